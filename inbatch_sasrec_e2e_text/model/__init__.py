"""`from model import Model` (inbatch_sasrec_e2e_text/run.py:13): the B200-native drop-in of model/model.py:7-69."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from idvs.morec_b200.model import Model  # noqa: E402,F401
from idvs.morec_b200.model.encoders import Bert_Encoder, Text_Encoder, User_Encoder  # noqa: E402,F401
