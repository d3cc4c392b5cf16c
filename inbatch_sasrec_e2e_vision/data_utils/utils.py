"""`from data_utils.utils import *` (inbatch_sasrec_e2e_vision/run.py:13, parameters.py:1): the reference's utils module
also leaks `os`, `time`, `torch`, `argparse`, `math` and `logging` into its importers -- kept, run.py relies on it."""
import argparse  # noqa: F401
import logging  # noqa: F401
import math  # noqa: F401
import os  # noqa: F401
import sys
import time  # noqa: F401

import torch  # noqa: F401
import torch.distributed as dist  # noqa: F401

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from idvs.morec_b200.host.utils import (get_checkpoint, get_time, latest_checkpoint, para_and_log, report_time_eval,  # noqa: E402,F401
                                        report_time_train, save_model, setuplogger, str2bool)
