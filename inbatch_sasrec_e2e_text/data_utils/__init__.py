"""`from data_utils import ...` (inbatch_sasrec_e2e_text/run.py:14-16, data_utils/__init__.py): the same names."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from idvs.morec_b200.host.utils import *  # noqa: E402,F401,F403
from idvs.morec_b200.host.utils import (get_checkpoint, get_time, latest_checkpoint, para_and_log, report_time_eval,  # noqa: E402,F401
                                        report_time_train, save_model, setuplogger, str2bool)
from idvs.morec_b200.host.preprocess import read_news, read_news_bert, get_doc_input_bert, read_behaviors  # noqa: E402,F401
from idvs.morec_b200.host.dataset import BuildTrainDataset, BuildEvalDataset, SequentialDistributedSampler, DeviceBatcher  # noqa: E402,F401
from idvs.morec_b200.host.metrics import eval_model, get_item_embeddings, eval_ranks, metrics_topK  # noqa: E402,F401
