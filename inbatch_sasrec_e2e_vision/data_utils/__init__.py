"""`from data_utils import ...` (inbatch_sasrec_e2e_vision/run.py:10-12, data_utils/__init__.py): the same names.
`lmdb` is only imported when an image database is actually opened (the reference imports it at module level,
data_utils/dataset.py:10, and the package is absent from this image)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from idvs.morec_b200.host.utils import (get_checkpoint, get_time, latest_checkpoint, para_and_log, report_time_eval,  # noqa: E402,F401
                                        report_time_train, save_model, setuplogger, str2bool)
from idvs.morec_b200.host.preprocess import read_images, read_behaviors_vision as read_behaviors  # noqa: E402,F401
from idvs.morec_b200.host.vision_data import LMDB_Image, Build_Lmdb_Dataset, Build_Id_Dataset, ImageStore  # noqa: E402,F401
from idvs.morec_b200.host.dataset import BuildEvalDataset, SequentialDistributedSampler, DeviceBatcher  # noqa: E402,F401
from idvs.morec_b200.host.metrics import eval_model, eval_ranks, metrics_topK  # noqa: E402,F401
from idvs.morec_b200.host.metrics import get_item_embeddings  # noqa: E402,F401


def get_itemId_embeddings(model, item_num, test_batch_size, args, local_rank):
    """ID tower (data_utils/metrics.py:50-61)"""
    return get_item_embeddings(model, None, test_batch_size, args, False, local_rank)


def get_itemLMDB_embeddings(model, item_num, item_id_to_keys, lmdb_data, test_batch_size, args, local_rank):
    """image tower (data_utils/metrics.py:64-77): every catalogue image through model.module.cv_encoder, sharded over
    the ranks; images are read from the LMDB store batch by batch"""
    args.lmdb_data = lmdb_data
    return get_item_embeddings(model, item_id_to_keys, test_batch_size, args, True, local_rank)
