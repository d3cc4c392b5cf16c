"""run.py of the text package -- started by the reference's launchers unchanged, e.g.
    python -m torch.distributed.launch --nproc_per_node 2 --master_port 1234 run.py --root_data_dir ... --item_tower modal ...
(inbatch_sasrec_e2e_text/train_*.py).  Same flags, log lines and checkpoint layout as the reference's run.py; the
training step, evaluation and optimizer run on the morec_b200 CUDA kernels (idvs/morec_b200/host/train.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from parameters import parse_args  # noqa: E402,F401
from model import Model  # noqa: E402,F401
from data_utils.utils import *  # noqa: E402,F401,F403
from idvs.morec_b200.host.train import main, run_eval, setup_seed, train  # noqa: E402,F401

if __name__ == "__main__":
    main(kind="text")
