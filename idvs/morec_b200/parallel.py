"""Multi-GPU plumbing of the MoRec step: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

`local` mode (reference-exact, inbatch_sasrec_e2e_text/run.py:148): every rank owns B users and their C = B(L+1)
item slots, negatives are rank-local, the loss is the mean over the rank's valid rows and parameter gradients are
averaged by DistributedDataParallel's bucketed all-reduce.  There is no collective inside Model.forward.
"""
import torch
from torch.nn.parallel import DistributedDataParallel as DDP


def wrap_ddp(model, local_rank):
    # gradient_as_bucket_view avoids one copy of the 462 MB fp32 gradient set per step
    return DDP(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False,
               gradient_as_bucket_view=True, bucket_cap_mb=100)
