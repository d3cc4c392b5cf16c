"""Multi-GPU plumbing of the MoRec step: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

Two modes:

`local`  (reference-exact, inbatch_sasrec_e2e_text/run.py:148): every rank owns B users and their C = B(L+1) item
         slots, negatives are rank-local, the loss is the mean over the rank's valid rows and parameter gradients
         are averaged by DistributedDataParallel's bucketed all-reduce.  No collective inside Model.forward.

`global` (north star): the G ranks behave like ONE process with batch G*B.
         1. all-gather of the int64 item ids and of the items' token rows (tiny: C*(1+2T)*8 bytes per rank),
         2. `plan_global_batch` (pure index arithmetic, identical on every rank): global unique non-pad items,
            dealt round-robin to ranks -> every item of the global batch is encoded exactly ONCE across the job,
         3. each rank runs the text tower on its share, then ONE NCCL all-gather of the item embeddings
            ([G*U_max, D]); the backward of that all-gather is a reduce-scatter (AllGatherRowsFn),
         4. columns of the scoring matrix = all G*C slots (gathered embeddings expanded by slot), rows = the rank's
            own B*L positions (col_offset = rank*C): softmax denominators are complete locally,
         5. the loss is normalised by the GLOBAL valid-row count (one 1-element all-reduce) and multiplied by G so
            that DDP's 1/G gradient averaging reproduces the single-process gradient at batch G*B.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP


class MorecDDP(DDP):
    """DistributedDataParallel whose no_sync() also suspends the text tower's own layer-wise gradient all-reduce
    (ops._GradSync), so gradient accumulation behaves as with stock DDP."""

    def no_sync(self):
        import contextlib
        outer = super().no_sync()
        mod = self.module

        @contextlib.contextmanager
        def ctx():
            prev = getattr(mod, "_grad_sync_suspended", False)
            mod._grad_sync_suspended = True
            try:
                with outer:
                    yield
            finally:
                mod._grad_sync_suspended = prev
        return ctx()


def wrap_ddp(model, local_rank, overlap_grad_sync=True, **ddp_kwargs):
    """DDP wrapper of a morec Model (run.py:148).  With overlap_grad_sync the text tower averages its own gradients
    layer by layer during its backward (Model.enable_overlap_grad_sync) and DDP handles the rest."""
    if overlap_grad_sync and getattr(model, "use_modal", False) and hasattr(model, "enable_overlap_grad_sync"):
        model.enable_overlap_grad_sync()
    kw = dict(find_unused_parameters=False, gradient_as_bucket_view=True, bucket_cap_mb=100)
    # gradient_as_bucket_view avoids one copy of the 462 MB fp32 gradient set per step
    kw.update(ddp_kwargs)
    if local_rank is None or (isinstance(local_rank, str) and local_rank == "cpu"):
        return MorecDDP(model, **kw)
    return MorecDDP(model, device_ids=[local_rank], output_device=local_rank, **kw)


@dataclass
class GlobalPlan:
    G: int
    rank: int
    C: int                     # slots per rank
    n_unique: int              # distinct non-pad items in the global batch
    u_max: int                 # rows each rank contributes to the all-gather (padded share size)
    my_first_slots: np.ndarray  # [n_mine] global slot index holding the content of each item this rank encodes
    slot_to_row: np.ndarray    # [G*C] row of the gathered table [G*u_max, D] for every global slot (-1 = pad slot)


def unique_first(x: np.ndarray):
    """(uniq, first, inv) == np.unique(x, return_index=True, return_inverse=True) for non-negative integer ids.
    When the id range is comparable to the number of slots (the usual case: in-batch ids of a catalogue) a direct
    table over the id range replaces np.unique's sort: O(n + max_id) instead of O(n log n) -- the global plan of an
    8-GPU step (13,312 slots) sits between the size-determining sync and the first kernel launch, where the GPU idles."""
    x = np.asarray(x)
    n = x.size
    if n == 0:
        return x[:0], np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    hi = int(x.max())
    if int(x.min()) < 0 or hi > 16 * n:                 # sparse id range: the sort is cheaper than touching the table
        return np.unique(x, return_index=True, return_inverse=True)
    first_of_id = np.full(hi + 1, -1, dtype=np.int64)
    first_of_id[x[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)     # last write wins = smallest slot index
    uniq = np.flatnonzero(first_of_id >= 0)
    rank_of_id = np.empty(hi + 1, dtype=np.int64)
    rank_of_id[uniq] = np.arange(uniq.size, dtype=np.int64)
    return uniq.astype(x.dtype, copy=False), first_of_id[uniq], rank_of_id[x]


def plan_global_batch(ids_all: np.ndarray, G: int, rank: int) -> GlobalPlan:
    """ids_all: int64 [G*C] item ids of every slot of the global batch (rank-major).  Deterministic, host-side,
    identical on every rank.  Unique non-pad items are sorted by id and dealt round-robin: item k -> rank k % G,
    local index k // G  (balances encoder work to within one item and removes cross-rank duplicates)."""
    ids_all = np.asarray(ids_all, dtype=np.int64).reshape(-1)
    C = ids_all.size // G
    nz = ids_all != 0
    uniq, first, inv = unique_first(ids_all[nz])
    nz_slots = np.nonzero(nz)[0]
    first_slots = nz_slots[first]                         # global slot of the first occurrence of each unique item
    n_unique = int(uniq.size)
    u_max = max((n_unique + G - 1) // G, 1)
    owner = np.arange(n_unique) % G
    local = np.arange(n_unique) // G
    slot_to_row = np.full(ids_all.size, -1, dtype=np.int64)
    slot_to_row[nz_slots] = (owner * u_max + local)[inv]
    mine = owner == rank
    return GlobalPlan(G=G, rank=rank, C=C, n_unique=n_unique, u_max=u_max, my_first_slots=first_slots[mine],
                      slot_to_row=slot_to_row)


class AllGatherRowsFn(torch.autograd.Function):
    """out[G*u_max, D] = all_gather(x[u_max, D]);  backward = reduce-scatter (sum) of the gathered-table gradient:
    the ONE NCCL all-gather of item embeddings named by the north star, and its matching collective."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        G = dist.get_world_size(group)
        x = x.contiguous()
        out = torch.empty((G * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        if dist.get_backend(group) == "gloo":              # CPU tests: no all_gather_into_tensor on gloo
            parts = [torch.empty_like(x) for _ in range(G)]
            dist.all_gather(parts, x, group=group)
            out.copy_(torch.cat(parts, dim=0))
        else:
            dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, dout):
        group = ctx.group
        G = dist.get_world_size(group)
        dout = dout.contiguous()
        n = dout.shape[0] // G
        if dist.get_backend(group) == "gloo":
            full = dout.clone()
            dist.all_reduce(full, group=group)
            r = dist.get_rank(group)
            return full[r * n:(r + 1) * n].clone(), None
        dx = torch.empty((n,) + tuple(dout.shape[1:]), device=dout.device, dtype=dout.dtype)
        dist.reduce_scatter_tensor(dx, dout, op=dist.ReduceOp.SUM, group=group)
        return dx, None


def all_gather_small(t: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of a small integer tensor along a new leading rank dimension"""
    G = dist.get_world_size(group)
    t = t.contiguous()
    if dist.get_backend(group) == "gloo":
        parts = [torch.empty_like(t) for _ in range(G)]
        dist.all_gather(parts, t, group=group)
        return torch.stack(parts, dim=0)
    out = torch.empty((G,) + tuple(t.shape), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t, group=group)
    return out
