"""Datasets of the training / evaluation loops.

Host datasets keep the reference's class names, constructor arguments and item layout
(inbatch_sasrec_e2e_text/data_utils/dataset.py:10-93), so a DataLoader-based loop is unchanged:
`BuildTrainDataset`, `BuildEvalDataset`, `SequentialDistributedSampler`.

`DeviceBatcher` is the B200-first replacement of that DataLoader: the catalogue's content
(`item_content` int32 [N+1, 2T]) and every user's sequence live in HBM once; a batch is assembled ON THE DEVICE from a
list of user indices (left padding, log_mask, content gather), so the per-step host work is one small index upload
instead of B*(L+1)*2T int64 collated by worker processes and copied every step.  The batch tensors have exactly the
layout and dtypes the reference's collate produces (sample_items_id i64 [B, L+1], sample_items i64 [B, L+1, 2T] | [B, L+1],
log_mask f32 [B, L]).
"""
import math

import numpy as np
import torch
from torch.utils.data import Dataset

from .preprocess import pack_sequences


class BuildTrainDataset(Dataset):
    def __init__(self, u2seq, item_content, item_num, max_seq_len, use_modal):
        self.u2seq = u2seq
        self.item_content = item_content
        self.item_num = item_num
        self.max_seq_len = max_seq_len + 1
        self.use_modal = use_modal

    def __len__(self):
        return len(self.u2seq)

    def __getitem__(self, user_id):
        seq = self.u2seq[user_id]
        pad = self.max_seq_len - len(seq)
        ids = torch.zeros(self.max_seq_len, dtype=torch.long)
        ids[pad:] = torch.as_tensor(seq, dtype=torch.long)
        log_mask = torch.zeros(self.max_seq_len - 1, dtype=torch.float32)
        log_mask[pad:] = 1.0
        items = torch.as_tensor(self.item_content[ids.numpy()], dtype=torch.long) if self.use_modal else ids
        return ids, items, log_mask


class BuildEvalDataset(Dataset):
    """(user id, input item embeddings [L, D], log_mask [L], one-hot label [N]) -- the reference's eval sample;
    the morec eval path (host/metrics.py) does not build the dense one-hot, this class exists for API parity"""

    def __init__(self, u2seq, item_content, max_seq_len, item_num):
        self.u2seq = u2seq
        self.item_content = item_content
        self.max_seq_len = max_seq_len + 1
        self.item_num = item_num

    def __len__(self):
        return len(self.u2seq)

    def __getitem__(self, user_id):
        seq = self.u2seq[user_id]
        tokens, target = seq[:-1], seq[-1]
        pad = self.max_seq_len - len(seq)
        padded = [0] * pad + list(tokens)
        log_mask = torch.zeros(len(padded), dtype=torch.float32)
        log_mask[pad:] = 1.0
        labels = np.zeros(self.item_num)
        labels[target - 1] = 1.0
        return torch.LongTensor([user_id]), self.item_content[padded], log_mask, labels


class SequentialDistributedSampler(torch.utils.data.sampler.Sampler):
    """contiguous, padded shards in order (rank r gets samples [r*n, (r+1)*n), the tail repeats the last index)"""

    def __init__(self, dataset, batch_size, rank=None, num_replicas=None):
        import torch.distributed as dist
        if num_replicas is None:
            num_replicas = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.dataset, self.batch_size, self.rank, self.num_replicas = dataset, batch_size, rank, num_replicas
        self.num_samples = int(math.ceil(len(dataset) / batch_size / num_replicas)) * batch_size
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        n = len(self.dataset)
        idx = np.minimum(np.arange(self.rank * self.num_samples, (self.rank + 1) * self.num_samples), n - 1)
        return iter(idx.tolist())

    def __len__(self):
        return self.num_samples


class DeviceBatcher:
    """Device-resident sequences (+ item content) and on-device batch assembly (see the module docstring)."""

    def __init__(self, u2seq, item_content, max_seq_len, use_modal, device):
        flat, ptr = pack_sequences(u2seq)
        self.device = torch.device(device)
        self.n_users = len(u2seq)
        self.Lp1 = max_seq_len + 1
        self.use_modal = use_modal
        self.flat = torch.from_numpy(flat).to(self.device)
        self.ptr = torch.from_numpy(ptr).to(self.device)
        self.flat_host, self.ptr_host = flat, ptr.astype(np.int64)
        self.content = None
        if use_modal:
            self.content = torch.as_tensor(np.asarray(item_content), dtype=torch.int64).to(self.device)   # [N+1, 2T]
        self._ar = torch.arange(self.Lp1, device=self.device, dtype=torch.int64)

    def __len__(self):
        return self.n_users

    def epoch_order(self, epoch, rank, world, seed=0):
        """DistributedSampler's partition (torch.utils.data.distributed.DistributedSampler.__iter__, shuffle=True,
        drop_last=False): one seeded permutation, padded by wrapping, strided by rank"""
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(self.n_users, generator=g)
        total = int(math.ceil(self.n_users / world)) * world
        if total > idx.numel():
            idx = torch.cat([idx, idx[:total - idx.numel()]])
        return idx[rank:total:world]

    def host_ids(self, user_idx):
        """the batch's item ids [B, L+1] as a host array (same arithmetic as batch(), in numpy): lets the model plan
        the step without a device->host wait (Model.forward(..., host_ids=))"""
        u = np.asarray(user_idx, dtype=np.int64)
        start, end = self.ptr_host[u], self.ptr_host[u + 1]
        n = np.minimum(end - start, self.Lp1)
        pos = np.arange(self.Lp1, dtype=np.int64)[None, :] - (self.Lp1 - n)[:, None]
        src = (end - n)[:, None] + pos
        return np.where(pos >= 0, self.flat_host[np.maximum(src, 0)].astype(np.int64), 0)

    def batch(self, user_idx):
        """user_idx: int64 tensor (host or device) -> (sample_items_id [B, L+1], sample_items, log_mask [B, L])"""
        u = user_idx.to(self.device, non_blocking=True)
        start, end = self.ptr[u].long(), self.ptr[u + 1].long()
        n = (end - start).clamp(max=self.Lp1)
        pos = self._ar.view(1, -1) - (self.Lp1 - n).view(-1, 1)              # position inside the sequence, < 0 = pad
        src = (end - n).view(-1, 1) + pos                                    # the LAST n items of the sequence
        ids = torch.where(pos >= 0, self.flat[src.clamp(min=0)].long(), torch.zeros((), dtype=torch.int64, device=self.device))
        log_mask = (ids[:, :-1] != 0).to(torch.float32)
        items = self.content[ids] if self.use_modal else ids
        return ids, items, log_mask
