"""Evaluation: catalogue encoding + full-catalogue ranking (Hit@10 / nDCG@10) with the reference's entry points
(inbatch_sasrec_e2e_text/data_utils/metrics.py:60-107): `get_item_embeddings`, `eval_model`, `metrics_topK`.

What changed underneath:
  * the catalogue is encoded ONCE ACROSS THE JOB: every rank encodes a contiguous 1/G share and one all-gather
    assembles the table (the reference makes every rank encode everything, metrics.py:60-74), and the table stays on
    the device (the reference moves it to the CPU and back);
  * per eval batch the reference runs, for every user, an argsort over all N items against a dense N-long one-hot label
    built on the host (dataset.py:60-61, metrics.py:49-57, 96-101); here ONE tcgen05 GEMM per batch counts the items
    ranked before the target in its epilogue (csrc/eval_rank.cu) -- no [U, N] score matrix, no sort, no one-hot.
rank(u) = 1 + #{items not in history(u) scoring above the target} (ties as in a stable descending sort).
"""
import math
import os

import numpy as np
import torch
import torch.distributed as dist

from .. import lib
from .dataset import SequentialDistributedSampler


def _dist():
    on = dist.is_available() and dist.is_initialized()
    return (dist.get_rank(), dist.get_world_size()) if on else (0, 1)


def _module(model):
    return model.module if hasattr(model, "module") else model


def print_metrics(x, Log_file, v_or_t):
    Log_file.info(v_or_t + "_results   {}".format('\t'.join(["{:0.5f}".format(i * 100) for i in x])))


def metrics_topK(rank, topK, local_rank=None):
    """(hit, ndcg) of one user from its 1-based rank -- the tail of the reference's metrics_topK (metrics.py:53-57)"""
    out = torch.zeros(2)
    if rank <= topK:
        out[0] = 1
        out[1] = 1 / math.log2(rank + 1)
    return out


@torch.no_grad()
def get_item_embeddings(model, item_content, test_batch_size, args, use_modal, local_rank):
    """-> [N+1, D] embedding of every catalogue item (row 0 = the pad item), ON THE DEVICE.  Sharded over the ranks of
    the job; fp32 parity arithmetic unless args.eval_dtype says otherwise (the reference evaluates outside autocast)."""
    m = _module(model)
    m.eval()
    dev = torch.device("cuda", local_rank) if isinstance(local_rank, int) else torch.device(local_rank)
    rank, world = _dist()
    prev = m.compute_dtype
    m.set_compute_dtype(getattr(args, "eval_dtype", "fp32"))
    try:
        if not use_modal:
            return m.id_embedding.weight.detach().to(torch.float32).clone()
        store = txn = None
        if isinstance(item_content, dict):            # vision, real data: {item id: LMDB key} (V/data_utils/metrics.py:64-77)
            from .vision_data import ImageStore
            store = ImageStore(os.path.join(args.root_data_dir, args.dataset, args.lmdb_data), args.CV_resize)
            txn = store.env.begin()
            n = len(item_content) + 1
        else:
            content = item_content if torch.is_tensor(item_content) else torch.as_tensor(np.asarray(item_content))
            n = content.shape[0]
        share = (n + world - 1) // world
        lo, hi = min(rank * share, n), min((rank + 1) * share, n)
        D = args.embedding_dim
        mine = torch.zeros(share, D, device=dev, dtype=torch.float32)
        enc = m.bert_encoder if hasattr(m, "bert_encoder") else m.cv_encoder
        for s in range(lo, hi, test_batch_size):
            e = min(s + test_batch_size, hi)
            if store is not None:
                x = torch.stack([store.get(txn, item_content[i]) if i > 0 else torch.zeros(3, args.CV_resize, args.CV_resize)
                                 for i in range(s, e)]).to(dev)
            else:
                x = content[s:e].to(dev)
                x = x.float() if x.is_floating_point() else x.long()
            mine[s - lo:e - lo] = enc(x).float()
        if world == 1:
            return mine[:n]
        table = torch.empty(world * share, D, device=dev, dtype=torch.float32)
        dist.all_gather_into_tensor(table, mine)
        return table[:n].contiguous()
    finally:
        m.set_compute_dtype(prev)


def _pack_eval(eval_seq, user_history, users, max_seq_len):
    """padded input item ids [U, L], log_mask [U, L], targets [U] and the CSR history of the given users (host numpy)"""
    U, L = len(users), max_seq_len
    ids = np.zeros((U, L), dtype=np.int32)
    tgt = np.zeros(U, dtype=np.int32)
    hist, ptr = [], np.zeros(U + 1, dtype=np.int32)
    for i, u in enumerate(users):
        seq = eval_seq[u]
        tokens = seq[:-1][-L:]
        if tokens:
            ids[i, L - len(tokens):] = tokens
        tgt[i] = seq[-1]
        h = user_history[u]
        h = h.numpy() if torch.is_tensor(h) else np.asarray(h)
        hist.append(h.astype(np.int64))
        ptr[i + 1] = ptr[i] + h.size
    return ids, tgt, (np.concatenate(hist) if hist else np.zeros(0, dtype=np.int64)), ptr


@torch.no_grad()
def eval_ranks(model, user_history, eval_seq, item_embeddings, test_batch_size, args, local_rank, users=None):
    """1-based rank of every user's held-out item (int64 tensor, in `users` order, on the device)"""
    m = _module(model)
    m.eval()
    dev = item_embeddings.device
    L = args.max_seq_len
    users = list(range(len(eval_seq))) if users is None else list(users)
    E = item_embeddings.to(torch.float32).contiguous()
    prev = m.compute_dtype
    m.set_compute_dtype(getattr(args, "eval_dtype", "fp32"))
    out = []
    try:
        for s in range(0, len(users), test_batch_size):
            chunk = users[s:s + test_batch_size]
            ids, tgt, hist, ptr = _pack_eval(eval_seq, user_history, chunk, L)
            ids_d = lib.h2d(ids, dev)
            log_mask = (ids_d != 0).to(torch.float32)
            X = lib.gather_rows(E, ids_d.reshape(-1))                               # [U*L, D] input embeddings
            prec = m.user_encoder(X.view(len(chunk), L, -1), log_mask, local_rank)[:, -1].float().contiguous()
            hist_t = lib.h2d(hist, dev) if hist.size else torch.zeros(0, dtype=torch.int64, device=dev)
            bits = lib.eval_hist_bits(lib.h2d(ptr, dev), hist_t, E.shape[0])
            tgt_d = lib.h2d(tgt, dev)
            with lib.fp32_mode(True):
                rank, _, _ = lib.eval_rank(prec, E, bits, tgt_d)
            in_hist = ((bits.view(-1)[torch.arange(len(chunk), device=dev) * bits.shape[1] + (tgt_d.long() >> 5)]
                        >> (tgt_d & 31)) & 1) != 0
            rank = torch.where(in_hist, torch.full_like(rank, E.shape[0]), rank)    # masked target: never a hit
            out.append(rank.long())
    finally:
        m.set_compute_dtype(prev)
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.int64, device=dev)


@torch.no_grad()
def eval_model(model, user_history, eval_seq, item_embeddings, test_batch_size, args, item_num, Log_file, v_or_t,
               local_rank):
    """mean Hit@10 over all users (the reference's return value); logs Hit10 / nDCG10 like metrics.py:87,105"""
    topK = 10
    Log_file.info(v_or_t + "_methods   {}".format('\t'.join(['Hit{}'.format(topK), 'nDCG{}'.format(topK)])))
    rank_id, world = _dist()
    n_users = len(eval_seq)
    sampler = SequentialDistributedSampler(range(n_users), batch_size=test_batch_size, rank=rank_id, num_replicas=world)
    users = list(iter(sampler))
    ranks = eval_ranks(model, user_history, eval_seq, item_embeddings, test_batch_size, args, local_rank, users)
    hit = (ranks <= topK).float()
    ndcg = torch.where(ranks <= topK, 1.0 / torch.log2(ranks.float() + 1.0), torch.zeros((), device=ranks.device))
    res = torch.stack([hit, ndcg], 0)                                               # [2, n_local]
    if world > 1:
        parts = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(parts, res.contiguous())
        res = torch.cat(parts, dim=1)
    res = res[:, :n_users]
    mean_eval = [float(res[0].mean()), float(res[1].mean())]
    print_metrics(mean_eval, Log_file, v_or_t)
    return mean_eval[0]
