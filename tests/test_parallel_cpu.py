"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the `global` mode's batch plan, the embedding
all-gather with its reduce-scatter backward, and the loss / gradient decomposition over ranks -- checked against
the single-process oracle at batch G*B.  The arithmetic is done by the oracle here (no GPU); the
CUDA kernels that consume the same plan are covered by tests/test_multigpu_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import morec_oracle as O
from idvs.morec_b200 import parallel as par
from idvs.morec_b200.synth import synth_batch


def test_plan_partition_properties():
    g = np.random.default_rng(0)
    G, C = 4, 60
    ids = g.integers(0, 25, size=G * C)
    plans = [par.plan_global_batch(ids, G, r) for r in range(G)]
    uniq = np.unique(ids[ids != 0])
    assert all(p.n_unique == uniq.size for p in plans)
    # every unique item is encoded by exactly one rank; shares differ by at most one item
    firsts = np.concatenate([p.my_first_slots for p in plans])
    assert sorted(ids[firsts].tolist()) == uniq.tolist()
    sizes = [p.my_first_slots.size for p in plans]
    assert max(sizes) - min(sizes) <= 1 and max(sizes) <= plans[0].u_max
    # slot_to_row is identical on every rank and reconstructs the ids through the gathered table layout
    table = np.zeros(G * plans[0].u_max, dtype=np.int64)
    for r, p in enumerate(plans):
        table[r * p.u_max: r * p.u_max + p.my_first_slots.size] = ids[p.my_first_slots]
        assert np.array_equal(p.slot_to_row, plans[0].slot_to_row)
    s2r = plans[0].slot_to_row
    assert np.array_equal(np.where(s2r >= 0, table[np.maximum(s2r, 0)], 0), ids)
    # edge cases: all pad, single item, G larger than the number of unique items
    p = par.plan_global_batch(np.zeros(8, dtype=np.int64), 2, 0)
    assert p.n_unique == 0 and p.u_max == 1 and (p.slot_to_row == -1).all()
    p = par.plan_global_batch(np.array([0, 7, 7, 0]), 4, 3)
    assert p.n_unique == 1 and p.my_first_slots.size == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B, L, N, D = 5, 6, 30, 16
        batches = [synth_batch(B, L, N, 0, seed=100 + r, modal=False, n_users_pop=50) for r in range(world)]
        pop = batches[0]["pop_prob"]
        table = torch.randn(N + 1, D, dtype=torch.float64, requires_grad=True)       # "encoder": item id -> embedding
        W = torch.randn(D, D, dtype=torch.float64, requires_grad=True)               # "user encoder": P = X @ W
        mine = batches[rank]
        ids_flat = mine["ids"].reshape(-1)
        C = ids_flat.numel()
        ids_all = par.all_gather_small(ids_flat).reshape(-1)
        plan = par.plan_global_batch(ids_all.numpy(), world, rank)
        # each rank "encodes" only its share, pads to u_max, all-gathers (autograd: reduce-scatter in backward)
        my_ids = ids_all[torch.from_numpy(plan.my_first_slots)]
        E_pad = torch.zeros(plan.u_max, D, dtype=torch.float64)
        E_pad = torch.cat([table[my_ids], E_pad[my_ids.numel():]], dim=0)
        E_table = par.AllGatherRowsFn.apply(E_pad, dist.group.WORLD)
        s2r = torch.from_numpy(plan.slot_to_row)
        E_slots = torch.where((s2r >= 0).view(-1, 1), E_table[s2r.clamp(min=0)], torch.zeros((), dtype=torch.float64))
        assert torch.allclose(E_slots.detach(), torch.where((ids_all != 0).view(-1, 1), table.detach()[ids_all],
                                                           torch.zeros((), dtype=torch.float64)))
        # local rows vs global columns
        rows = (torch.arange(B).view(B, 1) * (L + 1) + torch.arange(L).view(1, L)).reshape(-1) + rank * C
        P = E_slots[rows] @ W
        logp = torch.log(pop.float()[ids_all]).double()
        S = P @ E_slots.t() - logp.view(1, -1)
        own = mine["ids"]
        member = (ids_all.view(1, 1, -1) == own.view(B, L + 1, 1)).any(1)          # [B, G*C]
        member = member.view(B, 1, -1).expand(B, L, -1).reshape(B * L, -1).clone()
        tgt = O.ce_labels(B, L) + rank * C
        member[torch.arange(B * L), tgt] = False
        masked = member | (ids_all == 0).view(1, -1)
        S = torch.where(masked, torch.full((), O.NEG_MASK, dtype=torch.float64), S)
        valid = O.valid_rows(mine["log_mask"])
        row = torch.logsumexp(S, 1) - S[torch.arange(B * L), tgt]
        s_local = (row * valid.double()).sum()
        n = valid.double().sum().reshape(1)
        dist.all_reduce(n)
        loss_rank = s_local / n.reshape(()) * world                                   # what Model._forward_global returns
        loss_rank.backward()
        # DDP would average gradients over ranks
        for p in (table, W):
            dist.all_reduce(p.grad)
            p.grad /= world
        mean_loss = loss_rank.detach().clone().reshape(1)
        dist.all_reduce(mean_loss)
        mean_loss /= world
        # single-process oracle at batch G*B
        ids_cat = torch.cat([b["ids"] for b in batches], 0)
        lm_cat = torch.cat([b["log_mask"] for b in batches], 0)
        t2 = table.detach().clone().requires_grad_(True)
        W2 = W.detach().clone().requires_grad_(True)
        E2 = torch.where((ids_cat.reshape(-1) != 0).view(-1, 1), t2[ids_cat.reshape(-1)], torch.zeros((), dtype=torch.float64))
        X2 = E2.view(world * B, L + 1, D)[:, :-1].reshape(-1, D)
        loss2, _, _, _ = O.inbatch_ce(X2 @ W2, E2, ids_cat, torch.log(pop.float()[ids_cat.reshape(-1)]).double(), lm_cat)
        loss2.backward()
        ok = (abs(float(mean_loss) - float(loss2)) < 1e-10 and torch.allclose(table.grad, t2.grad, atol=1e-10)
              and torch.allclose(W.grad, W2.grad, atol=1e-10))
        q.put((rank, bool(ok), float(mean_loss), float(loss2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_global_mode_equals_single_process_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, a, b in res:
        assert ok, (rank, a, b)


def _gradsync_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from idvs.morec_b200 import ops
        shapes = [(7, 3), (5,), (4, 4), (9,), (2, 6)]
        arena = ops._Arena(torch.device("cpu"), shapes)
        assert ops._GradSync.enabled({"grad_sync": True}) and not ops._GradSync.enabled({})
        sync = ops._GradSync(arena)
        vals = []
        for i, sh in enumerate(shapes):                 # "layers": take a slice, fill it, flush after every second one
            g = arena.take(sh)
            g.copy_(torch.full(sh, float((rank + 1) * (i + 1))))
            vals.append(g)
            if i % 2 == 1:
                sync.flush()
        sync.wait()                                      # flushes the tail
        ok = all(torch.allclose(g, torch.full_like(g, (i + 1) * (1 + world) / 2.0)) for i, g in enumerate(vals))
        sync.wait()                                      # idempotent: nothing left to reduce
        ok = ok and all(torch.allclose(g, torch.full_like(g, (i + 1) * (1 + world) / 2.0)) for i, g in enumerate(vals))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_layerwise_gradient_sync_world2():
    """ops._GradSync: slices of the gradient arena are averaged over ranks exactly once, in flush order"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gradsync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_overlap_grad_sync_is_explicit():
    """the tower's own gradient averaging is switched on explicitly (enable_overlap_grad_sync / parallel.wrap_ddp);
    a bare Model never starts collectives and attribute introspection has no side effects"""
    import types
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.model import Model
    cfg = BertConfig(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64, vocab_size=100,
                     max_position_embeddings=16)
    a = types.SimpleNamespace(max_seq_len=4, embedding_dim=16, num_attention_heads=2, drop_rate=0.0, transformer_block=1,
                              num_words_title=8, num_words_abstract=0, num_words_body=0, news_attributes=["title"],
                              bert_model_load="bert_tiny", word_embedding_dim=32)
    m = Model(a, 10, True, BertModel(cfg), np.ones(11) / 11)
    import inspect
    assert m._overlap_grad_sync is False and not hasattr(m, "_ddp_params_and_buffers_to_ignore")
    inspect.getmembers(m)                                        # introspection must not flip anything
    assert m._overlap_grad_sync is False and not hasattr(m, "_ddp_params_and_buffers_to_ignore")
    m.enable_overlap_grad_sync()                                 # (no process group here: nothing to broadcast)
    names = m._ddp_params_and_buffers_to_ignore
    assert m._overlap_grad_sync is True
    all_names = [n for n, _ in m.named_parameters()] + [n for n, _ in m.named_buffers()]
    assert names and all(n in all_names and n.startswith("bert_encoder.") for n in names)
    assert not any(n.startswith("user_encoder.") for n in names)
    assert "_ddp_params_and_buffers_to_ignore" not in m.state_dict()
    m_id = Model(a, 10, False, None, np.ones(11) / 11)          # ID tower: nothing to ignore
    m_id.enable_overlap_grad_sync()
    assert m_id._overlap_grad_sync is False and not hasattr(m_id, "_ddp_params_and_buffers_to_ignore")


def test_unique_first_matches_numpy():
    g = np.random.default_rng(3)
    for n, hi in [(1, 5), (50, 7), (1664, 50000), (13312, 50000), (40, 10 ** 9)]:     # the last one takes the np.unique path
        x = g.integers(1, hi + 1, size=n)
        u0, f0, i0 = np.unique(x, return_index=True, return_inverse=True)
        u1, f1, i1 = par.unique_first(x)
        assert np.array_equal(u0, u1) and np.array_equal(f0, f1) and np.array_equal(i0, i1)
    u, f, i = par.unique_first(np.zeros(0, dtype=np.int64))
    assert u.size == 0 and f.size == 0 and i.size == 0
