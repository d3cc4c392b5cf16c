"""2-GPU NCCL test of the `global` multi-GPU mode (needs >= 2 B200s: `gpurun --gpus 2`): the two ranks together must
reproduce the single-process CUDA step at batch 2B -- loss and DDP-averaged parameter gradients -- and the `local`
mode must reproduce the per-rank single-process step (reference DDP semantics)."""
import os
import socket
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(seed, N, T, D, L, dev):
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.model import Model
    torch.manual_seed(seed)
    cfg = BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, vocab_size=1024,
                     max_position_embeddings=64)
    a = types.SimpleNamespace(max_seq_len=L, embedding_dim=D, num_attention_heads=2, drop_rate=0.1, transformer_block=2,
                              num_words_title=T, num_words_abstract=0, num_words_body=0, news_attributes=["title"],
                              bert_model_load="bert_tiny", word_embedding_dim=128)
    bert = BertModel(cfg)
    for n, p in bert.named_parameters():      # run.py:73-75 freezes the (unused) pooler; DDP then sees no unused params
        if "pooler" in n:
            p.requires_grad = False
    return a, bert


def _worker(rank, world, port, q, use_wrap, host_plan=False, one_item=False):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from idvs.morec_b200.model import Model
    from idvs.morec_b200.synth import synth_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, L, N, T, D = 6, 8, 40, 12, 64
        full = synth_batch(world * B, L, N, T, seed=77, modal=True, n_users_pop=100, vocab_lo=10, vocab_hi=900)
        if one_item:      # every real slot holds the SAME item: n_unique = 1 < G, rank 1 owns no item of the global batch
            flat = full["ids"].reshape(-1)
            k = int(torch.nonzero(flat)[0])
            real = flat != 0
            full["items"][real] = full["items"][k].clone()
            flat[real] = int(flat[k])
        a, bert = _build(5, N, T, D, L, dev)
        model = Model(a, N, True, bert, full["pop_prob"].numpy()).to(dev).eval()
        model.parallel_mode = "global"
        if host_plan:     # the step is planned from host ids + the catalogue's token counts: no device->host wait
            model.set_item_content(full["item_content"])
        if use_wrap:      # tower gradients averaged layer by layer inside its backward (ops._GradSync), rest by DDP
            from idvs.morec_b200.parallel import wrap_ddp
            ddp = wrap_ddp(model, rank)
        else:             # stock DistributedDataParallel exactly as run.py:148 constructs it
            ddp = DDP(model, device_ids=[rank], output_device=rank, find_unused_parameters=True)
        sl = slice(rank * B, (rank + 1) * B)
        ids, lm = full["ids"][sl].to(dev), full["log_mask"][sl].to(dev)
        items = full["items"].view(world * B, L + 1, -1)[sl].reshape(B * (L + 1), -1).to(dev)
        loss = ddp(ids.reshape(-1), items, lm, rank, host_ids=full["ids"][sl].numpy() if host_plan else None)
        loss.backward()
        mean_loss = loss.detach().clone()
        dist.all_reduce(mean_loss)
        mean_loss /= world
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        ok, msg = True, ""
        if rank == 0:
            a2, bert2 = _build(5, N, T, D, L, dev)
            ref = Model(a2, N, True, bert2, full["pop_prob"].numpy()).to(dev).eval()
            ref.load_state_dict(model.state_dict())
            l2 = ref(full["ids"].reshape(-1).to(dev), full["items"].to(dev), full["log_mask"].to(dev), 0)
            l2.backward()
            if abs(float(mean_loss) - float(l2)) > 1e-4:
                ok, msg = False, f"loss {float(mean_loss)} vs {float(l2)}"
            for n, p in ref.named_parameters():
                if p.grad is None:
                    continue
                sc = float(p.grad.abs().max()) + 1e-12
                err = float((grads[n] - p.grad).abs().max())
                if err > 2e-3 * sc + 1e-7:
                    ok, msg = False, f"{n}: {err} vs scale {sc}"
                    break
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_wrap,host_plan,one_item", [(False, False, False), (True, False, False), (True, True, False),
                                                         (True, True, True), (True, False, True)])
def test_global_mode_two_gpus_equals_single_process(use_wrap, host_plan, one_item):
    """host_plan: ids exchanged between the hosts (gloo), no device->host wait in the step.  one_item: a rank that owns
    no distinct item of the global batch still issues every gradient collective (formerly a hang)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, use_wrap, host_plan, one_item)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, ok, msg in res:
        assert ok, (rank, msg)
