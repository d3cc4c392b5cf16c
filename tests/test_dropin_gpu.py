"""Drop-in boundary on the GPU: the reference-shaped training loop of run.py:231-249 (autocast,
GradScaler, DistributedDataParallel(find_unused_parameters=True), two-group AdamW) drives the morec Model with both
torch.optim.AdamW and FusedAdamW; the eval entry points of data_utils/metrics.py work under no_grad; the packaged
run.py trains, evaluates, checkpoints and resumes end to end; the fused rank kernel reproduces the reference's
per-user argsort procedure; the BCE head matches the reference."""
import os
import socket
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny(L=8, D=64, T=12, N=120, drop=0.0, seed=3):
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.model import Model
    from idvs.morec_b200.synth import synth_batch
    torch.manual_seed(seed)
    cfg = BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, vocab_size=1024,
                     max_position_embeddings=64, hidden_dropout_prob=drop, attention_probs_dropout_prob=drop)
    bert = BertModel(cfg)
    for i, (n, p) in enumerate(bert.named_parameters()):            # run.py:73-75
        if i in (37, 38):
            p.requires_grad = False
    a = types.SimpleNamespace(max_seq_len=L, embedding_dim=D, num_attention_heads=2, drop_rate=drop, transformer_block=2,
                              num_words_title=T, num_words_abstract=50, num_words_body=50, news_attributes=["title"],
                              bert_model_load="bert_tiny", word_embedding_dim=128)
    batches = [synth_batch(16, L, N, T, seed + i, modal=True, n_users_pop=100, vocab_lo=10, vocab_hi=900) for i in range(4)]
    from idvs.morec_b200.synth import pop_from_batches
    model = Model(a, N, True, bert, pop_from_batches(batches).numpy())
    return a, model, batches


@pytest.mark.parametrize("opt_kind", ["torch", "fused"])
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_reference_shaped_loop(opt_kind, mode):
    """the loop body of inbatch_sasrec_e2e_text/run.py:231-249, verbatim in structure"""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from idvs.morec_b200.optim import FusedAdamW
    if not dist.is_initialized():
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(_free_port())
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    local_rank = 0
    a, model, batches = _tiny()
    model = model.to(local_rank)
    model.set_compute_dtype(mode)
    model = DDP(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=True)       # run.py:148
    bert_params = [p for n, p in model.module.named_parameters() if p.requires_grad and 'bert_model' in n]
    recsys_params = [p for n, p in model.module.named_parameters() if p.requires_grad and 'bert_model' not in n]
    groups = [{'params': bert_params, 'lr': 1e-4, 'weight_decay': 0.01}, {'params': recsys_params, 'lr': 1e-3, 'weight_decay': 0.01}]
    optimizer = torch.optim.AdamW(groups) if opt_kind == "torch" else FusedAdamW(groups)
    if opt_kind == "fused":
        model.module.attach_optimizer(optimizer)
    scaler = torch.cuda.amp.GradScaler()
    model.train()
    loss, losses = 0.0, []
    for it in range(6):
        d = batches[it % 2]
        sample_items_id, sample_items, log_mask = d["ids"].to(local_rank), d["items"].view(16, a.max_seq_len + 1, -1).to(local_rank), \
            d["log_mask"].to(local_rank)
        sample_items = sample_items.view(-1, sample_items.size(-1))
        sample_items_id = sample_items_id.view(-1)
        optimizer.zero_grad()
        with torch.cuda.amp.autocast():
            bz_loss = model(sample_items_id, sample_items, log_mask, local_rank)
            loss += bz_loss.data.float()
        scaler.scale(bz_loss).backward()
        scaler.step(optimizer)
        scaler.update()
        assert not torch.isnan(loss.data)
        losses.append(float(bz_loss))
    assert losses[4] < losses[0] and losses[5] < losses[1], losses          # the same two batches: the loss goes down
    # ---- eval entry points (data_utils/metrics.py:70, 95)
    model.eval()
    with torch.no_grad():
        emb = model.module.bert_encoder(batches[0]["item_content"][:40].to(local_rank))
        assert emb.shape == (40, a.embedding_dim) and torch.isfinite(emb.float()).all() and float(emb[0].abs().max()) == 0.0
        inp = emb[:32].float().view(4, a.max_seq_len, -1)
        prec = model.module.user_encoder(inp.to(emb.dtype), torch.ones(4, a.max_seq_len, device="cuda"), local_rank)[:, -1]
        assert prec.shape == (4, a.embedding_dim) and torch.isfinite(prec.float()).all()
    sd = model.module.state_dict()
    assert all(k.startswith(("user_encoder.", "bert_encoder.")) for k in sd)


def test_fused_and_torch_adamw_train_identically():
    """3 steps of the same data with torch.optim.AdamW and FusedAdamW (+ optimizer-written 16-bit weight copies in
    fp16 mode) leave the same parameters"""
    from idvs.morec_b200.optim import FusedAdamW
    res = {}
    for kind in ("torch", "fused"):
        a, model, batches = _tiny(seed=9)
        model = model.cuda().eval()                                        # dropout off: deterministic
        model.set_compute_dtype("fp16")
        ps = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(ps, lr=1e-3) if kind == "torch" else FusedAdamW(ps, lr=1e-3)
        if kind == "fused":
            model.attach_optimizer(opt)
        scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
        for it in range(3):
            d = batches[it]
            opt.zero_grad()
            l = model(d["ids"].reshape(-1).cuda(), d["items"].cuda(), d["log_mask"].cuda(), 0)
            scaler.scale(l).backward()
            scaler.step(opt)
            scaler.update()
        res[kind] = {n: p.detach().clone() for n, p in model.named_parameters()}
        if kind == "fused":       # the shadows the next forward would read equal a fresh cast of the parameters
            ss = model.user_encoder._shadows
            assert ss.plan is not None and ss.fresh_epoch >= 0
            for dst, own in zip(ss.plan.dst, ss.owners):
                src = torch.cat([p.detach() for p, _, _ in own], 0)
                assert torch.equal(dst, src.to(dst.dtype))
    # Adam's first updates are ~ +-lr per element whatever the gradient's size, so the (atomics-order) noise of a
    # near-zero gradient can flip single elements by up to 2*lr per step: compare the bulk tightly, the tail by that bound
    for n in res["torch"]:
        if n.endswith("attention.self.key.bias"):       # exactly-zero gradient (softmax shift invariance): pure noise
            continue
        diff = (res["torch"][n] - res["fused"][n]).abs()
        assert float(diff.mean()) <= 1e-4 and float(diff.max()) <= 6.5e-3, (n, float(diff.mean()), float(diff.max()))


def _run_pkg(pkg, cwd, extra, timeout=900):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, pkg, "run.py")] + extra, cwd=cwd, env=env, capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return r.stdout + r.stderr


def test_text_run_py_end_to_end_and_resume(tmp_path):
    """packaged run.py on a synthetic dataset: 2 epochs of training + evaluation + epoch-N.pt checkpoints, then a
    resumed third epoch from the saved checkpoint (run.py:132-144,192-194,210-212; data_utils/utils.py:107-114)"""
    common = ["--synthetic", "700,300", "--item_tower", "modal", "--bert_model_load", "bert_tiny", "--mode", "train",
              "--batch_size", "32", "--max_seq_len", "10", "--embedding_dim", "64", "--num_words_title", "12",
              "--freeze_paras_before", "0", "--lr", "1e-3", "--fine_tune_lr", "1e-4", "--logging_num", "2", "--news", "unused.tsv",
              "--label_screen", "t", "--eval_batch_size", "256"]
    out = _run_pkg("inbatch_sasrec_e2e_text", str(tmp_path), common + ["--epoch", "2", "--local_rank", "0"])
    assert "train_loss" in out and "train_results" in out and "max eval Hit10" in out
    ck_dirs = [d for d in os.listdir(tmp_path) if d.startswith("checkpoint_modal_bert_tiny_freeze_0")]
    assert ck_dirs, os.listdir(tmp_path)
    sub = os.path.join(tmp_path, ck_dirs[0])
    cpt = os.path.join(sub, os.listdir(sub)[0])
    files = sorted(os.listdir(cpt))
    assert files and all(f.startswith("epoch-") and f.endswith(".pt") for f in files), files
    ck = torch.load(os.path.join(cpt, files[-1]), map_location="cpu", weights_only=False)
    assert set(ck) == {"model_state_dict", "optimizer", "rng_state", "cuda_rng_state", "scaler_state"}
    st0 = next(iter(ck["optimizer"]["state"].values()))
    assert set(st0) == {"step", "exp_avg", "exp_avg_sq"} and float(st0["step"]) > 0           # torch.optim.AdamW layout
    out2 = _run_pkg("inbatch_sasrec_e2e_text", str(tmp_path), common + ["--epoch", "1", "--local-rank", "0",
                                                                       "--load_ckpt_name", files[-1]])
    n = int(files[-1][len("epoch-"):-3])
    assert f"epoch {n + 1} start" in out2 and "optimizer loaded from" in out2 and "train_loss" in out2


def test_id_tower_and_vision_run_py(tmp_path):
    out = _run_pkg("inbatch_sasrec_e2e_text", str(tmp_path), ["--synthetic", "500,200", "--item_tower", "id", "--mode", "train",
                                                             "--batch_size", "32", "--max_seq_len", "10", "--embedding_dim", "64",
                                                             "--lr", "1e-3", "--epoch", "1", "--news", "x", "--logging_num", "2"])
    assert "train_loss" in out and "train_results" in out
    out = _run_pkg("inbatch_sasrec_e2e_vision", str(tmp_path), ["--synthetic", "64,40", "--item_tower", "modal", "--mode", "train",
                                                               "--CV_model_load", "swin_tiny", "--batch_size", "4", "--max_seq_len", "3",
                                                               "--embedding_dim", "64", "--lr", "1e-3", "--fine_tune_lr", "1e-4",
                                                               "--epoch", "1", "--max_steps", "3", "--freeze_paras_before", "0",
                                                               "--logging_num", "1", "--eval_batch_size", "32"])
    assert "train_loss" in out and "train_results" in out


def test_eval_ranks_match_reference_eval_procedure():
    """host.metrics.eval_ranks (sharded catalogue table + fused rank kernel) == the reference's eval loop
    (data_utils/metrics.py:88-101: user_encoder(...)[:, -1], matmul, history -> -inf, [1:], argsort, rank of target)"""
    from idvs.morec_b200.host import metrics as M
    from idvs.morec_b200.host.preprocess import synthetic_dataset
    from idvs.morec_b200.model import Model
    torch.manual_seed(17)
    L, D, N = 10, 64, 400
    item_num, _, tr, va, te, hv, ht, _, _, pop = synthetic_dataset(300, N, L, seed=5)
    a = types.SimpleNamespace(max_seq_len=L, embedding_dim=D, num_attention_heads=2, drop_rate=0.1, transformer_block=2,
                              num_words_title=0, num_words_abstract=0, num_words_body=0, news_attributes=["title"],
                              bert_model_load="none", word_embedding_dim=0, eval_dtype="fp32")
    model = Model(a, N, False, None, pop).cuda().eval()
    wrap = types.SimpleNamespace(module=model)
    E = M.get_item_embeddings(wrap, np.arange(N + 1), 128, a, False, 0)
    ranks = M.eval_ranks(wrap, hv, va, E, 128, a, 0).cpu()
    ref = []
    with torch.no_grad():
        for u in range(len(va)):
            seq = va[u]
            toks = [0] * (L + 1 - len(seq)) + seq[:-1]
            lm = torch.tensor([0.0] * (L + 1 - len(seq)) + [1.0] * (len(seq) - 1)).view(1, L).cuda()
            prec = model.user_encoder(E[toks].view(1, L, D), lm, 0)[:, -1]
            score = (prec.double() @ E.double().t()).reshape(-1)
            score[hv[u].cuda()] = -float("inf")
            score = score[1:]
            order = torch.sort(score, descending=True, stable=True).indices
            pos = (order == (seq[-1] - 1)).nonzero()
            ref.append(int(pos[0, 0]) + 1 if score[seq[-1] - 1] != -float("inf") else N + 1)
    ref = torch.tensor(ref)
    mism = (ranks != ref).sum()
    assert int(mism) <= 3, (int(mism), ranks[:10], ref[:10])          # fp32-grade scores: only near-ties may differ
    hit_a, hit_b = (ranks <= 10).float().mean(), (ref <= 10).float().mean()
    assert abs(float(hit_a) - float(hit_b)) <= 0.01


def test_bce_model_matches_reference_golden():
    """bce_text/main-end2end model (N4): CUDA step vs the unmodified reference's loss and gradients"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_bce as GB
    import real_cases as RC
    from idvs.morec_b200.model_bce import Model
    c = GB.CASE
    g = torch.load(os.path.join(ROOT, "tests", "golden", "real_bce_tiny.pt"), map_location="cpu", weights_only=False)
    d = GB.bce_inputs(c)
    model = GB.build(c, Model).cuda().eval()
    loss = model(d["items"].cuda(), d["log_mask"].cuda(), 0)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-3, (float(loss), float(g["loss"]))
    grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    for k, ref in g["grads"].items():
        if "pooler" in k or RC.is_null_gradient(k):
            continue
        f = grads[k].double().reshape(-1)
        smp = f[RC.grad_sample_index(f.numel())].float()
        assert float((smp - ref["sample"]).abs().max()) <= 2e-2 * ref["absmax"] + 1e-7, k
