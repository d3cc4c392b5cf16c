#!/usr/bin/env python
"""bench.py -- training sequences/sec of the MoRec in-batch step (SASRec + BERT-base, L=25) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|tf32|bf16] [--impl morec|reference]

One "step" = one pass of the hot path over one synthetic MIND-shape batch: H2D (e2e only) -> Model.forward (item
encoder over the batch's items -> SASRec -> in-batch debiased CE) -> backward -> fused AdamW, i.e. the loop body of
inbatch_sasrec_e2e_text/run.py:231-247.  Workload = BASELINE.json configs[2] (the configuration the metric is
quoted on): BERT-base, B=64 users per GPU, L=25, T=30 word pieces, D=512, 2 SASRec blocks x 2 heads, N=50k items.

Prints ONE JSON line (rank 0).  `value` = whole-job sequences/s with the W+K batches already resident in HBM;
`e2e` = the same through the public Model API with host buffers (pinned H2D of each step's batch and a D2H read of
the loss inside the timed region); `roofline` = the dominant kernel (the tcgen05 GEMM) timed with CUDA events on the
launching stream inside the timed region; `cpu_baseline` = the oracle port on the host cores on a bounded sample.
`--impl reference` times the UNMODIFIED reference Model (baseline/_ref, mirrored by __graft_entry__.build()) on the
host cores at the full batch; without that copy, the oracle port on a bounded sample (`cpu_baseline.kind`).
Extra keys: `roofline_scoring` (K8, the fused scoring + CE kernel, at the cfg-3 and the 8-GPU all-gathered shapes,
against both rooflines), `flops_dense_vs_executed` (tensor work actually executed vs the reference's dense count),
`modes` (seq/s of every precision mode), `roofline.traffic` (ncu DRAM bytes of the dominant kernel, per launch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# batches differ in packed token count, so activation sizes change every step: expandable segments keep the caching
# allocator from falling back to cudaMalloc/cudaFree (device-synchronising) while it is still growing
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

BERT_BASE = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12)
CFG = dict(B=64, L=25, T=30, D=512, heads=2, blocks=2, N=50000, drop=0.1,
           lr=1e-4, fine_tune_lr=5e-5, l2=0.01, fine_tune_l2=0.01)      # train_bert_base.py:22-28
# BASELINE.json configs[1]: SASRec + BERT-tiny, titles of 128 word pieces (run.py:55-57: H=128, 2 layers, 2 heads, I=512)
BERT_TINY = dict(BERT_BASE, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512)
CFG_TINY = dict(CFG, T=128)
# BASELINE.json configs[3]: SASRec + Swin-T, HM-shape synthetic 3x224x224, B=32, L=10 (V/parameters.py:38), D=512
SWIN_T = dict(image_size=224, patch_size=4, num_channels=3, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
              window_size=7, mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
              drop_path_rate=0.1, hidden_act="gelu", layer_norm_eps=1e-5)
CFG_VISION = dict(B=32, L=10, T=0, D=512, heads=2, blocks=2, N=50000, drop=0.1,
                  lr=1e-4, fine_tune_lr=1e-4, l2=0.1, fine_tune_l2=0.1)    # train_swin_tiny.py
# BASELINE.json configs[4]: SASRec + Swin-B, B=16 per GPU
SWIN_B = dict(SWIN_T, embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32])
CFG_VISION_B = dict(CFG_VISION, B=16)


# DRAM traffic of the dominant kernel from one `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum
# of ONE launch); a tensor-bound kernel: the figure only shows that nothing is re-read (algorithmic = A + B + C once)
TRAFFIC = {"bytes": 110831360, "algorithmic_bytes": 23207424 + 2 * 73955328,
           "launch": "gemm2_kernel<f16, GELU+GELU' epilogue> 12037 x 3072 x 768 (FFN1 forward: A 18.5 MB + B 4.7 MB read, two "
                     "12037 x 3072 fp16 outputs written; 87.5 MB of the 147.9 MB written had left L2 when the kernel ended)",
           "other": "plain 12037 x 768 x 768: 19.72 MB read = A + B exactly, 0 written back inside the kernel",
           "source": "profiles/r02_ncu_gemm_epilogues.txt"}


def make_args(cfg):
    a = types.SimpleNamespace()
    a.max_seq_len = cfg["L"]; a.embedding_dim = cfg["D"]; a.num_attention_heads = cfg["heads"]
    a.drop_rate = cfg["drop"]; a.transformer_block = cfg["blocks"]; a.num_words_title = cfg["T"]
    a.num_words_abstract = 50; a.num_words_body = 50; a.news_attributes = ["title"]
    tiny = cfg.get("bert") is BERT_TINY
    a.bert_model_load = "bert_tiny" if tiny else "bert_base_uncased"
    a.word_embedding_dim = 128 if tiny else 768
    return a


# ------------------------------------------------------------------------------------------------ clocks
class ClockMonitor:
    """SM clock and throttle reasons of the timed region.

    NVML / nvidia-smi queries issued WHILE the step loop runs were measured to stall CUDA launches on these hosts
    (host-bound step time 35 ms -> 82 ms with a 100 ms polling thread), so the in-region clock is measured on the
    device itself: after every timed step a one-thread probe kernel (morec_clock_probe, ~20 us) reads clock64 and
    globaltimer -> SM MHz under load.  NVML is queried immediately before the first timed step is enqueued (GPU hot
    from the warm-up steps) and immediately after the last one completes: max clock + the union of the throttle /
    event reasons seen at both edges."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index, n_probes):
        import torch
        self.buf = torch.zeros(n_probes, 2, dtype=torch.int64, device=f"cuda:{index}")
        self.n = 0
        self.bits = 0
        self.smax = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
        self.nvml_sm = []

    def edge(self):
        if self.h is None:
            return
        try:
            try:
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.nvml_sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        except Exception:
            pass

    def probe(self):
        from idvs.morec_b200 import lib
        if self.n < self.buf.shape[0]:
            lib.clock_probe(self.buf[self.n])
            self.n += 1

    def result(self):
        mhz = sorted(float(c) / float(ns) * 1e3 for c, ns in self.buf[:self.n].cpu().tolist() if ns > 0)
        if not mhz:
            return None
        return {"sm_mhz": round(mhz[len(mhz) // 2], 1), "sm_mhz_min": round(mhz[0], 1), "sm_max_mhz": self.smax,
                "reasons": sorted(k for k, m in self.REASONS.items() if self.bits & m), "samples": len(mhz),
                "nvml_sm_mhz_at_edges": self.nvml_sm,
                "method": "device clock64/globaltimer probe after every timed step; NVML reasons at both edges"}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_oracle_step_fn(cfg, B_sample, seed):
    """returns (step_fn, cores): one training step of the oracle port (fwd + bwd + AdamW) on B_sample users"""
    import torch
    from transformers import BertConfig, BertModel
    from oracle import morec_oracle as O
    from idvs.morec_b200.synth import synth_batch
    torch.manual_seed(seed)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bert = BertModel(BertConfig(**cfg.get("bert", BERT_BASE)))
    sd = {("bert_encoder.text_encoders.title.bert_model." + k): v for k, v in bert.state_dict().items()}
    D, L = cfg["D"], cfg["L"]
    g = torch.Generator().manual_seed(seed)
    sd["bert_encoder.text_encoders.title.fc.weight"] = torch.randn(D, 768, generator=g) * 0.03
    sd["bert_encoder.text_encoders.title.fc.bias"] = torch.zeros(D)
    pre = "user_encoder.transformer_encoder."
    sd[pre + "position_embedding.weight"] = torch.randn(L, D, generator=g) * 0.05
    sd[pre + "layer_norm.weight"] = torch.ones(D); sd[pre + "layer_norm.bias"] = torch.zeros(D)
    for b in range(cfg["blocks"]):
        q = pre + f"transformer_blocks.{b}."
        for n in ("w_Q", "w_K", "w_V", "fc"):
            sd[q + f"multi_head_attention.{n}.weight"] = torch.randn(D, D, generator=g) * 0.04
        sd[q + "multi_head_attention.layer_norm.weight"] = torch.ones(D); sd[q + "multi_head_attention.layer_norm.bias"] = torch.zeros(D)
        sd[q + "feed_forward.w_1.weight"] = torch.randn(4 * D, D, generator=g) * 0.03; sd[q + "feed_forward.w_1.bias"] = torch.zeros(4 * D)
        sd[q + "feed_forward.w_2.weight"] = torch.randn(D, 4 * D, generator=g) * 0.03; sd[q + "feed_forward.w_2.bias"] = torch.zeros(D)
        sd[q + "feed_forward.layer_norm.weight"] = torch.ones(D); sd[q + "feed_forward.layer_norm.bias"] = torch.zeros(D)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "pooler" not in k) for k, v in sd.items()
              if v.is_floating_point()}
    bert_p = [p for k, p in params.items() if "bert_model" in k and p.requires_grad]
    rec_p = [p for k, p in params.items() if "bert_model" not in k and p.requires_grad]
    opt = torch.optim.AdamW([{"params": bert_p, "lr": cfg["fine_tune_lr"], "weight_decay": cfg["fine_tune_l2"]},
                             {"params": rec_p, "lr": cfg["lr"], "weight_decay": cfg["l2"]}])
    batch = synth_batch(B_sample, L, cfg["N"], cfg["T"], seed, modal=True)

    def step():
        opt.zero_grad()
        out = O.model_forward(params, batch["ids"], batch["items"], batch["log_mask"], batch["pop_prob"], use_modal=True,
                              n_heads_user=cfg["heads"], n_heads_bert=12)
        out.loss.backward()
        opt.step()
        return float(out.loss)

    return step, cores


def time_cpu_oracle(cfg, B_sample, steps, warmup, seed=12345):
    import torch
    step, cores = cpu_oracle_step_fn(cfg, B_sample, seed)
    # PyTorch's CPU kernels do not always scale to every core of a big host: calibrate the thread count on one step
    best = None
    for nt in sorted({min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(nt)
        t0 = time.time()
        step()
        dt = time.time() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    cores = best[1]
    torch.set_num_threads(cores)
    for _ in range(max(warmup - 1, 0)):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / max(steps, 1)
    return dict(value=B_sample / dt, unit="sequences/s", cores=cores, kind="port",
                sample=f"oracle port (plain PyTorch CPU fp32, oracle/morec_oracle.py), B={B_sample} of {cfg['B']} users "
                       f"per step (same L/T/D/encoder), {steps} timed step(s) after thread-count calibration "
                       f"(best of {os.cpu_count()} host cores: {cores} threads), {dt:.2f} s/step"), dt, steps


# ------------------------------------------------------------------------------------------------ unmodified reference arm
def ref_root():
    """directory holding the UNMODIFIED reference `model/` packages: baseline/_ref (a git-ignored copy made by
    __graft_entry__.build() in the authoring container; it travels to the GPU box with the gpurun snapshot), else
    /root/reference when this runs in the authoring container itself."""
    for r in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(r, "inbatch_sasrec_e2e_text", "model", "model.py")):
            return r
    return None


def load_reference_model_cls(root, pkg="inbatch_sasrec_e2e_text"):
    """import <root>/<pkg>/model as a uniquely named package (its own relative imports resolve inside it; nothing of
    this repo is on that path) and return the reference's Model class"""
    import importlib.util
    name = "_ref_" + pkg + "_model"
    if name in sys.modules:
        return sys.modules[name].Model
    d = os.path.join(root, pkg, "model")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod.Model


def reference_step_fn(cfg, B, seed, root):
    """one training step of the unmodified reference Model on CPU fp32 exactly as BASELINE.md §2 prescribes:
    zero_grad -> forward(local_rank='cpu') -> backward -> AdamW (two groups, run.py:150-162), same synthetic batch
    generator and same BERT-base config as the B200 arm"""
    import torch
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.synth import synth_batch
    RefModel = load_reference_model_cls(root)
    torch.manual_seed(seed)
    bert = BertModel(BertConfig(**cfg.get("bert", BERT_BASE)))
    for i, (n, p) in enumerate(bert.named_parameters()):          # run.py:73-75 (freeze_paras_before=0; pooler frozen)
        if i in (197, 198):
            p.requires_grad = False
    batch = synth_batch(B, cfg["L"], cfg["N"], cfg["T"], seed, modal=True)
    model = RefModel(make_args(cfg), cfg["N"], True, bert, batch["pop_prob"].numpy())
    model.train()
    bert_p = [p for n, p in model.named_parameters() if p.requires_grad and "bert_model" in n]
    rec_p = [p for n, p in model.named_parameters() if p.requires_grad and "bert_model" not in n]
    opt = torch.optim.AdamW([{"params": bert_p, "lr": cfg["fine_tune_lr"], "weight_decay": cfg["fine_tune_l2"]},
                             {"params": rec_p, "lr": cfg["lr"], "weight_decay": cfg["l2"]}])
    ids, items, lm = batch["ids"].reshape(-1), batch["items"], batch["log_mask"]

    def step():
        opt.zero_grad()
        loss = model(ids, items, lm, "cpu")
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_reference(cfg, steps, warmup, budget_s, seed=12345, users=0):
    """unmodified reference at the FULL batch (B = cfg['B']).  A BERT-base CPU step takes tens of seconds, so the
    number of timed steps is bounded by a time budget: min(steps, max(3, budget / step time)); `warmup` >= 1 steps
    are capped the same way.  Returns (cpu_baseline dict, seconds per step, steps actually timed, warm-ups done)."""
    import torch
    root = ref_root()
    cores = os.cpu_count() or 1
    # thread count: calibrated on a 4-user step of the same model (cheap), as PyTorch's CPU kernels do not always
    # scale to every core of a big host
    cal = reference_step_fn(cfg, 4, seed, root)
    best = None
    for nt in sorted({min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(nt)
        cal()
        t0 = time.time()
        cal()
        dt = time.time() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    del cal
    threads = best[1]
    torch.set_num_threads(threads)
    B = users or cfg["B"]
    step = reference_step_fn(cfg, B, seed, root)
    t0 = time.time()
    step()
    t_first = time.time() - t0
    n_warm = 1
    while n_warm < warmup and (n_warm + 1) * t_first < 0.25 * budget_s:
        step()
        n_warm += 1
    n_timed = int(min(max(steps, 1), max(3 if steps >= 3 else steps, (budget_s - n_warm * t_first) // max(t_first, 1e-3))))
    times = []
    for _ in range(n_timed):
        t0 = time.time()
        step()
        times.append(time.time() - t0)
    times.sort()
    dt = times[len(times) // 2]
    return dict(value=B / dt, unit="sequences/s", cores=threads, kind="reference",
                sample=f"UNMODIFIED reference Model ({os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}/"
                       f"inbatch_sasrec_e2e_text/model), CPU fp32, {'full batch' if B == cfg['B'] else 'SAMPLE of the batch:'} B={B} (same L/T/D/BERT-base config, same "
                       f"synthetic generator, AdamW 2 groups), {n_warm} warm-up + {n_timed} timed step(s) (median "
                       f"{dt:.2f} s/step; min {times[0]:.2f}, max {times[-1]:.2f}); {threads} threads (calibrated on a "
                       f"4-user step) of {cores} host cores"), dt, n_timed, n_warm


def _synth_images(ids, seed, dev=None):
    """HM-shape synthetic item images: a deterministic randn image per item id (post-normalisation range,
    V/data_utils/dataset.py:71), zero image for pad slots; generated per distinct id so duplicates are identical."""
    import torch
    flat = ids.reshape(-1)
    uniq, inv = torch.unique(flat, return_inverse=True)
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn(uniq.numel(), 3, 224, 224, generator=g)
    imgs[uniq == 0] = 0
    return imgs[inv]


def setup_training_vision(cfg, mode, n_batches, rank=0, world=1, local_rank=0, parallel="global", use_d2h_plan=False):
    import torch
    from transformers import SwinConfig, SwinForImageClassification
    from idvs.morec_b200.model_vision import Model
    from idvs.morec_b200.optim import FusedAdamW
    from idvs.morec_b200.synth import synth_batch
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(12345)
    net = SwinForImageClassification(SwinConfig(**cfg.get("swin", SWIN_T)))
    net.classifier = torch.nn.Linear(net.classifier.in_features, cfg["D"])          # V/run.py:49-54
    torch.nn.init.xavier_normal_(net.classifier.weight.data)
    torch.nn.init.constant_(net.classifier.bias.data, 0)
    batches = [synth_batch(cfg["B"], cfg["L"], cfg["N"], 0, 777 + 1000 * rank + i, modal=False, mind_shape=False)
               for i in range(n_batches)]
    a = types.SimpleNamespace(max_seq_len=cfg["L"], embedding_dim=cfg["D"], num_attention_heads=cfg["heads"],
                              drop_rate=cfg["drop"], transformer_block=cfg["blocks"],
                              CV_model_load="swin_base" if cfg.get("swin") is SWIN_B else "swin_tiny")
    from idvs.morec_b200.synth import pop_from_batches
    model = Model(a, cfg["N"], True, net, pop_from_batches(batches).numpy()).to(dev)
    model.set_compute_dtype(mode)
    model.item_dedup = "always"      # north star: "forward over the batch's unique items"
    model.parallel_mode = "local" if world > 1 else parallel     # (global mode exchanges token rows; images stay local)
    model.train()
    if world > 1:
        from idvs.morec_b200.parallel import wrap_ddp
        model_run = wrap_ddp(model, local_rank)
    else:
        model_run = model
    net_params = [p for n, p in model.named_parameters() if p.requires_grad and "image_net" in n and "classifier" not in n]
    rec_params = [p for n, p in model.named_parameters() if p.requires_grad and not ("image_net" in n and "classifier" not in n)]
    opt = FusedAdamW([{"params": net_params, "lr": cfg["fine_tune_lr"], "weight_decay": cfg["fine_tune_l2"]},
                      {"params": rec_params, "lr": cfg["lr"], "weight_decay": cfg["l2"]}])
    host = [(b["ids"].pin_memory(), _synth_images(b["ids"], 99 + i).pin_memory(), b["log_mask"].pin_memory())
            for i, b in enumerate(batches)]
    resident = [(a_.to(dev), b_.to(dev), c_.to(dev)) for (a_, b_, c_) in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 14)      # fp16 storage: dynamic loss scaling (V/run.py:186, 213-215)

    def step(ids, items, lm):
        opt.zero_grad(set_to_none=True)
        loss = model_run(ids.view(-1), items, lm, local_rank)
        if model.compute_dtype == "fp16":
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        else:
            loss.backward()
            opt.step()
        return loss

    return step, host, resident, h2d_bytes


def setup_training(cfg, mode, n_batches, rank=0, world=1, local_rank=0, parallel="global", use_d2h_plan=False):
    """model + optimizer + synthetic batches exactly as the reference loop builds them (run.py:127-162);
    returns (step_fn, pinned host batches, device-resident batches, H2D bytes per step)"""
    import torch
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.model import Model
    from idvs.morec_b200.optim import FusedAdamW
    from idvs.morec_b200.synth import synth_batch
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(12345)
    bert = BertModel(BertConfig(**cfg.get("bert", BERT_BASE)))
    pooler = (37, 38) if cfg.get("bert") is BERT_TINY else (197, 198)
    for i, (n, p) in enumerate(bert.named_parameters()):          # run.py:55-75 (freeze_paras_before=0; pooler frozen)
        if i in pooler:
            p.requires_grad = False
    from idvs.morec_b200.synth import pop_from_batches
    batches = [synth_batch(cfg["B"], cfg["L"], cfg["N"], cfg["T"], 12345 + 1000 * rank + i, modal=True)
               for i in range(n_batches)]
    pop = pop_from_batches(batches).numpy()       # every in-batch id of every batch has p > 0
    from idvs.morec_b200.synth import synth_catalogue
    catalogue = synth_catalogue(cfg["N"], cfg["T"], 4242)      # ONE catalogue for every batch and every rank
    for b in batches:
        b["items"] = catalogue[b["ids"].reshape(-1)]
    model = Model(make_args(cfg), cfg["N"], True, bert, pop).to(dev)
    model.set_compute_dtype(mode)
    model.item_dedup = "always"      # north star: "forward over the batch's unique items" (at every N)
    model.parallel_mode = parallel
    model.train()
    if world > 1:
        from idvs.morec_b200.parallel import wrap_ddp
        model_run = wrap_ddp(model, local_rank)
    else:
        model_run = model
    bert_params = [p for n, p in model.named_parameters() if p.requires_grad and "bert_model" in n]
    rec_params = [p for n, p in model.named_parameters() if p.requires_grad and "bert_model" not in n]
    opt = FusedAdamW([{"params": bert_params, "lr": cfg["fine_tune_lr"], "weight_decay": cfg["fine_tune_l2"]},
                      {"params": rec_params, "lr": cfg["lr"], "weight_decay": cfg["l2"]}])
    model.attach_optimizer(opt)      # 16-bit weight copies are written by the AdamW kernel (no per-step cast pass)
    model.set_item_content(catalogue)             # static per-item token counts
    host = [(b["ids"].pin_memory(), b["items"].pin_memory(), b["log_mask"].pin_memory()) for b in batches]
    resident = [(a.to(dev), b.to(dev), c.to(dev)) for (a, b, c) in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    host_ids = {}                         # device batch -> its ids as a host array (the data loader had them there)
    for (hi, _, _), (ri, _, _) in zip(host, resident):
        host_ids[ri.data_ptr()] = hi.numpy()
    # fp16 storage needs dynamic loss scaling exactly like the reference's autocast loop (run.py:210, 245-247)
    scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 14)

    def step(ids, items, lm, ids_host=None):
        opt.zero_grad(set_to_none=True)
        if ids_host is None:
            ids_host = host_ids.get(ids.data_ptr())
        loss = model_run(ids.view(-1), items.view(-1, items.size(-1)), lm, local_rank,
                         host_ids=None if use_d2h_plan else ids_host)
        if model.compute_dtype == "fp16":
            scaler.scale(loss).backward()
            scaler.step(opt)         # unscale + overflow skip fused into the AdamW kernel (no host wait)
            scaler.update()
        else:
            loss.backward()
            opt.step()
        return loss

    step.model = model
    return step, host, resident, h2d_bytes



# ------------------------------------------------------------------------------------------------ K8 roofline
def scoring_roofline(cfg, peaks, dev):
    """K8 = fused scoring + debias + masks + CE (model/model.py:45-67).  Forward = mask pre-pass + scoring GEMM with the
    CE-partials epilogue + combine; backward = dlogits GEMM + dP and dE GEMMs.  Timed alone (CUDA events, operands
    L2-warm as in the step, where E has just been written) at the single-GPU shape and at the 8-GPU all-gathered
    shape, in the arithmetic the step uses for it (3xTF32 on fp32 operands) and in fp16.  Algorithmic bytes per
    SURVEY.md 8(d): s*(R+C)*D + 12*C + 12*R forward (the [R,C] logits are never counted), the same plus the dP / dE
    outputs backward; FLOPs 2*R*C*D forward, 6*R*C*D backward."""
    import torch
    from idvs.morec_b200 import lib
    from idvs.morec_b200.synth import synth_batch
    B, L, D = cfg["B"], cfg["L"], cfg["D"]
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tens = float(peaks.get("bf16_tflops", 1604.0))
    out = []
    for G in (1, 8):
        bs = [synth_batch(B, L, cfg["N"], 0, seed=100 + g, modal=False, n_users_pop=2000) for g in range(G)]
        ids_all = torch.cat([b["ids"].reshape(-1) for b in bs]).to(dev)
        ids_loc = bs[0]["ids"].to(dev)
        lm = bs[0]["log_mask"].reshape(-1).to(dev)
        R, C = B * L, ids_all.numel()
        logp = torch.rand(C, device=dev).log()
        g1 = torch.ones(1, device=dev)
        for dt, x3, name, es in ((torch.float32, True, "fp32 (3xTF32, as in the step)", 4), (torch.float16, False, "fp16", 2)):
            P = (torch.randn(R, D, device=dev) * 0.3).to(dt)
            E = (torch.randn(C, D, device=dev) * 0.3).to(dt)

            def fwd():
                member, pad = lib.inbatch_mask(ids_loc, ids_all, B, L)
                return (member, pad) + tuple(lib.inbatch_ce_fwd(P, E, member, pad, logp, lm, B, L))

            def bwd(st):
                member, pad, loss, row_lse, row_loss, sum_cnt = st[:6]
                dS = lib.inbatch_ce_dlogits(P, E, member, pad, logp, lm, row_lse, g1, sum_cnt[1:2].contiguous(), B, L)
                dP = torch.empty(R, D, device=dev, dtype=dt)
                dE = torch.empty(C, D, device=dev, dtype=dt)
                lib.gemm(dS, E, dP, M=R, N=D, K=C, lda=dS.stride(0), ldb=E.stride(0), ldc=D, a_mn=False, b_mn=True)
                lib.gemm(dS, P, dE, M=C, N=D, K=R, lda=dS.stride(0), ldb=P.stride(0), ldc=D, a_mn=True, b_mn=True)

            with lib.fp32_mode(x3):
                st = fwd()
                bwd(st)
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                n = 20
                e0.record()
                for _ in range(n):
                    st = fwd()
                e1.record()
                for _ in range(n):
                    bwd(st)
                e2.record()
                torch.cuda.synchronize()
            us_f, us_b = e0.elapsed_time(e1) / n * 1e3, e1.elapsed_time(e2) / n * 1e3
            by_f = es * (R + C) * D + 12 * C + 12 * R
            by_b = by_f + es * (R + C) * D
            fl_f, fl_b = 2.0 * R * C * D, 6.0 * R * C * D
            tpk = tens * (1.0 / 6.0 if x3 else 1.0)       # 3 passes at the half-rate tf32 kind
            out.append({"shape": f"R={R} C={C} D={D} (G={G})", "arith": name,
                        "fwd_us": round(us_f, 1), "bwd_us": round(us_b, 1), "launches": {"fwd": 3, "bwd": 3},
                        "alg_bytes_fwd": by_f, "fwd_gbs": round(by_f / us_f / 1e3, 1),
                        "fwd_frac_hbm": round(by_f / us_f / 1e3 / hbm, 4),
                        "fwd_tflops": round(fl_f / us_f / 1e6, 1), "fwd_frac_tensor": round(fl_f / us_f / 1e6 / tpk, 3),
                        "bwd_tflops": round(fl_b / us_b / 1e6, 1), "bwd_frac_tensor": round(fl_b / us_b / 1e6 / tpk, 3),
                        "hbm_floor_us": round(by_f / hbm / 1e3, 2), "tensor_floor_us": round(fl_f / tpk / 1e6, 2)})
    return {"kernel": "K8: morec::inbatch_mask_kernel + gemm_kernel<.,CeFwdEpi> + inbatch_ce_combine_kernel (fwd); "
                      "gemm_kernel<.,CeBwdEpi> + 2 x gemm (bwd)",
            "peaks": {"hbm_gbs": hbm, "tensor_tflops_f16": tens, "tensor_tflops_3xtf32": round(tens / 6.0, 1)},
            "shapes": out,
            "note": "bound is the tensor roofline, not HBM: the algorithmic traffic is 3-30 MB for 2.7-22 GFLOP, so 60 % of "
                    "the HBM peak would need > 3 PFLOP/s; *_frac_tensor is against the rate the arithmetic allows "
                    "(3xTF32 = 1/6 of the f16 peak); at the G=1 shape the three launches sit on the launch-latency floor"}

# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="morec", choices=["morec", "reference"])
    ap.add_argument("--mode", default=os.environ.get("MOREC_MODE", "fp16"), choices=["fp32", "tf32", "bf16", "fp16"],
                    help="fp16 (default) = the arithmetic of the reference's own loop (autocast + GradScaler, run.py:242-247)")
    ap.add_argument("--no-modes", action="store_true", help="skip the short per-mode throughput block")
    ap.add_argument("--parallel", default="global", choices=["global", "local"],
                    help="multi-GPU semantics for N > 1 (idvs/morec_b200/parallel.py)")
    ap.add_argument("--workload", default="text", choices=["text", "text_tiny", "vision", "vision_b"],
                    help="text = SASRec+BERT-base (headline, configs[2]); text_tiny = BERT-tiny, T=128 (configs[1]); "
                         "vision = SASRec+Swin-T (configs[3]); vision_b = Swin-B, B=16 (configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scoring", action="store_true", help="skip the K8 (scoring + CE kernel) roofline block")
    ap.add_argument("--d2h-plan", action="store_true", help="plan each step from a device->host copy of the ids and token "
                    "counts (one host wait per step) instead of from the ids the host already has")
    ap.add_argument("--cpu-sample-users", type=int, default=4)
    ap.add_argument("--cpu-port", action="store_true", help="reference arm / cpu_baseline: time the oracle port even "
                    "when a copy of the unmodified reference is available")
    ap.add_argument("--ref-users", type=int, default=0, help="reference arm: users per step (0 = the full batch B)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="wall-clock budget of the reference arm's steps (a BERT-base CPU step takes tens of seconds)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    vision = args.workload.startswith("vision")
    cfg = {"text": dict(CFG, bert=BERT_BASE), "text_tiny": dict(CFG_TINY, bert=BERT_TINY),
           "vision": dict(CFG_VISION, swin=SWIN_T), "vision_b": dict(CFG_VISION_B, swin=SWIN_B)}[args.workload]
    workload = (f"MoRec SASRec+Swin-{'B' if args.workload == 'vision_b' else 'T'} end2end in-batch debiased CE, B={cfg['B']}/GPU, "
                f"L={cfg['L']}, 3x224x224 images, D={cfg['D']}, HM-shape synthetic "
                f"(BASELINE.json configs[{4 if args.workload == 'vision_b' else 3}])") if vision else \
               (f"MoRec SASRec+BERT-{'tiny' if args.workload == 'text_tiny' else 'base'} end2end in-batch debiased CE, "
                f"B={cfg['B']}/GPU, L={cfg['L']}, T={cfg['T']}, D={cfg['D']}, N={cfg['N']} items, MIND-shape synthetic "
                f"(BASELINE.json configs[{1 if args.workload == 'text_tiny' else 2}])")

    if args.impl == "reference":
        if rank != 0:
            return
        W = max(args.warmup, 1)
        if ref_root() is not None and not args.cpu_port:
            # the UNMODIFIED reference Model at the full batch; timed steps bounded by a wall-clock budget
            base, dt, n_timed, n_warm = time_reference(cfg, args.steps, W, args.ref_budget_s, users=args.ref_users)
        else:
            # no copy of the reference on this box: the oracle port (pinned to the reference's goldens) on a sample
            base, dt, n_timed = time_cpu_oracle(cfg, args.cpu_sample_users, args.steps, W)
            n_warm = W
        line = {"impl": "reference", "metric": "training sequences/sec", "value": base["value"], "unit": "sequences/s",
                "n_gpus": args.gpus, "steps": n_timed, "warmup": n_warm, "steps_requested": args.steps,
                "warmup_requested": args.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from idvs.morec_b200 import lib

    lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    W, K = max(args.warmup, 3), args.steps
    step, host, resident, h2d_bytes = (setup_training_vision if vision else setup_training)(
        cfg, args.mode, W + K, rank, world, local_rank, args.parallel, args.d2h_plan)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`): K steps, batches already in HBM
    # warm-up: W steps, the first one on the batch with the most real tokens so the allocator reaches its
    # steady-state footprint before anything is timed
    if vision:  # activation sizes follow the number of DISTINCT images of a step
        import numpy as np
        big = max(range(len(host)), key=lambda i: int(np.unique(host[i][0].numpy()).size))
    else:       # packed token count of a step = tokens of the batch's DISTINCT items
        import numpy as np
        lens = step.model._item_lens
        if world > 1 and args.parallel == "global":
            # a rank encodes its SHARE of the global batch's distinct items: regenerate every rank's ids (the synthetic
            # generator is seeded per rank and batch) and take the batch whose largest share has the most tokens --
            # the same batch index on every rank, the collectives of a step pair up by position
            from idvs.morec_b200.parallel import plan_global_batch
            from idvs.morec_b200.synth import synth_batch

            def share_tokens(i):
                ids_all = np.concatenate([synth_batch(cfg["B"], cfg["L"], cfg["N"], cfg["T"], 12345 + 1000 * r + i,
                                                      modal=False)["ids"].numpy().reshape(-1) for r in range(world)])
                worst, n_unique = 0, 0
                for r in range(world):
                    pl = plan_global_batch(ids_all, world, r)
                    worst = max(worst, int(lens[ids_all[pl.my_first_slots]].sum()))
                    n_unique = pl.n_unique
                return worst, n_unique
            shares = [share_tokens(i) for i in range(len(host))]
            big = max(range(len(host)), key=lambda i: shares[i][0])
            big2 = max(range(len(host)), key=lambda i: shares[i][1])     # most distinct items: largest all-gather buffers
        else:
            big = max(range(len(host)), key=lambda i: int(lens[np.unique(host[i][0].numpy())].sum()))
    step(*resident[big])
    if not vision and world > 1 and args.parallel == "global":
        step(*resident[big2])
    for i in range(W - 1):
        step(*resident[i])
    sync()
    mon = ClockMonitor(local_rank, K) if rank == 0 else None
    lib.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    if mon:
        mon.edge()
    ms0 = torch.cuda.memory_stats(dev)
    t_wall0 = time.time()
    step_t = []
    e0.record()
    for i in range(K):
        step(*resident[W + i])
        if mon:
            mon.probe()
        step_t.append(time.time())
    e1.record()
    ms1 = torch.cuda.memory_stats(dev)
    alloc_delta = {k: int(ms1.get(k, 0) - ms0.get(k, 0)) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    step_host_ms = [round(1e3 * (b - a), 1) for a, b in zip([t_wall0] + step_t[:-1], step_t)]
    t_issue = time.time() - t_wall0          # host time to enqueue K steps (includes the in-step size syncs)
    sync()
    if mon:
        mon.edge()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count()
    clocks = mon.result() if mon else None
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms) / K
    value = cfg["B"] * world / (ms_step / 1e3)

    # ---------------- roofline leg: the same K steps again with every tcgen05 GEMM launch bracketed by CUDA events
    # on the launching stream (kept out of the `value` loop: ~350 event records per step perturb the host side)
    lib.set_gemm_timing(True)
    step(*resident[W - 1])
    sync()
    _, _, per_step = lib.collect_gemm_timing()
    lib.prepare_gemm_timing(per_step * K + 64)
    for i in range(K):
        step(*resident[W + i])
    sync()
    lib.set_gemm_timing(False)
    gemm_ms, gemm_flops, gemm_n = lib.collect_gemm_timing()

    # ---------------- end to end through the public API: pinned H2D of each batch + D2H of the loss, every step
    for i in range(2):
        if vision:
            step(*[t.to(dev, non_blocking=True) for t in host[i]])
        else:
            step(*[t.to(dev, non_blocking=True) for t in host[i]], host[i][0].numpy())
    sync()

    def e2e_loop(lagged):
        """K steps from pinned host batches; every step's loss is read back to the host inside the timed region --
        immediately (the host waits for the GPU before it may prepare the next step: the reference's run.py:249), or
        one step late (lagged: the loss of step i is fetched after step i+1 has been enqueued, through a pinned
        buffer and an event, so the read never idles the GPU; what the drop-in run.py's logging does)"""
        losses, pending = [], None
        e0.record()
        for i in range(K):
            ids, items, lm = [t.to(dev, non_blocking=True) for t in host[W + i]]
            loss = step(ids, items, lm, host[W + i][0].numpy()) if not vision else step(ids, items, lm)
            if not lagged:
                losses.append(float(loss.detach()))          # D2H read + host wait
                continue
            buf = loss_pinned[i & 1]
            buf.copy_(loss.detach().reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                losses.append(float(pending[0][0]))
            pending = (buf, ev)
        if pending is not None:
            pending[1].synchronize()
            losses.append(float(pending[0][0]))
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return cfg["B"] * world / (float(t) / K / 1e3), losses

    loss_pinned = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_sync_value, e2e_losses = e2e_loop(lagged=False)
    e2e_value, e2e_losses_lagged = e2e_loop(lagged=True)
    e2e_losses = e2e_losses + e2e_losses_lagged
    import math
    assert all(math.isfinite(v) for v in e2e_losses), f"non-finite training loss in the timed steps: {e2e_losses}"

    # ---------------- per-mode throughput (few steps each): makes the cost of the parity mode driver-visible
    modes = None
    if not vision and not args.no_modes and hasattr(step, "model"):
        modes = {}
        for m in ("fp32", "tf32", "bf16", "fp16"):
            if m == args.mode:
                modes[m] = {"seq_per_s": value, "ms_per_step": ms_step, "steps": K}
                continue
            step.model.set_compute_dtype(m)
            for i in range(2):
                step(*resident[i])
            sync()
            e0.record()
            for i in range(3):
                step(*resident[W + (i % max(K, 1))])
            e1.record()
            sync()
            tms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            modes[m] = {"seq_per_s": cfg["B"] * world / (float(tms) / 3 / 1e3), "ms_per_step": float(tms) / 3, "steps": 3}
        step.model.set_compute_dtype(args.mode)
        modes["note"] = ("fp32 = 3xTF32 parity mode (loss/logits <= 1e-3 vs the reference CPU fp32 path, tests/test_parity_gpu.py); "
                         "fp16 = the reference's own autocast arithmetic (run.py:242) with GradScaler; measured deviations of "
                         "every mode: profiles/r02_parity_modes.json")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "morec::gemm2_kernel / gemm_kernel (tcgen05 CTA-pair and single-CTA GEMMs: every fwd / dgrad / wgrad / scoring launch)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": TRAFFIC["bytes"],
                "traffic_detail": TRAFFIC,
                "peak_source": peak_src, "launches_timed": gemm_n, "gemm_ms_per_step": gemm_ms / K,
                "note": {"fp32": "3xTF32 parity mode: 3 tensor-core passes per algorithmic FLOP and kind::tf32 runs at half "
                                 "the bf16 rate, so the attainable fraction of the bf16 peak is 1/6",
                         "tf32": "kind::tf32 runs at half the bf16 rate: attainable fraction of the bf16 peak is 1/2",
                         "bf16": "kind::f16 on bf16 operands", "fp16": "kind::f16 on fp16 operands"}[args.mode]}
    # tensor work executed vs the reference's dense count (it encodes all C slots x T tokens, pads and duplicates included)
    flops_dv = None
    scoring = None
    if not vision:
        import numpy as np
        lens = step.model._item_lens
        tok_exec = float(np.mean([int(lens[np.unique(host[W + i][0].numpy())].sum()) for i in range(K)]))
        bc = cfg.get("bert", BERT_BASE)
        Hh, Ii, nl = bc["hidden_size"], bc["intermediate_size"], bc["num_hidden_layers"]
        per_tok = nl * 2.0 * (4 * Hh * Hh + 2 * Hh * Ii)           # forward GEMM FLOPs of one token through the tower
        tok_dense = cfg["B"] * (cfg["L"] + 1) * cfg["T"]
        executed = gemm_flops / K
        dense = executed + 3.0 * per_tok * (tok_dense - tok_exec)
        flops_dv = {"executed_tflop_per_step": executed / 1e12, "reference_dense_tflop_per_step": dense / 1e12,
                    "executed_over_dense": executed / dense, "tokens_executed_per_step": tok_exec,
                    "tokens_dense_per_step": tok_dense,
                    "note": "pad word pieces, pad item slots and duplicate items are not encoded (bit-identical loss with "
                            "dropout off, tests/test_parity_gpu.py); the roofline counts executed FLOPs only"}
        if not args.no_scoring:
            scoring = scoring_roofline(cfg, peaks, dev)
    cpu_base = None
    if not args.no_cpu_baseline and not vision:
        if ref_root() is not None and not args.cpu_port:
            cpu_base = time_reference(cfg, 1, 1, 40.0)[0]          # 1 warm-up + 1 timed full-batch step
        else:
            cpu_base = time_cpu_oracle(cfg, args.cpu_sample_users, 1, 1)[0]
    line = {"metric": "training sequences/sec", "value": value, "unit": "sequences/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32 (3xTF32 tensor-core emulation)", "tf32": "tf32", "bf16": "bf16", "fp16": "fp16"}[args.mode],
            "data": "synthetic",
            "config": {"workload": workload, "mode": args.mode,
                       "parallelism": f"dp{world}" + (f" ({args.parallel}: " + ("item embeddings all-gathered, global negatives, "
                                      "reduce-scatter in backward; text-tower gradients all-reduced layer by layer inside its backward, the rest by DDP)" if args.parallel == "global"
                                      else "reference DDP semantics, rank-local negatives; text-tower gradients all-reduced layer by layer inside its backward)") if world > 1 else ""),
                       "l2_policy": "every step uses a different batch and streams >10 GB of activations (>> 126 MB L2)",
                       "dropout": cfg["drop"], "optimizer": "FusedAdamW (2 groups)",
                       "items_encoded": "each distinct non-pad item of the (global) batch once; pad tokens skipped"},
            "e2e": {"value": e2e_value, "unit": "sequences/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "loss_read": "every step's loss is copied to pinned host memory and read one step late (after the next "
                                 "step has been enqueued), as the drop-in run.py logs it",
                    "value_host_waits_every_step": e2e_sync_value},
            "gpu_launches": launches, "host_issue_ms_per_step": 1e3 * t_issue / K, "host_ms_each_step": step_host_ms,
            "loss_first_last_e2e": [e2e_losses[0], e2e_losses[-1]] if e2e_losses else None,
            "allocator_events_in_timed_region": alloc_delta, "clocks": clocks, "roofline": roofline,
            "roofline_scoring": scoring, "flops_dense_vs_executed": flops_dv, "modes": modes,
            "cpu_baseline": cpu_base}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
