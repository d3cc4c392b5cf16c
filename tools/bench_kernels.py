"""Micro-benchmarks of the non-GEMM kernels of a BERT-base layer at the bench shape (CUDA events, 20 launches each),
plus the GELU-epilogue GEMMs.  usage: bench_kernels.py [bf16|tf32] [n_seq]   (for `ncu --set full -k regex:...` too)"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from idvs.morec_b200 import lib

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 662
dt = torch.bfloat16 if mode == "bf16" else torch.float32
es = 2 if mode == "bf16" else 4
torch.manual_seed(0)
lens = torch.randint(6, 31, (n_seq,))
cu = torch.zeros(n_seq + 1, dtype=torch.int32)
cu[1:] = torch.cumsum(lens, 0)
n_tok = int(cu[-1])
cu = cu.cuda()
H, heads, dh, I = 768, 12, 64, 3072


def timeit(name, fn, bytes_=None, flops=None, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    extra = ""
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:.0f} GB/s"
    if flops:
        extra += f"  {flops / us / 1e6:.0f} TFLOP/s"
    print(f"{name:42s} {us:8.1f} us{extra}", flush=True)


with lib.fp32_mode(False):
    qkv = torch.randn(n_tok, 3 * H, device="cuda").to(dt)
    o = torch.empty(n_tok, H, device="cuda", dtype=dt)
    do = torch.randn(n_tok, H, device="cuda").to(dt)
    dqkv = torch.empty_like(qkv)
    kw = dict(cu_seqlens=cu, n_seq=n_seq, seqlen=30, n_heads=heads, head_dim=dh, scale=1 / math.sqrt(dh), dropout_p=0.1,
              seed=1, offset=1 << 36)
    print(f"mode={mode} n_seq={n_seq} n_tok={n_tok}")
    timeit("attn fwd (BERT, T<=30, d=64)", lambda: lib.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, **kw),
           bytes_=n_tok * 4 * H * es)
    timeit("attn bwd", lambda: lib.attn_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], do, dqkv[:, :H], dqkv[:, H:2 * H],
                                           dqkv[:, 2 * H:], **kw), bytes_=n_tok * 7 * H * es)
    # SASRec attention: B=64, L=25, 2 heads x 256
    B, L, D = 64, 25, 512
    q2 = torch.randn(B * L, 3 * D, device="cuda").to(dt)
    o2 = torch.empty(B * L, D, device="cuda", dtype=dt)
    lm = torch.ones(B, L, device="cuda")
    kw2 = dict(key_mask=lm, causal=True, n_seq=B, seqlen=L, n_heads=2, head_dim=256, scale=1 / 16.0, dropout_p=0.1, seed=1,
               offset=1 << 36)
    timeit("attn fwd (SASRec, L=25, d=256)", lambda: lib.attn_fwd(q2[:, :D], q2[:, D:2 * D], q2[:, 2 * D:], o2, **kw2))
    dq2 = torch.empty_like(q2)
    timeit("attn bwd (SASRec)", lambda: lib.attn_bwd(q2[:, :D], q2[:, D:2 * D], q2[:, 2 * D:], o2, dq2[:, :D], dq2[:, D:2 * D],
                                                    dq2[:, 2 * D:], **kw2))
    # LayerNorm
    x = torch.randn(n_tok, H, device="cuda").to(dt)
    r = torch.randn(n_tok, H, device="cuda").to(dt)
    g = torch.ones(H, device="cuda"); b = torch.zeros(H, device="cuda")
    y = torch.empty_like(x)
    timeit("ln fwd (residual, dropout 0.1)", lambda: lib.layernorm_fwd(x, g, b, 1e-12, residual=r, p_pre=0.1, seed=1, off_pre=1 << 36, out=y),
           bytes_=n_tok * 3 * H * es)
    _, _, rstd = lib.layernorm_fwd(x, g, b, 1e-12, residual=r, p_pre=0.1, seed=1, off_pre=1 << 36, out=y)
    dg, db, dbi = (torch.zeros(H, device="cuda") for _ in range(3))
    timeit("ln bwd (dy+dy2, dropout 0.1)", lambda: lib.layernorm_bwd(x, y, g, b, rstd, dy2=r, dgamma=dg, dbeta=db, dbias=dbi, p_pre=0.1,
                                                                     seed=1, off_pre=1 << 36), bytes_=n_tok * 5 * H * es)
    timeit("ln fwd (residual, no dropout)", lambda: lib.layernorm_fwd(x, g, b, 1e-12, residual=r, out=y), bytes_=n_tok * 3 * H * es)
    timeit("ln bwd (dy+dy2, no dropout)", lambda: lib.layernorm_bwd(x, y, g, b, rstd, dy2=r, dgamma=dg, dbeta=db, dbias=dbi),
           bytes_=n_tok * 4 * H * es)
    kw0 = dict(kw); kw0["dropout_p"] = 0.0
    timeit("attn fwd (no dropout)", lambda: lib.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, **kw0), bytes_=n_tok * 4 * H * es)
    timeit("attn bwd (no dropout)", lambda: lib.attn_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], do, dqkv[:, :H], dqkv[:, H:2 * H],
                                                        dqkv[:, 2 * H:], **kw0), bytes_=n_tok * 7 * H * es)
    # GEMMs of the FFN
    w_i = (torch.randn(I, H, device="cuda") * 0.02).to(dt)
    w_o = (torch.randn(H, I, device="cuda") * 0.02).to(dt)
    bi = torch.zeros(I, device="cuda")
    pre = torch.empty(n_tok, I, device="cuda", dtype=dt)
    act = torch.empty(n_tok, I, device="cuda", dtype=dt)
    fl = 2.0 * n_tok * H * I
    timeit("FFN1 fwd (GELU, 2 outputs)", lambda: lib.linear_fwd(x, w_i, bi, epilogue=lib.EPI_GELU, pre=pre, out=act), flops=fl)
    timeit("FFN1 fwd (GELU + GELU' outputs)", lambda: lib.linear_fwd(x, w_i, bi, epilogue=lib.EPI_GELU_DGELU, pre=pre, out=act), flops=fl)
    timeit("FFN1 fwd (plain bias)", lambda: lib.linear_fwd(x, w_i, bi, out=act), flops=fl)
    timeit("FFN2 fwd (plain bias)", lambda: lib.linear_fwd(act, w_o, g, out=y), flops=fl)
    timeit("dpre = (dy Wo2) * gelu'(pre)", lambda: lib.linear_dgrad(x, w_o, epilogue=lib.EPI_MUL_GELU_GRAD, aux=pre, out=act), flops=fl)
    timeit("dpre = (dy Wo2) * aux", lambda: lib.linear_dgrad(x, w_o, epilogue=lib.EPI_MUL_AUX, aux=pre, out=act), flops=fl)
    timeit("dgrad plain same shape", lambda: lib.linear_dgrad(x, w_o, out=act), flops=fl)
    dw = torch.zeros(I, H, device="cuda")
    timeit("wgrad dW_i (split-K)", lambda: lib.linear_wgrad(act, x, dw), flops=fl)
    timeit("colsum [n_tok, 3072]", lambda: lib.colsum(act, bi), bytes_=n_tok * I * es)
