"""Run one GEMM shape repeatedly (for `ncu --set full -k regex:gemm_kernel`).  usage: one_gemm.py M N K mode layout"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from idvs.morec_b200 import lib
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "tf32"
layout = sys.argv[5] if len(sys.argv) > 5 else "kk"
dt = torch.bfloat16 if mode == "bf16" else torch.float16 if mode == "fp16" else torch.float32
a_mn, b_mn = {"kk": (False, False), "kmn": (False, True), "mnmn": (True, True)}[layout]
A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(dt)
B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(dt)
C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if layout == "mnmn" else dt)
with lib.fp32_mode(mode == "fp32"):
    for _ in range(6):
        lib.gemm(A, B, C, M=M, N=N, K=K, lda=A.stride(0), ldb=B.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn,
                 accumulate=(layout == "mnmn"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.gemm(A, B, C, M=M, N=N, K=K, lda=A.stride(0), ldb=B.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn,
                 accumulate=(layout == "mnmn"))
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"{mode} {layout} M={M} N={N} K={K}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
