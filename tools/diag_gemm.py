"""GPU diagnostic for the tcgen05 GEMM: every (dtype, A-major, B-major, epilogue) combo vs torch fp64 matmul.
Each group runs in its own subprocess with a timeout so a deadlocked kernel cannot hang the whole call.

    python tools/diag_gemm.py            # driver: spawns groups
    python tools/diag_gemm.py GROUP      # one group in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = os.environ.get("DIAG_GROUPS", "f32_kk,f32_kmn,f32_mnmn,bf16_kk,bf16_kmn,bf16_mnmn,epi,perf").split(",")


def ref_mm(A, B, a_mn, b_mn):
    import torch
    a = A.double().t() if a_mn else A.double()
    b = B.double().t() if b_mn else B.double()
    return a @ b.t()


def run_case(dt, a_mn, b_mn, M, N, K, accumulate=False, out_bf16=False):
    import torch
    from idvs.morec_b200 import lib
    torch.manual_seed(M * 7 + N * 3 + K)
    tdt = torch.float32 if dt == "f32" else torch.bfloat16
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(tdt)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(tdt)
    odt = torch.bfloat16 if out_bf16 else torch.float32
    C = torch.full((M, N), 7.0 if accumulate else float("nan"), device="cuda", dtype=odt)
    lib.gemm(A, B, C, M=M, N=N, K=K, lda=A.stride(0), ldb=B.stride(0), ldc=C.stride(0), a_mn=a_mn, b_mn=b_mn,
             accumulate=accumulate)
    torch.cuda.synchronize()
    ref = ref_mm(A, B, a_mn, b_mn) + (7.0 if accumulate else 0.0)
    err = (C.double() - ref).abs()
    scale = ref.abs().max().item()
    bad = torch.isnan(C).sum().item()
    print(f"  {dt} a_mn={int(a_mn)} b_mn={int(b_mn)} M={M} N={N} K={K} acc={int(accumulate)} obf={int(out_bf16)}: "
          f"max_err={err.max().item():.4e} mean_err={err.mean().item():.4e} ref_max={scale:.3f} nan={bad}", flush=True)
    return err.max().item() / max(scale, 1e-9)


def group(name):
    import torch
    from idvs.morec_b200 import lib
    print(f"[{name}] sms={lib.load().morec_device_sms()}", flush=True)
    shapes = [(128, 128, 32), (128, 256, 64), (256, 256, 256), (300, 200, 100), (1600, 512, 512), (4096, 768, 768),
              (1000, 3072, 768), (777, 64, 2048)]
    if name in ("f32_kk", "f32_kmn", "f32_mnmn", "bf16_kk", "bf16_kmn", "bf16_mnmn"):
        dt, lay = name.split("_")
        a_mn, b_mn = {"kk": (False, False), "kmn": (False, True), "mnmn": (True, True)}[lay]
        worst = 0.0
        for (M, N, K) in shapes:
            if dt == "bf16" and (K % 8 or M % 8 or N % 8):
                continue
            if (a_mn and M % 4) or (b_mn and N % 4) or K % 4:
                continue
            worst = max(worst, run_case(dt, a_mn, b_mn, M, N, K))
        if lay == "mnmn":
            worst = max(worst, run_case(dt, True, True, 768, 768, 49920 // 4, accumulate=True))
            worst = max(worst, run_case(dt, True, True, 512, 2048, 1600, accumulate=True))
        if dt == "bf16":
            worst = max(worst, run_case(dt, a_mn, b_mn, 1024, 768, 512, out_bf16=True))
        print(f"[{name}] worst_rel={worst:.3e}", flush=True)
    elif name == "epi":
        import torch.nn.functional as F
        for dt in (torch.float32, torch.bfloat16):
            M, N, K = 900, 512, 256
            x = torch.randn(M, K, device="cuda").to(dt)
            w = (torch.randn(N, K, device="cuda") * 0.1).to(dt)
            b = torch.randn(N, device="cuda")
            pre = torch.empty(M, N, device="cuda", dtype=dt)
            y = lib.linear_fwd(x, w, b, epilogue=lib.EPI_GELU, pre=pre)
            torch.cuda.synchronize()
            rp = x.double() @ w.double().t() + b.double()
            print(f"  gelu {dt}: pre_err={(pre.double()-rp).abs().max().item():.3e} "
                  f"act_err={(y.double()-F.gelu(rp)).abs().max().item():.3e}", flush=True)
            y2 = lib.linear_fwd(x, w, b, epilogue=lib.EPI_RELU)
            print(f"  relu {dt}: err={(y2.double()-F.relu(rp)).abs().max().item():.3e}", flush=True)
            dy = torch.randn(M, N, device="cuda").to(dt)
            aux = torch.randn(M, K, device="cuda").to(dt)
            dx = lib.linear_dgrad(dy, w, epilogue=lib.EPI_MUL_GELU_GRAD, aux=aux)
            a64 = aux.double().requires_grad_(True)
            F.gelu(a64).sum().backward()
            rdx = (dy.double() @ w.double()) * a64.grad
            print(f"  dgrad*gelu' {dt}: err={(dx.double()-rdx).abs().max().item():.3e} ref_max={rdx.abs().max().item():.2f}", flush=True)
            dx2 = lib.linear_dgrad(dy, w, epilogue=lib.EPI_MUL_RELU_GRAD, aux=aux)
            rdx2 = (dy.double() @ w.double()) * (aux.double() > 0)
            print(f"  dgrad*relu' {dt}: err={(dx2.double()-rdx2).abs().max().item():.3e}", flush=True)
            dw = torch.zeros(N, K, device="cuda")
            lib.linear_wgrad(dy, x, dw)
            rdw = dy.double().t() @ x.double()
            print(f"  wgrad {dt}: err={(dw.double()-rdw).abs().max().item():.3e} ref_max={rdw.abs().max().item():.2f}", flush=True)
    elif name == "perf":
        for dt in (torch.float32, torch.bfloat16):
            for (M, N, K) in [(49920, 768, 768), (49920, 3072, 768), (49920, 768, 3072), (8192, 8192, 8192)]:
                x = torch.randn(M, K, device="cuda").to(dt)
                w = torch.randn(N, K, device="cuda").to(dt)
                y = torch.empty(M, N, device="cuda", dtype=dt)
                for _ in range(3):
                    lib.linear_fwd(x, w, out=y)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                it = 10
                for _ in range(it):
                    lib.linear_fwd(x, w, out=y)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / it
                tf = 2.0 * M * N * K / ms / 1e9
                torch.backends.cuda.matmul.allow_tf32 = True
                for _ in range(3):
                    torch.matmul(x, w.t(), out=y)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(it):
                    torch.matmul(x, w.t(), out=y)
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / it
                print(f"  perf {dt} M={M} N={N} K={K}: {ms:.3f} ms {tf:.1f} TFLOP/s | cublas {ms2:.3f} ms "
                      f"{2.0*M*N*K/ms2/1e9:.1f} TFLOP/s", flush=True)
            # wgrad perf
            M, N, K = 49920, 768, 768
            dy = torch.randn(M, N, device="cuda").to(dt)
            x = torch.randn(M, K, device="cuda").to(dt)
            dw = torch.zeros(N, K, device="cuda")
            for _ in range(3):
                lib.linear_wgrad(dy, x, dw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.linear_wgrad(dy, x, dw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"  perf wgrad {dt}: {ms:.3f} ms {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
            dx = torch.empty(M, K, device="cuda", dtype=dt)
            for _ in range(3):
                lib.linear_dgrad(dy, w[:768, :768].contiguous() if w.shape != (768, 768) else w, out=dx)
            torch.cuda.synchronize()
    print(f"[{name}] done", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        group(sys.argv[1])
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        for gname in GROUPS:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), gname], timeout=240,
                                   capture_output=True, text=True)
                out = r.stdout + ("\nSTDERR:\n" + r.stderr[-3000:] if r.returncode else "")
                print(out)
                print(f"[{gname}] rc={r.returncode} {time.time()-t0:.1f}s", flush=True)
            except subprocess.TimeoutExpired as e:
                print(f"[{gname}] TIMEOUT (hang?) partial:\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else e.stdout}", flush=True)
