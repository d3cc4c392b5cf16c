"""ctypes binding of the morec_b200 C ABI (include/morec_b200.h).

The product path has NO fallback: if libmorec_b200.so is missing or a call fails, a RuntimeError is raised.
(`python -m idvs.morec_b200.build` or `__graft_entry__.build()` produces the library in-tree.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmorec_b200.so")

_lib = None


class MorecError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MorecError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m idvs.morec_b200.build). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.morec_last_error.restype = c_char_p
    lib.morec_abi_version.restype = c_int
    lib.morec_device_sms.restype = c_int
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().morec_last_error().decode("utf-8", "replace")
        raise MorecError(f"{what} failed (rc={rc}): {msg}")


def _ptr(t):
    if t is None:
        return c_void_p(0)
    assert t.is_cuda, "morec_b200 kernels take device tensors only"
    return c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# enum mirrors
EPI_LINEAR, EPI_GELU, EPI_GELU_NOSAVE, EPI_RELU, EPI_MUL_GELU_GRAD, EPI_MUL_RELU_GRAD = range(6)
DT_F32, DT_BF16 = 0, 1


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.bfloat16:
        return DT_BF16
    raise MorecError(f"unsupported dtype {t.dtype}")


def gemm(A, B, C, *, C2=None, bias=None, aux=None, M, N, K, lda, ldb, ldc, ldaux=0, a_mn=False, b_mn=False,
         epilogue=EPI_LINEAR, alpha=1.0, accumulate=False):
    """Raw morec_gemm.  A/B dtype decides the math kind; C dtype decides the output element type."""
    lib = load()
    dt = dtype_code(A)
    assert dtype_code(B) == dt
    out_bf16 = 1 if C.dtype == torch.bfloat16 else 0
    rc = lib.morec_gemm(_ptr(A), _ptr(B), _ptr(C), _ptr(C2), _ptr(bias), _ptr(aux), c_int(M), c_int(N), c_int(K),
                        c_int(lda), c_int(ldb), c_int(ldc), c_int(ldaux), c_int(int(a_mn)), c_int(int(b_mn)),
                        c_int(dt), c_int(out_bf16), c_int(epilogue), c_float(alpha), c_int(int(accumulate)), _stream())
    _check(rc, "morec_gemm")


def linear_fwd(x, w, bias=None, *, epilogue=EPI_LINEAR, out=None, pre=None, out_dtype=None):
    """y[M,N] = epi(x[M,K] @ w[N,K]^T + bias)."""
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=out_dtype or x.dtype)
    gemm(x, w, out, C2=pre, bias=bias, M=M, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldc=out.stride(0),
         epilogue=epilogue)
    return out


def linear_dgrad(dy, w, *, epilogue=EPI_LINEAR, aux=None, out=None, out_dtype=None):
    """dx[M,K] = (dy[M,N] @ w[N,K]) (* act'(aux))."""
    M, N = dy.shape
    K = w.shape[1]
    if out is None:
        out = torch.empty(M, K, device=dy.device, dtype=out_dtype or dy.dtype)
    gemm(dy, w, out, aux=aux, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldc=out.stride(0),
         ldaux=(aux.stride(0) if aux is not None else 0), a_mn=False, b_mn=True, epilogue=epilogue)
    return out


def linear_wgrad(dy, x, dw):
    """dw[N,K] += dy[M,N]^T @ x[M,K]   (fp32 dw, split-K + TMA reduce-add; dw must be initialised)."""
    M, N = dy.shape
    K = x.shape[1]
    assert dw.dtype == torch.float32 and dw.shape == (N, K)
    gemm(dy, x, dw, M=N, N=K, K=M, lda=dy.stride(0), ldb=x.stride(0), ldc=dw.stride(0), a_mn=True, b_mn=True,
         accumulate=True)
    return dw
