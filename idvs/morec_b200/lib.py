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
    P, I, F, U, L = c_void_p, c_int, c_float, c_uint64, c_int64
    sig = {
        "morec_gemm": [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, F, I, P],
        "morec_layernorm_fwd": [P, P, P, I, P, P, P, P, P, I, I, F, I, F, F, U, U, U, P],
        "morec_layernorm_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, F, F, U, U, U, P],
        "morec_attn_fwd": [P, P, P, P, P, P, I, I, I, I, I, I, I, F, F, I, F, U, U, P],
        "morec_attn_bwd": [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, F, F, I, F, U, U, P],
        "morec_inbatch_mask": [P, P, P, P, I, I, I, P],
        "morec_inbatch_ce_num_tiles": [I, I],
        "morec_inbatch_ce_fwd": [P, P, P, P, P, P, I, I, I, I, I, I, P, P, P, P, P, P, P, P],
        "morec_inbatch_ce_dlogits": [P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P, I, P],
        "morec_bert_embed_fwd": [P, P, P, P, P, P, I, I, I, P],
        "morec_bert_embed_bwd": [P, P, P, P, P, I, I, I, P],
        "morec_gather_rows": [P, P, P, I, I, I, I, I, I, P],
        "morec_scatter_add_rows": [P, P, P, I, I, I, I, I, P],
        "morec_colsum": [P, P, I, I, I, I, P],
        "morec_act_bwd": [P, P, P, L, I, I, P],
        "morec_cast_f32_to_16": [P, P, L, I, P],
        "morec_cast_f32_to_16_multi": [P, P, I, I, I, P],
        "morec_adamw_multi": [P, P, I, I, P, P, P, I, I, P],
        "morec_clock_probe": [P, P],
        "morec_mask_row_lens": [P, I, I, I, P, P],
        "morec_pack_tokens": [P, I, I, P, P, I, P, P, P],
        "morec_attn_gen_fwd": [P, P, P, P, P, P, P, I, I, I, I, I, I, I, F, I, F, U, U, P],
        "morec_attn_gen_bwd": [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, F, I, F, U, U, P],
        "morec_scale_add_rows": [P, P, P, P, I, F, P, I, I, I, I, P],
        "morec_mean_rows": [P, P, I, I, I, I, P],
        "morec_eval_hist_bits": [P, P, I, I, I, P, P],
        "morec_eval_rank": [P, P, P, P, P, I, I, I, I, P, P, P],
        "morec_bce_fwd": [P, P, P, P, I, I, I, P, P, P, P],
        "morec_bce_bwd": [P, P, P, P, P, P, P, P, I, I, I, P, P, P, P],
        "morec_bert_layers_fwd": [P, I, P],
        "morec_bert_layers_bwd": [P, I, P],
        "morec_bert_layers_bwd_ex": [P, I, I, P],
        "morec_bert_layer_fwd": [P, P],
        "morec_bert_layer_bwd": [P, P],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


_LAUNCHES = 0          # kernels launched through the C ABI since reset_counters()
_EXTRA = {"morec_inbatch_ce_fwd": 1}   # entry points that launch more than one kernel
_GEMM_TIMING = False
_GEMM_EVENTS = []      # (start_event, end_event, flops)


def _check(rc, what):
    global _LAUNCHES
    if rc != 0:
        msg = load().morec_last_error().decode("utf-8", "replace")
        raise MorecError(f"{what} failed (rc={rc}): {msg}")
    _LAUNCHES += 1 + _EXTRA.get(what, 0)


def reset_counters():
    global _LAUNCHES
    _LAUNCHES = 0
    _GEMM_EVENTS.clear()


def launch_count():
    return _LAUNCHES


def set_gemm_timing(on: bool):
    """bracket every tcgen05 GEMM launch with CUDA events on the launching stream (bench.py roofline leg)"""
    global _GEMM_TIMING
    _GEMM_TIMING = bool(on)


def collect_gemm_timing():
    """-> (total ms, total algorithmic FLOPs, launches); call after torch.cuda.synchronize()"""
    ms = sum(a.elapsed_time(b) for a, b, _ in _GEMM_EVENTS)
    fl = sum(f for _, _, f in _GEMM_EVENTS)
    n = len(_GEMM_EVENTS)
    _GEMM_EVENTS.clear()
    return ms, fl, n


_EVENT_POOL = []       # pre-created CUDA events (creation costs ~10 us; a record ~1 us)
_EVENT_NEXT = 0


def prepare_gemm_timing(n_launches: int):
    """pre-create the events for n_launches timed GEMM launches so the timed region only pays cudaEventRecord"""
    global _EVENT_NEXT
    while len(_EVENT_POOL) < 2 * n_launches:
        _EVENT_POOL.append(torch.cuda.Event(enable_timing=True))
    _EVENT_NEXT = 0


def _next_event():
    global _EVENT_NEXT
    if _EVENT_NEXT < len(_EVENT_POOL):
        e = _EVENT_POOL[_EVENT_NEXT]
    else:
        e = torch.cuda.Event(enable_timing=True)
        _EVENT_POOL.append(e)
    _EVENT_NEXT += 1
    return e


class _timed_gemm:
    def __init__(self, flops):
        self.flops = flops

    def __enter__(self):
        if _GEMM_TIMING:
            self.e0 = _next_event()
            self.e1 = _next_event()
            self.e0.record()

    def __exit__(self, *a):
        if _GEMM_TIMING:
            self.e1.record()
            _GEMM_EVENTS.append((self.e0, self.e1, self.flops))


def _ptr(t):
    """raw device address (argtypes convert the int); None -> NULL"""
    if t is None:
        return None
    if not t.is_cuda:
        raise MorecError("morec_b200 kernels take device tensors only (no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


# fp32-storage GEMM precision: True -> error-compensated 3xTF32 ("fp32" parity mode, C-ABI dtype 2),
# False -> one TF32 pass ("tf32" mode, dtype 0).  A plain module global (the autograd engine runs backward on its
# own thread), set by the ops Functions from their saved meta through `fp32_mode`.
_FP32_X3 = True


class fp32_mode:
    def __init__(self, x3: bool):
        self.x3 = bool(x3)

    def __enter__(self):
        global _FP32_X3
        self.prev = _FP32_X3
        _FP32_X3 = self.x3

    def __exit__(self, *a):
        global _FP32_X3
        _FP32_X3 = self.prev


def gemm_dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 2 if _FP32_X3 else 0
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float16:
        return 3
    raise MorecError(f"unsupported dtype {t.dtype}")


# ------------------------------------------------------------------------------------------------
# small host<->device transfers of the per-step index plans: pinned staging ring (a pageable cudaMemcpyAsync
# synchronises the stream; ~0.3 ms each, measured) and one combined device->host fetch per step
# ------------------------------------------------------------------------------------------------
_PIN_RING = []
_PIN_EVENTS = []       # per slot: event recorded behind the copy that last read it
_PIN_NEXT = 0
_PIN_SLOTS = 64
_PIN_BYTES = 1 << 20


def h2d(arr, device):
    """numpy array -> device tensor through a ring of pinned staging buffers (async, no stream sync).  A slot is
    reused 64 uploads later and is guarded by a CUDA event recorded behind its copy: a step without any host wait
    lets the host run several steps ahead of the GPU, so the reuse must wait until that copy has really executed
    (it never blocks unless the host is > 64 uploads ahead)."""
    import numpy as np
    global _PIN_NEXT
    arr = np.ascontiguousarray(arr)
    nbytes = arr.nbytes
    if nbytes == 0:
        return torch.empty(arr.shape, dtype=torch.from_numpy(arr).dtype, device=device)
    if nbytes > _PIN_BYTES:
        return torch.from_numpy(arr).pin_memory().to(device, non_blocking=True)
    if not _PIN_RING:
        for _ in range(_PIN_SLOTS):
            _PIN_RING.append(torch.empty(_PIN_BYTES, dtype=torch.uint8).pin_memory())
            _PIN_EVENTS.append(None)
    i = _PIN_NEXT
    slot = _PIN_RING[i]
    _PIN_NEXT = (_PIN_NEXT + 1) % _PIN_SLOTS
    if _PIN_EVENTS[i] is not None:
        _PIN_EVENTS[i].synchronize()
    src = torch.from_numpy(arr)
    staged = slot[:nbytes].view(src.dtype).view(arr.shape)
    staged.copy_(src)
    out = staged.to(device, non_blocking=True)
    if out.is_cuda:
        ev = _PIN_EVENTS[i] if _PIN_EVENTS[i] is not None else torch.cuda.Event()
        ev.record(torch.cuda.current_stream(out.device))
        _PIN_EVENTS[i] = ev
    return out


_D2H_BUF = {}


def d2h_begin(tensors):
    """enqueue the device->host copies of several small tensors (pinned buffers) and record an event; the caller can
    keep issuing independent work and collects the arrays with d2h_end (ONE host wait for all of them)"""
    outs = []
    for i, t in enumerate(tensors):
        t = t.contiguous()
        key = (i, t.dtype, t.numel())
        buf = _D2H_BUF.get(key)
        if buf is None:
            buf = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
            _D2H_BUF[key] = buf
        buf.copy_(t.reshape(-1), non_blocking=True)
        outs.append((buf, t.shape))
    ev = torch.cuda.Event()
    ev.record()
    return ev, outs


def d2h_end(handle):
    ev, outs = handle
    ev.synchronize()
    return [b.numpy().reshape(sh).copy() for b, sh in outs]


def d2h_many(tensors):
    """several small device tensors -> numpy arrays with ONE stream synchronisation"""
    return d2h_end(d2h_begin(tensors))


# enum mirrors
EPI_LINEAR, EPI_GELU, EPI_GELU_NOSAVE, EPI_RELU, EPI_MUL_GELU_GRAD, EPI_MUL_RELU_GRAD, EPI_GELU_DGELU, EPI_MUL_AUX = range(8)
DT_F32, DT_BF16, DT_F16 = 0, 1, 3


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float16:
        return DT_F16
    raise MorecError(f"unsupported dtype {t.dtype}")


def gemm(A, B, C, *, C2=None, bias=None, aux=None, M, N, K, lda, ldb, ldc, ldaux=0, a_mn=False, b_mn=False,
         epilogue=EPI_LINEAR, alpha=1.0, accumulate=False):
    """Raw morec_gemm.  A/B dtype decides the math kind; C dtype decides the output element type."""
    lib = load()
    dt = gemm_dtype_code(A)
    assert gemm_dtype_code(B) == dt
    with _timed_gemm(2.0 * M * N * K):
        rc = lib.morec_gemm(_ptr(A), _ptr(B), _ptr(C), _ptr(C2), _ptr(bias), _ptr(aux), M, N, K,
                            lda, ldb, ldc, ldaux, int(a_mn), int(b_mn),
                            dt, dtype_code(C), epilogue, alpha, int(accumulate), _stream())
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


def linear_dgrad(dy, w, *, epilogue=EPI_LINEAR, aux=None, out=None, out_dtype=None, accumulate=False):
    """dx[M,K] = (dy[M,N] @ w[N,K]) (* act'(aux));  accumulate=True: out (fp32) += dy @ w via TMA reduce-add."""
    M, N = dy.shape
    K = w.shape[1]
    if out is None:
        assert not accumulate
        out = torch.empty(M, K, device=dy.device, dtype=out_dtype or dy.dtype)
    gemm(dy, w, out, aux=aux, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldc=out.stride(0),
         ldaux=(aux.stride(0) if aux is not None else 0), a_mn=False, b_mn=True, epilogue=epilogue,
         accumulate=accumulate)
    return out


def linear_wgrad(dy, x, dw):
    """dw[N,K] += dy[M,N]^T @ x[M,K]   (fp32 dw, split-K + TMA reduce-add; dw must be initialised)."""
    M, N = dy.shape
    K = x.shape[1]
    assert dw.dtype == torch.float32 and dw.shape == (N, K)
    gemm(dy, x, dw, M=N, N=K, K=M, lda=dy.stride(0), ldb=x.stride(0), ldc=dw.stride(0), a_mn=True, b_mn=True,
         accumulate=True)
    return dw


# ------------------------------------------------------------------------------------------------
# thin wrappers over the remaining entry points (device tensors in, pre-allocated outputs)
# ------------------------------------------------------------------------------------------------
def _u64(x):
    return int(x) & 0xFFFFFFFFFFFFFFFF


def layernorm_fwd(x, gamma, beta, eps, *, residual=None, pos=None, pos_period=0, p_pre=0.0, p_post=0.0, seed=0,
                  off_pre=0, off_post=0, out=None, y_pre=None):
    """returns (y, y_pre_or_None, rstd)"""
    M, H = x.shape
    y = out if out is not None else torch.empty_like(x)
    if p_post > 0 and y_pre is None:
        y_pre = torch.empty_like(x)
    if not (p_post > 0):
        y_pre = None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    rc = load().morec_layernorm_fwd(_ptr(x), _ptr(residual), _ptr(pos), pos_period, _ptr(gamma), _ptr(beta),
                                    _ptr(y), _ptr(y_pre), _ptr(rstd), M, H, eps,
                                    dtype_code(x), p_pre, p_post, _u64(seed), _u64(off_pre),
                                    _u64(off_post), _stream())
    _check(rc, "morec_layernorm_fwd")
    return y, y_pre, rstd


def layernorm_bwd(dy, y, gamma, beta, rstd, *, dy2=None, dgamma, dbeta, dbias=None, dpos=None, pos_period=0,
                  p_pre=0.0, p_post=0.0, seed=0, off_pre=0, off_post=0):
    """returns (dz, dx_branch) ; dx_branch is dz itself when p_pre == 0"""
    M, H = dy.shape
    dz = torch.empty_like(dy)
    dxb = torch.empty_like(dy) if p_pre > 0 else None
    rc = load().morec_layernorm_bwd(_ptr(dy), _ptr(dy2), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(rstd), _ptr(dz),
                                    _ptr(dxb), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(dpos), pos_period,
                                    M, H, dtype_code(dy), p_pre, p_post,
                                    _u64(seed), _u64(off_pre), _u64(off_post), _stream())
    _check(rc, "morec_layernorm_bwd")
    return dz, (dxb if dxb is not None else dz)


def attn_fwd(q, k, v, o, *, cu_seqlens=None, key_mask=None, causal=False, n_seq, seqlen, n_heads, head_dim, scale,
             masked_add=-1e9, dropout_p=0.0, seed=0, offset=0):
    assert q.stride(0) == k.stride(0) == v.stride(0)
    rc = load().morec_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(cu_seqlens), _ptr(key_mask), int(causal),
                               n_seq, seqlen, n_heads, head_dim, q.stride(0),
                               o.stride(0), scale, masked_add, gemm_dtype_code(q), dropout_p, _u64(seed),
                               _u64(offset), _stream())
    _check(rc, "morec_attn_fwd")
    return o


def attn_bwd(q, k, v, do, dq, dk, dv, *, cu_seqlens=None, key_mask=None, causal=False, n_seq, seqlen, n_heads,
             head_dim, scale, masked_add=-1e9, dropout_p=0.0, seed=0, offset=0):
    assert q.stride(0) == k.stride(0) == v.stride(0) == dq.stride(0) == dk.stride(0) == dv.stride(0)
    rc = load().morec_attn_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(do), _ptr(dq), _ptr(dk), _ptr(dv), _ptr(cu_seqlens),
                               _ptr(key_mask), int(causal), n_seq, seqlen, n_heads,
                               head_dim, q.stride(0), do.stride(0), scale, masked_add,
                               gemm_dtype_code(q), dropout_p, _u64(seed), _u64(offset), _stream())
    _check(rc, "morec_attn_bwd")


def attn_gen_fwd(q, k, v, o, *, cu_seqlens=None, bias=None, mask=None, n_seq, seqlen, n_heads, head_dim, scale,
                 dropout_p=0.0, seed=0, offset=0):
    assert q.stride(0) == k.stride(0) == v.stride(0)
    rc = load().morec_attn_gen_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(cu_seqlens), _ptr(bias), _ptr(mask),
                                   (mask.shape[0] if mask is not None else 0), n_seq, seqlen, n_heads, head_dim,
                                   q.stride(0), o.stride(0), scale, gemm_dtype_code(q), dropout_p, _u64(seed), _u64(offset),
                                   _stream())
    _check(rc, "morec_attn_gen_fwd")
    return o


def attn_gen_bwd(q, k, v, do, dq, dk, dv, *, dbias=None, cu_seqlens=None, bias=None, mask=None, n_seq, seqlen, n_heads,
                 head_dim, scale, dropout_p=0.0, seed=0, offset=0):
    assert q.stride(0) == k.stride(0) == v.stride(0) == dq.stride(0) == dk.stride(0) == dv.stride(0)
    rc = load().morec_attn_gen_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(do), _ptr(dq), _ptr(dk), _ptr(dv), _ptr(dbias),
                                   _ptr(cu_seqlens), _ptr(bias), _ptr(mask), (mask.shape[0] if mask is not None else 0),
                                   n_seq, seqlen, n_heads, head_dim, q.stride(0), do.stride(0), scale, gemm_dtype_code(q),
                                   dropout_p, _u64(seed), _u64(offset), _stream())
    _check(rc, "morec_attn_gen_bwd")


def scale_add_rows(y, *, x=None, idx=None, group_scale=None, rows_per_group=1, alpha=1.0, out=None):
    """out[r] = (x[r] or 0) + alpha * group_scale[r // rows_per_group] * y[idx[r] or r]"""
    n = idx.numel() if idx is not None else (x.shape[0] if x is not None else y.shape[0])
    H = y.shape[1]
    if out is None:
        out = torch.empty(n, H, device=y.device, dtype=y.dtype)
    rc = load().morec_scale_add_rows(_ptr(x), _ptr(y), _ptr(idx), _ptr(group_scale), rows_per_group, alpha, _ptr(out), n, H,
                                     y.stride(0), dtype_code(y), _stream())
    _check(rc, "morec_scale_add_rows")
    return out


def mean_rows(x, n_groups, rows_per_group):
    H = x.shape[1]
    out = torch.empty(n_groups, H, device=x.device, dtype=x.dtype)
    rc = load().morec_mean_rows(_ptr(x), _ptr(out), n_groups, rows_per_group, H, dtype_code(x), _stream())
    _check(rc, "morec_mean_rows")
    return out


def inbatch_mask(row_ids, col_ids, B, L):
    C = col_ids.numel()
    Wc = (C + 31) // 32
    member = torch.empty(B, Wc, device=row_ids.device, dtype=torch.int32)
    pad = torch.empty(Wc, device=row_ids.device, dtype=torch.int32)
    rc = load().morec_inbatch_mask(_ptr(row_ids), _ptr(col_ids), _ptr(member), _ptr(pad), B, L, C,
                                   _stream())
    _check(rc, "morec_inbatch_mask")
    return member, pad


def inbatch_ce_fwd(P, E, member, pad, log_pop, log_mask, B, L, col_offset=0, want_loss=True):
    R, D = P.shape
    C = E.shape[0]
    NT = load().morec_inbatch_ce_num_tiles(C, gemm_dtype_code(P))
    dev = P.device
    part = torch.empty(2, R, NT, device=dev, dtype=torch.float32)
    tgt = torch.empty(R, device=dev, dtype=torch.float32)
    row_lse = torch.empty(R, device=dev, dtype=torch.float32)
    row_loss = torch.empty(R, device=dev, dtype=torch.float32)
    sum_cnt = torch.empty(2, device=dev, dtype=torch.float32)
    loss = torch.empty((), device=dev, dtype=torch.float32) if want_loss else None
    rc = load().morec_inbatch_ce_fwd(_ptr(P), _ptr(E), _ptr(member), _ptr(pad), _ptr(log_pop), _ptr(log_mask), B,
                                     L, D, C, col_offset, gemm_dtype_code(P),
                                     _ptr(part[0]), _ptr(part[1]), _ptr(tgt), _ptr(row_lse), _ptr(row_loss),
                                     _ptr(sum_cnt), _ptr(loss), _stream())
    _check(rc, "morec_inbatch_ce_fwd")
    return loss, row_lse, row_loss, sum_cnt, tgt


def inbatch_ce_dlogits(P, E, member, pad, log_pop, log_mask, row_lse, grad_out, n_valid, B, L, col_offset=0):
    R, D = P.shape
    C = E.shape[0]
    align = 4 if P.dtype == torch.float32 else 8
    ld = (C + align - 1) // align * align
    dS = torch.empty(R, ld, device=P.device, dtype=P.dtype)
    rc = load().morec_inbatch_ce_dlogits(_ptr(P), _ptr(E), _ptr(member), _ptr(pad), _ptr(log_pop), _ptr(log_mask),
                                         _ptr(row_lse), _ptr(grad_out), _ptr(n_valid), B, L, D,
                                         C, col_offset, gemm_dtype_code(P), _ptr(dS), ld, _stream())
    _check(rc, "morec_inbatch_ce_dlogits")
    return dS[:, :C]


def bert_embed_fwd(ids, pos, word, posemb, type0, out):
    n_tok, H = out.shape
    rc = load().morec_bert_embed_fwd(_ptr(ids), _ptr(pos), _ptr(word), _ptr(posemb), _ptr(type0), _ptr(out), n_tok,
                                     H, dtype_code(out), _stream())
    _check(rc, "morec_bert_embed_fwd")
    return out


def bert_embed_bwd(dz, ids, pos, dword, dposemb):
    n_tok, H = dz.shape
    rc = load().morec_bert_embed_bwd(_ptr(dz), _ptr(ids), _ptr(pos), _ptr(dword), _ptr(dposemb), n_tok, H,
                                     dtype_code(dz), _stream())
    _check(rc, "morec_bert_embed_bwd")


def gather_rows(src, idx, out=None, out_dtype=None):
    n = idx.numel()
    H = src.shape[1]
    if out is None:
        out = torch.empty(n, H, device=src.device, dtype=out_dtype or src.dtype)
    rc = load().morec_gather_rows(_ptr(src), _ptr(idx), _ptr(out), n, H, src.stride(0),
                                  out.stride(0), dtype_code(src), dtype_code(out), _stream())
    _check(rc, "morec_gather_rows")
    return out


def scatter_add_rows(src, idx, dst):
    n, H = src.shape
    assert dst.dtype == torch.float32
    rc = load().morec_scatter_add_rows(_ptr(src), _ptr(idx), _ptr(dst), n, H, src.stride(0),
                                       dst.stride(0), dtype_code(src), _stream())
    _check(rc, "morec_scatter_add_rows")
    return dst


# Bumped by every FusedAdamW.step(): the fused optimizer writes parameters through raw pointers, which torch's
# per-tensor version counters do not see.  Weight-shadow caches (ops.ShadowSet) compare against it.
PARAM_EPOCH = 0


class CastPlan:
    """persistent 16-bit shadows (bf16 or fp16) of a fixed set of fp32 tensors + the device table that casts all of
    them in one launch"""

    def __init__(self, srcs, dtype=torch.bfloat16):
        import numpy as np
        lib = load()
        lib.morec_cast_chunk_elems.restype = c_int
        chunk = lib.morec_cast_chunk_elems()
        dev = srcs[0].device
        self.key = tuple(t.data_ptr() for t in srcs)
        self.dtype = dtype
        self.dst = [torch.empty(t.shape, device=dev, dtype=dtype) for t in srcs]
        n = len(srcs)
        rec = np.zeros(n, dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("n", "<i8")], align=True))
        assert rec.itemsize == 24
        for i, (a, b) in enumerate(zip(srcs, self.dst)):
            assert a.dtype == torch.float32 and a.is_contiguous() and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
            rec[i] = (a.data_ptr(), b.data_ptr(), a.numel())
        counts = (rec["n"] + chunk - 1) // chunk
        start = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(counts, out=start[1:])
        self.n, self.n_chunks = n, int(start[-1])
        self.table = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
        self.start = torch.from_numpy(start).to(dev)

    def matches(self, srcs, dtype=torch.bfloat16):
        return len(srcs) == self.n and dtype == self.dtype and all(t.data_ptr() == k for t, k in zip(srcs, self.key))

    def run(self):
        rc = load().morec_cast_f32_to_16_multi(_ptr(self.table), _ptr(self.start), self.n, self.n_chunks,
                                               dtype_code(self.dst[0]), _stream())
        _check(rc, "morec_cast_f32_to_16_multi")
        return self.dst


def mask_row_lens(text, T):
    """text [n, >=2T] int64 (ids || attention mask), row stride arbitrary -> int32 [n] count of real tokens per row"""
    assert text.dtype == torch.int64 and text.stride(1) == 1
    n = text.shape[0]
    lens = torch.empty(n, device=text.device, dtype=torch.int32)
    rc = load().morec_mask_row_lens(_ptr(text), text.stride(0), T, n, _ptr(lens), _stream())
    _check(rc, "morec_mask_row_lens")
    return lens


def pack_tokens(text, T, enc_rows, cu, n_tok):
    """kept word pieces of the rows enc_rows (int32) of text -> (tok_ids int64 [n_tok], tok_pos int32 [n_tok])"""
    assert text.dtype == torch.int64 and text.stride(1) == 1
    tok_ids = torch.empty(n_tok, device=text.device, dtype=torch.int64)
    tok_pos = torch.empty(n_tok, device=text.device, dtype=torch.int32)
    rc = load().morec_pack_tokens(_ptr(text), text.stride(0), T, _ptr(enc_rows), _ptr(cu), enc_rows.numel(),
                                  _ptr(tok_ids), _ptr(tok_pos), _stream())
    _check(rc, "morec_pack_tokens")
    return tok_ids, tok_pos


def colsum(x, out):
    M, N = x.shape
    rc = load().morec_colsum(_ptr(x), _ptr(out), M, N, x.stride(0), dtype_code(x), _stream())
    _check(rc, "morec_colsum")
    return out


def act_bwd(dy, aux, mode, out=None):
    """out = dy * act'(aux); mode 0: erf-GELU (aux = pre-activation), 1: ReLU (aux = activation output)"""
    out = out if out is not None else torch.empty_like(dy)
    assert dy.is_contiguous() and aux.is_contiguous()
    rc = load().morec_act_bwd(_ptr(dy), _ptr(aux), _ptr(out), dy.numel(), mode, dtype_code(dy),
                              _stream())
    _check(rc, "morec_act_bwd")
    return out


def cast_f32_to_16(src, dst):
    """fp32 -> bf16 / fp16 (dst.dtype)"""
    rc = load().morec_cast_f32_to_16(_ptr(src), _ptr(dst), src.numel(), dtype_code(dst), _stream())
    _check(rc, "morec_cast_f32_to_16")
    return dst


class AdamTensor(ctypes.Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("p16", c_void_p),
                ("n", c_int), ("lr", c_float), ("wd", c_float), ("beta1", c_float), ("beta2", c_float),
                ("eps", c_float)]


AdamChunk = AdamTensor   # layout check in tests


def adamw_chunk_elems():
    h = load()
    h.morec_adamw_chunk_elems.restype = c_int
    return h.morec_adamw_chunk_elems()


_EXTRA["morec_adamw_multi"] = 1            # step-counter kernel + update kernel (+1 more with check_finite)


def adamw_multi(table_dev, chunk_start_dev, n_tensors, n_chunks, step_dev, grad_scale=None, found_inf=None,
                check_finite=False, p16_dtype=DT_BF16):
    """step_dev: device float32 scalar (advanced on the device when the update is applied)"""
    global _LAUNCHES
    rc = load().morec_adamw_multi(_ptr(table_dev), _ptr(chunk_start_dev), n_tensors, n_chunks, _ptr(step_dev),
                                  _ptr(grad_scale), _ptr(found_inf), int(check_finite), int(p16_dtype), _stream())
    _check(rc, "morec_adamw_multi")
    if check_finite and found_inf is not None:
        _LAUNCHES += 1


class BertLayerFwd(ctypes.Structure):
    _fields_ = ([(n, c_int) for n in ("n_tok", "n_seq", "H", "I", "n_heads", "max_len", "dtype", "_pad")]
                + [(n, c_float) for n in ("eps", "p_hidden", "p_attn", "_padf")]
                + [(n, c_uint64) for n in ("seed", "off_attn", "off_ln1", "off_ln2")]
                + [(n, c_void_p) for n in ("cu_seqlens", "wqkv", "bqkv", "w_ao", "b_ao", "g1", "b1", "w_i", "b_i", "w_o",
                                           "b_o", "g2", "b2", "x", "qkv", "ctx", "tmp_h", "x1", "rstd1", "pre", "act",
                                           "x2", "rstd2")])


class BertLayerBwd(ctypes.Structure):
    _fields_ = ([("fwd", BertLayerFwd)]
                + [(n, c_void_p) for n in ("dy", "dy2", "dz1", "dxq", "dz2", "dbr", "dx1b", "dctx", "dpre", "dqkv",
                                           "dwqkv", "dbqkv", "dw_ao", "db_ao", "dg1", "db1", "dw_i", "db_i", "dw_o",
                                           "db_o", "dg2", "db2")])


_N_LAYER_FWD_LAUNCHES, _N_LAYER_BWD_LAUNCHES = 7, 17


def bert_layer_fwd(args: BertLayerFwd):
    global _LAUNCHES
    rc = load().morec_bert_layer_fwd(ctypes.addressof(args), _stream())
    _check(rc, "morec_bert_layer_fwd")
    _LAUNCHES += _N_LAYER_FWD_LAUNCHES - 1


def bert_layer_bwd(args: BertLayerBwd):
    global _LAUNCHES
    rc = load().morec_bert_layer_bwd(ctypes.addressof(args), _stream())
    _check(rc, "morec_bert_layer_bwd")
    _LAUNCHES += _N_LAYER_BWD_LAUNCHES - 1


def eval_hist_bits(hist_ptr, hist_items, n_cols):
    """CSR history (hist_ptr int32 [U+1], hist_items int64) -> bit matrix [U, ceil(n_cols/32)] (column 0 always set)"""
    U = hist_ptr.numel() - 1
    bits = torch.empty(U, (n_cols + 31) // 32, device=hist_ptr.device, dtype=torch.int32)
    rc = load().morec_eval_hist_bits(_ptr(hist_ptr), _ptr(hist_items) if hist_items.numel() else None, U,
                                     hist_items.numel(), n_cols, _ptr(bits), _stream())
    _check(rc, "morec_eval_hist_bits")
    return bits


_EXTRA["morec_eval_hist_bits"] = 1     # memset + kernel
_EXTRA["morec_eval_rank"] = 1


def eval_rank(P, E, hist_bits, tgt, *, want_seen=False):
    """rank of item tgt[u] among the catalogue rows of E for every user row of P (see include/morec_b200.h).
    Returns (rank int32 [U], tgt_score [U], tgt_seen [U] | None)."""
    U, D = P.shape
    n_cols = E.shape[0]
    dev = P.device
    tgt = tgt.to(torch.int32).contiguous()
    # target scores through the SAME GEMM path: gathered target rows, padded to a tile-friendly column count so the
    # kernel variant (tile width, CTA pairing) matches the big pass
    n_pad = max(512, (U + 255) // 256 * 256)
    Et = torch.zeros(n_pad, D, device=dev, dtype=E.dtype)
    gather_rows(E, tgt, out=Et[:U])
    T = torch.empty(U, n_pad, device=dev, dtype=torch.float32)
    gemm(P, Et, T, M=U, N=n_pad, K=D, lda=P.stride(0), ldb=D, ldc=n_pad)
    tgt_score = T.diagonal()[:U].contiguous() if U <= n_pad else None
    count = torch.empty(U, device=dev, dtype=torch.int32)
    seen = torch.empty(U, device=dev, dtype=torch.float32) if want_seen else None
    with _timed_gemm(2.0 * U * n_cols * D):
        rc = load().morec_eval_rank(_ptr(P), _ptr(E), _ptr(hist_bits), _ptr(tgt_score), _ptr(tgt), U, n_cols, D,
                                    gemm_dtype_code(P), _ptr(count), _ptr(seen), _stream())
    _check(rc, "morec_eval_rank")
    return count + 1, tgt_score, seen


def bce_fwd(P, Epos, Eneg, log_mask):
    R, D = P.shape
    pos = torch.empty(R, device=P.device, dtype=torch.float32)
    neg = torch.empty(R, device=P.device, dtype=torch.float32)
    sum_cnt = torch.empty(2, device=P.device, dtype=torch.float32)
    rc = load().morec_bce_fwd(_ptr(P), _ptr(Epos), _ptr(Eneg), _ptr(log_mask), R, D, dtype_code(P), _ptr(pos), _ptr(neg),
                              _ptr(sum_cnt), _stream())
    _check(rc, "morec_bce_fwd")
    return pos, neg, sum_cnt


_EXTRA["morec_bce_fwd"] = 1


def bce_bwd(P, Epos, Eneg, log_mask, pos, neg, grad_out, sum_cnt):
    R, D = P.shape
    dP, dEp, dEn = torch.empty_like(P), torch.empty_like(Epos), torch.empty_like(Eneg)
    rc = load().morec_bce_bwd(_ptr(P), _ptr(Epos), _ptr(Eneg), _ptr(log_mask), _ptr(pos), _ptr(neg), _ptr(grad_out),
                              _ptr(sum_cnt), R, D, dtype_code(P), _ptr(dP), _ptr(dEp), _ptr(dEn), _stream())
    _check(rc, "morec_bce_bwd")
    return dP, dEp, dEn


# numpy mirrors of MorecBertLayerFwd / MorecBertLayerBwd (include/morec_b200.h): whole towers are described by ONE
# structured array filled column-wise (vectorised over layers) and handed to C++ in one call
import numpy as _np

_FWD_PTRS = ("cu_seqlens", "wqkv", "bqkv", "w_ao", "b_ao", "g1", "b1", "w_i", "b_i", "w_o", "b_o", "g2", "b2", "x", "qkv", "ctx",
             "tmp_h", "x1", "rstd1", "pre", "act", "x2", "rstd2")
_BWD_PTRS = ("dy", "dy2", "dz1", "dxq", "dz2", "dbr", "dx1b", "dctx", "dpre", "dqkv", "dwqkv", "dbqkv", "dw_ao", "db_ao", "dg1",
             "db1", "dw_i", "db_i", "dw_o", "db_o", "dg2", "db2")
LAYER_FWD_DT = _np.dtype([(n, "<i4") for n in ("n_tok", "n_seq", "H", "I", "n_heads", "max_len", "dtype", "_pad")]
                         + [(n, "<f4") for n in ("eps", "p_hidden", "p_attn", "_padf")]
                         + [(n, "<u8") for n in ("seed", "off_attn", "off_ln1", "off_ln2")]
                         + [(n, "<u8") for n in _FWD_PTRS], align=True)
LAYER_BWD_DT = _np.dtype([("fwd", LAYER_FWD_DT)] + [(n, "<u8") for n in _BWD_PTRS], align=True)
assert LAYER_FWD_DT.itemsize == ctypes.sizeof(BertLayerFwd) and LAYER_BWD_DT.itemsize == ctypes.sizeof(BertLayerBwd)


def bert_layers_fwd(rec):
    """rec: numpy array of LAYER_FWD_DT records in execution order"""
    global _LAUNCHES
    rc = load().morec_bert_layers_fwd(rec.ctypes.data, rec.shape[0], _stream())
    _check(rc, "morec_bert_layers_fwd")
    _LAUNCHES += _N_LAYER_FWD_LAUNCHES * rec.shape[0] - 1


def bert_layers_bwd(rec, join=True):
    """rec: numpy array of LAYER_BWD_DT records in execution order (last layer first), or None (join only).
    join=False leaves the last record's weight-gradient work in flight on the side stream (see the header)"""
    global _LAUNCHES
    n = 0 if rec is None else rec.shape[0]
    rc = load().morec_bert_layers_bwd_ex(rec.ctypes.data if n else None, n, int(join), _stream())
    _check(rc, "morec_bert_layers_bwd_ex")
    if n:
        _LAUNCHES += _N_LAYER_BWD_LAUNCHES * n - 1


def clock_probe(out):
    """out: int64[2] device tensor <- (SM cycles, ns) of a ~20 us spin on the current stream"""
    rc = load().morec_clock_probe(_ptr(out), _stream())
    _check(rc, "morec_clock_probe")
