"""Host-side sequencing of the morec_b200 kernels as torch.autograd Functions.

torch is plumbing here (device memory, streams, the autograd graph between the coarse Functions); every FLOP of the
hot path runs in the hand-written sm_100a kernels behind the C ABI (idvs/morec_b200/lib.py).  Each Function mirrors
one piece of the reference's Model.forward (inbatch_sasrec_e2e_text/model/model.py:31-69):

    BertTowerFn   model/encoders.py:63-70  HF BertModel(ids, mask)[0][:, 0] -> fc -> GELU       (packed tokens)
    GatherRowsFn  model/model.py:37-41     nn.Embedding / slot expansion / input_embs[:, :-1]
    SasrecFn      model/encoders.py:23-28 + model/modules.py:89-96                               (post-LN blocks)
    InbatchCEFn   model/model.py:45-67     scoring + debias + masks + CE

Gradients of parameters are produced in fp32 and returned through autograd, so DistributedDataParallel hooks and
torch.cuda.amp.GradScaler (run.py:148, 245-247) keep working unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import lib

_MASKED_ADD = -1e9


@dataclass
class DropCtx:
    """dropout configuration of one Function call: probabilities + Philox seed + a base offset unique to the call"""
    p_hidden: float = 0.0
    p_attn: float = 0.0
    seed: int = 0
    base: int = 0

    def off(self, site: int) -> int:
        return self.base + (site << 36)


def _cw(p: torch.Tensor, adt: torch.dtype) -> torch.Tensor:
    """weight in compute dtype (fp32 params are used in place; the 16-bit modes make a shadow copy)"""
    p = p.detach()
    if adt == torch.float32:
        return p
    sh = torch.empty(p.shape, device=p.device, dtype=adt)
    if p.numel() % 4 == 0 and p.is_contiguous():
        lib.cast_f32_to_16(p, sh)
    else:
        sh.copy_(p)
    return sh


def _with_prec(fn):
    """run a Function.forward/backward under the fp32 GEMM precision recorded in its meta (see lib.fp32_mode)"""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        meta = args[0] if isinstance(args[0], dict) else getattr(ctx, "meta", None)
        x3 = True if meta is None else bool(meta.get("x3", True))
        with lib.fp32_mode(x3):
            return fn(ctx, *args)
    return wrapper


def _z32(p):
    return torch.zeros(p.shape, device=p.device, dtype=torch.float32)


class _Arena:
    """one zero-filled fp32 buffer per backward pass from which every parameter gradient is carved (a single memset
    instead of ~180 fill kernels; 16-byte aligned slices so they can be TMA reduce-add targets)"""

    def __init__(self, dev, shapes):
        self.total = sum((math.prod(sh) + 3) // 4 * 4 for sh in shapes)
        self.buf = torch.zeros(max(self.total, 4), device=dev, dtype=torch.float32)
        self.off = 0

    def take(self, shape):
        n = math.prod(shape)
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += (n + 3) // 4 * 4
        assert self.off <= self.buf.numel()
        return v


class _GradSync:
    """Gradient averaging of the text tower overlapped with its own backward (multi-GPU).

    Every parameter gradient of one backward pass is a slice of ONE arena, filled layer by layer (last layer first).
    As soon as a layer's slice is complete it is all-reduced (average) asynchronously on NCCL's stream while the
    layers below are still being differentiated; the tower's parameters are excluded from DistributedDataParallel
    (Model sets `_ddp_params_and_buffers_to_ignore`).  Under DDP alone all of the tower's gradients become ready at the
    same instant -- when this Function returns -- so its 340 MB all-reduce ran entirely exposed."""

    def __init__(self, arena):
        self.arena, self.start, self.handles = arena, 0, []

    @staticmethod
    def enabled(meta):
        import torch.distributed as dist
        return bool(meta.get("grad_sync")) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def flush(self):
        import torch.distributed as dist
        end = self.arena.off
        if end > self.start:
            t = self.arena.buf[self.start:end]
            if dist.get_backend() == "gloo":              # CPU tests: no AVG / no stream overlap on gloo
                dist.all_reduce(t)
                t.div_(dist.get_world_size())
            else:
                self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True))
            self.start = end

    def wait(self):
        self.flush()
        for h in self.handles:
            h.wait()                      # the current stream waits for the collective; the host does not block
        self.handles = []


class _Workspace:
    """Grow-only device buffer with a bump allocator for the token-count-dependent activations of the text tower.

    Batches differ in packed token count, so every step would otherwise ask the caching allocator for a new set of
    block sizes (observed: cudaMalloc / cuMemMap stalls of tens of ms whenever unseen sizes arrive).  The workspace
    reaches its steady-state size on the largest batch and is then reused verbatim.  A workspace is owned by one
    forward pass until its backward has run; a second concurrent owner falls back to ordinary torch.empty."""

    def __init__(self):
        self.buf = None
        self.off = 0
        self.busy = False

    def acquire(self, dev, nbytes):
        if self.busy:
            return False
        if self.buf is None or self.buf.device != dev or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes * 1.1) + (1 << 20), device=dev, dtype=torch.uint8)
        self.off = 0
        self.busy = True
        return True

    def release(self):
        self.busy = False

    def take(self, shape, dtype):
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        start = (self.off + 255) // 256 * 256
        if self.buf is None or start + n > self.buf.numel():
            return torch.empty(shape, device=self.buf.device if self.buf is not None else None, dtype=dtype)
        self.off = start + n
        return self.buf[start:start + n].view(dtype).view(shape)


_WS_FWD = {}
_WS_BWD = {}


def _ws(table, dev):
    key = (dev.type, dev.index)
    if key not in table:
        table[key] = _Workspace()
    return table[key]


class FusedParamGroup:
    """Keeps several nn.Parameters (e.g. the Q, K, V projection weights) as consecutive row blocks of ONE contiguous
    buffer by re-pointing their .data, so the fused [3H, H] projection weight exists without a per-step concat.
    The Parameter objects, their names, order and values are unchanged; if something re-allocates
    the parameters (.to(), load_state_dict copies in place so that is fine) the group is rebuilt lazily."""

    def __init__(self, params):
        self.params = list(params)
        self._numels = [p.numel() for p in self.params]
        self.buf = None

    def _valid(self):
        if self.buf is None:
            return False
        # device pointers are unique across devices (unified addressing), so matching addresses imply the same device
        ptr = self.buf.data_ptr()
        for p, n in zip(self.params, self._numels):
            if p.data_ptr() != ptr:
                return False
            ptr += 4 * n
        return True

    @torch.no_grad()
    def buffer(self):
        if not self._valid():
            p0 = self.params[0]
            cols = p0.shape[1:] if p0.dim() > 1 else ()
            rows = sum(p.shape[0] for p in self.params)
            buf = torch.empty((rows,) + tuple(cols), device=p0.device, dtype=torch.float32)
            r = 0
            for p in self.params:
                buf[r:r + p.shape[0]].copy_(p.data)
                p.data = buf[r:r + p.shape[0]]
                r += p.shape[0]
            self.buf = buf
        return self.buf


def _dgrad_acc(dy, w, into):
    """into += dy @ w"""
    if into.dtype == torch.float32:
        lib.linear_dgrad(dy, w, out=into, accumulate=True)
    else:
        into.add_(lib.linear_dgrad(dy, w))
    return into


# ==================================================================================================
# BERT text tower
# ==================================================================================================
def _layer_struct(meta, l, n_tok, n_seq, H, I, cu, w, small, acts, tmp_h):
    """MorecBertLayerFwd for layer l.  w = (wqkv, w_ao, w_i, w_o) in compute dtype, small = fp32 (bqkv, b_ao, g1, b1,
    b_i, b_o, g2, b2), acts = (x, qkv, ctx, x1, rstd1, pre, act, x2, rstd2)."""
    drop, adt = meta["drop"], meta["adt"]
    a = lib.BertLayerFwd()
    a.n_tok, a.n_seq, a.H, a.I, a.n_heads, a.max_len = n_tok, n_seq, H, I, meta["n_heads"], meta["max_len"]
    a.dtype = lib.DT_BF16 if adt == torch.bfloat16 else lib.DT_F16 if adt == torch.float16 else (2 if meta.get("x3", True) else 0)
    a.eps, a.p_hidden, a.p_attn = meta["eps"], drop.p_hidden, drop.p_attn
    a.seed = drop.seed & 0xFFFFFFFFFFFFFFFF
    a.off_attn, a.off_ln1, a.off_ln2 = drop.off(1 + 4 * l), drop.off(2 + 4 * l), drop.off(3 + 4 * l)
    a.cu_seqlens = cu.data_ptr()
    a.wqkv, a.w_ao, a.w_i, a.w_o = (t.data_ptr() for t in w)
    a.bqkv, a.b_ao, a.g1, a.b1, a.b_i, a.b_o, a.g2, a.b2 = (t.data_ptr() for t in small)
    a.x, a.qkv, a.ctx, a.x1, a.rstd1, a.pre, a.act, a.x2, a.rstd2 = (t.data_ptr() for t in acts)
    a.tmp_h = tmp_h.data_ptr()
    return a



class ShadowSet:
    """Persistent compute-dtype (bf16 / fp16) copies of a fixed list of fp32 GEMM weights.

    They are refreshed by ONE multi-tensor cast launch (lib.CastPlan) -- or not at all: when a FusedAdamW is attached
    (Model.attach_optimizer) its update kernel writes the 16-bit copy of every parameter it updates next to the fp32
    master, so the per-step cast pass over the whole tower (0.23 ms for BERT-base) disappears.  Staleness is tracked
    two ways: torch's per-tensor version counters (any in-place torch op on a parameter: load_state_dict, another
    optimizer) and lib.PARAM_EPOCH (raw-pointer updates by FusedAdamW instances that do NOT feed this set)."""

    def __init__(self):
        self.plan = None
        self.owners = None
        self.fresh_epoch = -1
        self.versions = None
        self.opt = None            # weakref to the attached FusedAdamW

    def attach(self, opt):
        import weakref
        self.opt = weakref.ref(opt) if opt is not None else None
        self.plan = None           # re-register the shadows with the new optimizer on the next get()

    def _register(self):
        opt = self.opt() if self.opt is not None else None
        if opt is None:
            return
        for dst, own in zip(self.plan.dst, self.owners):
            for (p, r0, r1) in own:
                opt.register_shadow(p, dst if (r0 == 0 and r1 == dst.shape[0]) else dst[r0:r1], self)

    def get(self, srcs, owners, adt):
        """srcs: fp32 tensors (16-byte aligned, contiguous); owners[i]: [(parameter, row0, row1)] whose storage makes
        up srcs[i] (the fused QKV weight is three parameters).  Returns the list of 16-bit tensors."""
        if self.plan is None or not self.plan.matches(srcs, adt):
            self.plan, self.owners, self.fresh_epoch = lib.CastPlan(srcs, adt), owners, -1
            self._register()
        vers = tuple(p._version for own in owners for (p, _, _) in own)
        if self.fresh_epoch != lib.PARAM_EPOCH or vers != self.versions:
            self.plan.run()
            self.fresh_epoch, self.versions = lib.PARAM_EPOCH, vers
        return self.plan.dst


def prepare_tower_weights(wqkv_bufs, flat_params, adt, cache=None):
    """compute-dtype copies of every GEMM weight of the text tower (bf16 mode: 4 per layer + fc).  They do not depend
    on the batch, so the caller issues them BEFORE it waits for the packing plan's device->host copy.  With a `cache`
    dict the shadows are persistent buffers refreshed by ONE multi-tensor cast launch (lib.CastPlan) instead of 49
    launches -- host time that an end-to-end step (loss read back every step) cannot hide.  The shadows of forward i
    are overwritten by forward i+1, i.e. a backward must run before the weights change and the next forward starts
    (every training loop does; re-casting unchanged weights writes identical values)."""
    n_layers = (len(flat_params) - 7) // 16
    srcs, owners = [], []
    for l in range(n_layers):
        ps = flat_params[5 + 16 * l: 5 + 16 * (l + 1)]
        H = ps[0].shape[0]
        srcs += [wqkv_bufs[l].detach(), ps[6].detach(), ps[10].detach(), ps[12].detach()]
        owners += [[(ps[0], 0, H), (ps[2], H, 2 * H), (ps[4], 2 * H, 3 * H)], [(ps[6], 0, ps[6].shape[0])],
                   [(ps[10], 0, ps[10].shape[0])], [(ps[12], 0, ps[12].shape[0])]]
    srcs.append(flat_params[-2].detach())
    owners.append([(flat_params[-2], 0, flat_params[-2].shape[0])])
    if adt == torch.float32:
        out = srcs
    elif cache is not None and all(t.is_contiguous() and t.data_ptr() % 16 == 0 for t in srcs):
        ss = cache.get("shadows")
        if ss is None:
            ss = cache["shadows"] = ShadowSet()
        out = ss.get(srcs, owners, adt)
    else:
        out = [_cw(t, adt) for t in srcs]
    return dict(layers=[tuple(out[4 * l: 4 * l + 4]) for l in range(n_layers)], fc=out[-1])


def _al256(n):
    return (n + 255) // 256 * 256


class TowerLayout:
    """Static description of a text tower for the C++ layer sequencer (lib.bert_layers_fwd / _bwd): device pointers of
    the weights in compute dtype and of the fp32 biases / LayerNorm parameters (one uint64 row per layer), and the
    layout of the fp32 gradient arena (fc first, then the layers LAST to FIRST -- the order the backward completes
    them, so the finished prefix can be all-reduced while the rest is still being computed -- then the embeddings).
    Built once and cached by Text_Encoder.prepare(); rebuilt when a weight buffer moves."""

    def __init__(self, flat, cw, bqkv):
        n_layers = (len(flat) - 7) // 16
        self.n_layers = n_layers
        self.key = TowerLayout.make_key(cw, bqkv)
        w = np.zeros((n_layers, 4), dtype=np.uint64)
        sm = np.zeros((n_layers, 8), dtype=np.uint64)
        for l in range(n_layers):
            ps = flat[5 + 16 * l: 5 + 16 * (l + 1)]
            w[l] = [t.data_ptr() for t in cw["layers"][l]]
            sm[l] = [bqkv[l].data_ptr()] + [ps[j].data_ptr() for j in (7, 8, 9, 11, 13, 14, 15)]   # b_ao g1 b1 b_i b_o g2 b2
        self.w, self.sm = w, sm
        # ---- gradient arena (offsets in floats, every slice 4-float aligned: TMA reduce-add targets)
        off = 0
        views = [None] * len(flat)                   # per parameter: (offset, shape)

        def take(shape):
            nonlocal off
            o = off
            off += (math.prod(shape) + 3) // 4 * 4
            return o
        H = flat[0].shape[1]
        views[-2] = (take(flat[-2].shape), tuple(flat[-2].shape))
        views[-1] = (take(flat[-1].shape), tuple(flat[-1].shape))
        self.fc_end = off
        lay = np.zeros((n_layers, 12), dtype=np.int64)        # dwqkv dbqkv dw_ao db_ao dg1 db1 dw_i db_i dw_o db_o dg2 db2
        self.layer_end = [0] * n_layers
        for l in reversed(range(n_layers)):
            base = 5 + 16 * l
            ps = flat[base: base + 16]
            o_w, o_b = take((3 * H, H)), take((3 * H,))
            lay[l, 0], lay[l, 1] = o_w, o_b
            for j in range(3):
                views[base + 2 * j] = (o_w + j * H * H, (H, H))
                views[base + 2 * j + 1] = (o_b + j * H, (H,))
            for col, j in ((2, 6), (3, 7), (4, 8), (5, 9), (6, 10), (7, 11), (8, 12), (9, 13), (10, 14), (11, 15)):
                o = take(ps[j].shape)
                lay[l, col] = o
                views[base + j] = (o, tuple(ps[j].shape))
            self.layer_end[l] = off
        self.lay = lay
        for i in (3, 4):
            views[i] = (take(flat[i].shape), tuple(flat[i].shape))
        views[2] = (take(flat[2].shape), tuple(flat[2].shape))      # token-type table [2, H]: row 0 receives the column sum
        views[0] = (take(flat[0].shape), tuple(flat[0].shape))
        views[1] = (take(flat[1].shape), tuple(flat[1].shape))
        self.views = views
        self.total = off

    @staticmethod
    def make_key(cw, bqkv):
        return tuple(t.data_ptr() for lay in cw["layers"] for t in lay) + (cw["fc"].data_ptr(),) + tuple(b.data_ptr() for b in bqkv)


class BertTowerFn(torch.autograd.Function):
    """E[n_seq, D] = GELU(fc(BERT(tokens)[CLS])) over PACKED tokens (pad tokens and pad items are never computed).

    params (flat): word, pos, type, emb_ln_w, emb_ln_b,
        per layer: q_w,q_b,k_w,k_b,v_w,v_b, ao_w,ao_b, ln1_w,ln1_b, i_w,i_b, o_w,o_b, ln2_w,ln2_b,   then fc_w, fc_b.
    meta: n_layers, n_heads, eps, max_len, adt (activation dtype), drop (DropCtx)
    """

    @staticmethod
    @_with_prec
    def forward(ctx, meta, tok_ids, tok_pos, cu_seqlens, cls_rows, *params):
        n_layers, n_heads, eps, max_len, adt, drop = (meta["n_layers"], meta["n_heads"], meta["eps"], meta["max_len"],
                                                      meta["adt"], meta["drop"])
        word, posw, typew, eg, eb = params[:5]
        fc_w, fc_b = params[-2:]
        H = word.shape[1]
        n_tok = tok_ids.numel()
        n_seq = cu_seqlens.numel() - 1
        dh = H // n_heads
        dev = word.device
        scale = 1.0 / math.sqrt(dh)
        es = 4 if adt == torch.float32 else 2
        I0 = params[5 + 10].shape[0]
        ws = _ws(_WS_FWD, dev)
        own_ws = ws.acquire(dev, n_tok * es * (n_layers * (6 * H + 2 * I0) + 6 * H) + n_layers * n_tok * 8
                            + (n_layers * 8 + 16) * 256)
        new = (lambda shape, dtype=adt: ws.take(shape, dtype)) if own_ws else \
              (lambda shape, dtype=adt: torch.empty(shape, device=dev, dtype=dtype))
        # ---- embeddings: gather + LN (+ dropout after LN: HF BertEmbeddings)
        z = new((n_tok, H))
        lib.bert_embed_fwd(tok_ids, tok_pos, word.detach(), posw.detach(), typew.detach()[0].contiguous(), z)
        x, x_pre, rstd0 = lib.layernorm_fwd(z, eg.detach(), eb.detach(), eps, p_post=drop.p_hidden, seed=drop.seed,
                                            off_post=drop.off(0), out=new((n_tok, H)),
                                            y_pre=new((n_tok, H)) if drop.p_hidden > 0 else None)
        emb_saved = (x_pre if x_pre is not None else x, rstd0)
        del z
        layers = []
        cw = meta.get("cw")                     # weights already in the compute dtype (prepare_tower_weights)
        layout = meta.get("layout")
        use_seq = not lib._GEMM_TIMING and layout is not None   # C++ layer sequencer; the per-kernel path is kept for the roofline leg
        seq_saved = None
        if use_seq:
            # ONE C-ABI call for all layers: a structured numpy array of per-layer records, filled column-wise
            I = I0
            szh, sz3, szi, szr = _al256(n_tok * H * es), _al256(n_tok * 3 * H * es), _al256(n_tok * I * es), _al256(2 * n_tok * 4)
            block = sz3 + 3 * szh + 2 * szi + szr
            tmp_h = new((n_tok, H))
            blk = new((n_layers * block,), torch.uint8)
            rec = np.zeros(n_layers, dtype=lib.LAYER_FWD_DT)
            ls = np.arange(n_layers, dtype=np.uint64)
            b = np.uint64(blk.data_ptr()) + ls * np.uint64(block)
            rec["n_tok"], rec["n_seq"], rec["H"], rec["I"], rec["n_heads"], rec["max_len"] = n_tok, n_seq, H, I, n_heads, max_len
            rec["dtype"] = lib.DT_BF16 if adt == torch.bfloat16 else lib.DT_F16 if adt == torch.float16 else (2 if meta.get("x3", True) else 0)
            rec["eps"], rec["p_hidden"], rec["p_attn"] = eps, drop.p_hidden, drop.p_attn
            rec["seed"] = np.uint64(drop.seed & 0xFFFFFFFFFFFFFFFF)
            site = (np.uint64(1) + np.uint64(4) * ls) << np.uint64(36)
            dbase = np.uint64(drop.base & 0xFFFFFFFFFFFFFFFF)
            rec["off_attn"], rec["off_ln1"], rec["off_ln2"] = dbase + site, dbase + site + (np.uint64(1) << np.uint64(36)), \
                dbase + site + (np.uint64(2) << np.uint64(36))
            rec["cu_seqlens"], rec["tmp_h"] = cu_seqlens.data_ptr(), tmp_h.data_ptr()
            rec["wqkv"], rec["w_ao"], rec["w_i"], rec["w_o"] = (layout.w[:, j] for j in range(4))
            for j, nme in enumerate(("bqkv", "b_ao", "g1", "b1", "b_i", "b_o", "g2", "b2")):
                rec[nme] = layout.sm[:, j]
            rec["qkv"], rec["ctx"], rec["x1"], rec["x2"] = b, b + np.uint64(sz3), b + np.uint64(sz3 + szh), b + np.uint64(sz3 + 2 * szh)
            rec["pre"] = b + np.uint64(sz3 + 3 * szh)
            rec["act"] = rec["pre"] + np.uint64(szi)
            rec["rstd1"] = rec["act"] + np.uint64(szi)
            rec["rstd2"] = rec["rstd1"] + np.uint64(4 * n_tok)
            rec["x"][0] = x.data_ptr()
            rec["x"][1:] = rec["x2"][:-1]
            lib.bert_layers_fwd(rec)
            x0 = x
            o = (n_layers - 1) * block + sz3 + 2 * szh
            x = blk[o:o + n_tok * H * es].view(adt).view(n_tok, H)
            seq_saved = (rec, blk, x0, tmp_h)
        for l in range(n_layers if not use_seq else 0):
            (qw, qb, kw, kb, vw, vb, aow, aob, g1, b1, iw, ib, ow, ob, g2, b2) = params[5 + 16 * l: 5 + 16 * (l + 1)]
            bqkv = meta["bqkv"][l]                                       # fused [3H] (FusedParamGroup)
            if cw is not None:
                wqkv, w_ao, w_i, w_o = cw["layers"][l]
            else:
                wqkv = _cw(meta["wqkv"][l], adt)                         # fused [3H, H]
                w_ao, w_i, w_o = _cw(aow, adt), _cw(iw, adt), _cw(ow, adt)
            I = iw.shape[0]
            qkv = lib.linear_fwd(x, wqkv, bqkv)                        # [n_tok, 3H]
            ctxo = torch.empty(n_tok, H, device=dev, dtype=adt)
            (lib.attn_fwd if max_len <= 32 else lib.attn_gen_fwd)(
                qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctxo, cu_seqlens=cu_seqlens, n_seq=n_seq, seqlen=max_len,
                n_heads=n_heads, head_dim=dh, scale=scale, dropout_p=drop.p_attn, seed=drop.seed,
                offset=drop.off(1 + 4 * l))
            ao = lib.linear_fwd(ctxo, w_ao, aob.detach())
            x1, _, rstd1 = lib.layernorm_fwd(ao, g1.detach(), b1.detach(), eps, residual=x, p_pre=drop.p_hidden,
                                             seed=drop.seed, off_pre=drop.off(2 + 4 * l))
            del ao
            pre = torch.empty(n_tok, I, device=dev, dtype=adt)
            act = lib.linear_fwd(x1, w_i, ib.detach(), epilogue=lib.EPI_GELU_DGELU, pre=pre)   # pre <- gelu'(z)
            fo = lib.linear_fwd(act, w_o, ob.detach())
            x2, _, rstd2 = lib.layernorm_fwd(fo, g2.detach(), b2.detach(), eps, residual=x1, p_pre=drop.p_hidden,
                                             seed=drop.seed, off_pre=drop.off(3 + 4 * l))
            del fo
            layers.append([x, qkv, ctxo, x1, rstd1, pre, act, rstd2, (wqkv, w_ao, w_i, w_o)])
            x = x2
        # ---- CLS pooling + fc + GELU
        cls = lib.gather_rows(x, cls_rows)                                # [n_seq, H]
        D = fc_w.shape[0]
        fc_pre = torch.empty(n_seq, D, device=dev, dtype=adt)
        w_fc = cw["fc"] if cw is not None else _cw(fc_w, adt)
        E = lib.linear_fwd(cls, w_fc, fc_b.detach(), epilogue=lib.EPI_GELU, pre=fc_pre)
        ctx.meta = meta
        ctx.own_ws = own_ws
        if own_ws and not torch.is_grad_enabled():
            ws.release()                      # no backward will come (eval / no_grad): nothing stays saved
            ctx.own_ws = False
        ctx.saved = dict(emb=emb_saved, layers=layers, x_last=x, cls=cls, fc_pre=fc_pre, w_fc=w_fc, seq=seq_saved)
        ctx.idx = (tok_ids, tok_pos, cu_seqlens, cls_rows)
        ctx.params = params
        return E

    @staticmethod
    @_with_prec
    def backward(ctx, dE):
        meta, saved, params = ctx.meta, ctx.saved, ctx.params
        tok_ids, tok_pos, cu_seqlens, cls_rows = ctx.idx
        n_layers, n_heads, eps, max_len, adt, drop = (meta["n_layers"], meta["n_heads"], meta["eps"], meta["max_len"],
                                                      meta["adt"], meta["drop"])
        need = [p.requires_grad for p in params]
        word, posw, typew, eg, eb = params[:5]
        fc_w, fc_b = params[-2:]
        H = word.shape[1]
        n_tok = tok_ids.numel()
        n_seq = cu_seqlens.numel() - 1
        dh = H // n_heads
        dev = word.device
        scale = 1.0 / math.sqrt(dh)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        dE = dE.contiguous().to(adt)
        if saved.get("seq") is not None:
            return BertTowerFn._backward_seq(ctx, dE)
        arena = _Arena(dev, [p.shape for p in params] + [(H,)])
        _z = lambda p: arena.take(p.shape)   # noqa: E731  zero-initialised gradient slice
        sync = _GradSync(arena) if _GradSync.enabled(meta) else None
        # ---- fc + GELU backward
        cls, fc_pre = saved["cls"], saved["fc_pre"]
        dpre = lib.act_bwd(dE, fc_pre, 0)
        g_fcw = _z(fc_w)
        lib.linear_wgrad(dpre, cls, g_fcw)
        g_fcb = _z(fc_b)
        lib.colsum(dpre, g_fcb)
        grads[-2], grads[-1] = (g_fcw if need[-2] else None), (g_fcb if need[-1] else None)
        if sync:
            sync.flush()
        dcls = lib.linear_dgrad(dpre, saved["w_fc"])
        # ---- scatter CLS grads into the gradient of the last hidden state
        dx32 = torch.zeros(n_tok, H, device=dev, dtype=torch.float32)
        lib.scatter_add_rows(dcls, cls_rows, dx32)
        dx = dx32 if adt == torch.float32 else dx32.to(adt)
        del dx32
        x_out = saved["x_last"]
        dx2 = None                         # second addend of the running hidden-state gradient (sequencer path)
        use_seq = not lib._GEMM_TIMING
        wsb = _ws(_WS_BWD, dev)
        es = 4 if adt == torch.float32 else 2
        I = params[5 + 10].shape[0]
        own_b = use_seq and wsb.acquire(dev, n_tok * es * (12 * H + I) + 64 * 256)
        newb = (lambda shape, dtype=adt: wsb.take(shape, dtype)) if own_b else \
               (lambda shape, dtype=adt: torch.empty(shape, device=dev, dtype=dtype))
        if use_seq:
            ws_h = newb((4 if drop.p_hidden > 0 else 3, n_tok, H))   # dz2, dx1b, dctx, (dbr)
            dpre_ws = newb((n_tok, I))
            dqkv_ws = newb((n_tok, 3 * H))
            pp = [newb((n_tok, H)) for _ in range(4)]                # ping-pong pairs for (dz1, dxq)
        for l in reversed(range(n_layers)):
            (qw, qb, kw, kb, vw, vb, aow, aob, g1, b1, iw, ib, ow, ob, g2, b2) = params[5 + 16 * l: 5 + 16 * (l + 1)]
            x, qkv, ctxo, x1, rstd1, pre, act, rstd2, (wqkv, w_ao, w_i, w_o) = saved["layers"][l]
            base = 5 + 16 * l
            # gradients of the three projection weights / biases are row blocks of one fused buffer
            dwqkv, dbqkv = arena.take((3 * H, H)), arena.take((3 * H,))
            if use_seq:
                dg2, db2, dob, dow, dib, diw = _z(g2), _z(b2), _z(ob), _z(ow), _z(ib), _z(iw)
                dg1, db1, daob, daow = _z(g1), _z(b1), _z(aob), _z(aow)
                small = (meta["bqkv"][l], aob.detach(), g1.detach(), b1.detach(), ib.detach(), ob.detach(), g2.detach(),
                         b2.detach())
                b = lib.BertLayerBwd()
                b.fwd = _layer_struct(meta, l, n_tok, n_seq, H, iw.shape[0], cu_seqlens, (wqkv, w_ao, w_i, w_o), small,
                                      (x, qkv, ctxo, x1, rstd1, pre, act, x_out, rstd2), ws_h[0])
                dz1, dxq = pp[2 * (l & 1)], pp[2 * (l & 1) + 1]
                b.dy, b.dy2 = dx.data_ptr(), (dx2.data_ptr() if dx2 is not None else None)
                b.dz1, b.dxq = dz1.data_ptr(), dxq.data_ptr()
                b.dz2, b.dx1b, b.dctx = ws_h[0].data_ptr(), ws_h[1].data_ptr(), ws_h[2].data_ptr()
                b.dbr = ws_h[3].data_ptr() if drop.p_hidden > 0 else None
                b.dpre, b.dqkv = dpre_ws.data_ptr(), dqkv_ws.data_ptr()
                (b.dwqkv, b.dbqkv, b.dw_ao, b.db_ao, b.dg1, b.db1, b.dw_i, b.db_i, b.dw_o, b.db_o, b.dg2, b.db2) = (
                    t.data_ptr() for t in (dwqkv, dbqkv, daow, daob, dg1, db1, diw, dib, dow, dob, dg2, db2))
                lib.bert_layer_bwd(b)
                dx, dx2 = dz1, dxq
                x_out = x
                lay = (dwqkv[:H], dbqkv[:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:], dbqkv[2 * H:],
                       daow, daob, dg1, db1, diw, dib, dow, dob, dg2, db2)
                for j, g in enumerate(lay):
                    grads[base + j] = g if need[base + j] else None
                saved["layers"][l] = None
                if sync:
                    sync.flush()          # this layer's gradients: all-reduce while the layers below are differentiated
                continue
            # output LayerNorm (y = x_out)
            dg2, db2, dob = _z(g2), _z(b2), _z(ob)
            dz2, dfo = lib.layernorm_bwd(dx, x_out, g2.detach(), b2.detach(), rstd2, dgamma=dg2, dbeta=db2, dbias=dob,
                                         p_pre=drop.p_hidden, seed=drop.seed, off_pre=drop.off(3 + 4 * l))
            del dx
            # FFN2 / FFN1
            dow = _z(ow)
            lib.linear_wgrad(dfo, act, dow)
            dpre_i = lib.linear_dgrad(dfo, w_o, epilogue=lib.EPI_MUL_AUX, aux=pre)
            if dfo is not dz2:
                del dfo
            dib, diw = _z(ib), _z(iw)
            lib.colsum(dpre_i, dib)
            lib.linear_wgrad(dpre_i, x1, diw)
            dx1_b = lib.linear_dgrad(dpre_i, w_i)
            del dpre_i
            # attention-output LayerNorm: dy = dz2 + dx1_b, y = x1
            dg1, db1, daob = _z(g1), _z(b1), _z(aob)
            dz1, dao = lib.layernorm_bwd(dz2, x1, g1.detach(), b1.detach(), rstd1, dy2=dx1_b, dgamma=dg1, dbeta=db1,
                                         dbias=daob, p_pre=drop.p_hidden, seed=drop.seed, off_pre=drop.off(2 + 4 * l))
            del dz2, dx1_b
            daow = _z(aow)
            lib.linear_wgrad(dao, ctxo, daow)
            dctx = lib.linear_dgrad(dao, w_ao)
            if dao is not dz1:
                del dao
            # attention core
            dqkv = torch.empty_like(qkv)
            (lib.attn_bwd if max_len <= 32 else lib.attn_gen_bwd)(
                qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], dctx, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:],
                cu_seqlens=cu_seqlens, n_seq=n_seq, seqlen=max_len, n_heads=n_heads, head_dim=dh, scale=scale,
                dropout_p=drop.p_attn, seed=drop.seed, offset=drop.off(1 + 4 * l))
            del dctx
            # fused QKV projection backward
            lib.linear_wgrad(dqkv, x, dwqkv)
            lib.colsum(dqkv, dbqkv)
            dx = _dgrad_acc(dqkv, wqkv, dz1)
            del dqkv, dz1
            x_out = x
            lay = (dwqkv[:H], dbqkv[:H], dwqkv[H:2 * H], dbqkv[H:2 * H], dwqkv[2 * H:], dbqkv[2 * H:],
                   daow, daob, dg1, db1, diw, dib, dow, dob, dg2, db2)
            for j, g in enumerate(lay):
                grads[base + j] = g if need[base + j] else None
            saved["layers"][l] = None
            if sync:
                sync.flush()
        # ---- embeddings backward: y = dropout(LN(z)), z = word + pos + type
        y_emb, rstd0 = saved["emb"]
        deg, deb = _z(eg), _z(eb)
        dtype_sum = arena.take((H,))
        dz0, _ = lib.layernorm_bwd(dx, y_emb, eg.detach(), eb.detach(), rstd0, dy2=dx2, dgamma=deg, dbeta=deb,
                                   dbias=dtype_sum, p_post=drop.p_hidden, seed=drop.seed, off_post=drop.off(0))
        dword = _z(word) if need[0] else None
        dposw = _z(posw) if need[1] else None
        if dword is not None or dposw is not None:
            lib.bert_embed_bwd(dz0, tok_ids, tok_pos, dword, dposw)
        dtypew = None
        if need[2]:
            dtypew = _z(typew)
            dtypew[0].copy_(dtype_sum)
        grads[0], grads[1], grads[2] = dword, dposw, dtypew
        grads[3] = deg if need[3] else None
        grads[4] = deb if need[4] else None
        if sync:
            sync.wait()
        ctx.saved = None
        if own_b:
            wsb.release()
        if getattr(ctx, "own_ws", False):
            _ws(_WS_FWD, dev).release()
            ctx.own_ws = False
        return (None, None, None, None, None) + tuple(grads)


    @staticmethod
    def _backward_seq(ctx, dE):
        """backward on the C++ layer sequencer: the gradient arena and every per-layer record are laid out by
        TowerLayout, all layers run in ONE C-ABI call (or one per layer when the layer-wise gradient all-reduce of a
        multi-GPU run interleaves with them)"""
        meta, saved, params = ctx.meta, ctx.saved, ctx.params
        tok_ids, tok_pos, cu_seqlens, cls_rows = ctx.idx
        n_layers, adt, drop, eps = meta["n_layers"], meta["adt"], meta["drop"], meta["eps"]
        lay = meta["layout"]
        rec, blk, x0, _tmp = saved["seq"]
        need = [p.requires_grad for p in params]
        word, posw, typew, eg, eb = params[:5]
        fc_w, fc_b = params[-2:]
        H = word.shape[1]
        I = params[5 + 10].shape[0]
        n_tok = tok_ids.numel()
        dev = word.device
        es = 4 if adt == torch.float32 else 2
        buf = torch.zeros(max(lay.total, 4), device=dev, dtype=torch.float32)     # ONE memset for every parameter gradient
        abase = buf.data_ptr()

        def view(i):
            o, shape = lay.views[i]
            return buf[o:o + math.prod(shape)].view(shape)
        arena = type("_FlatArena", (), {})()
        arena.buf, arena.off = buf, 0
        sync = _GradSync(arena) if _GradSync.enabled(meta) else None
        # ---- fc + GELU backward
        cls, fc_pre = saved["cls"], saved["fc_pre"]
        dpre = lib.act_bwd(dE, fc_pre, 0)
        lib.linear_wgrad(dpre, cls, view(-2))
        lib.colsum(dpre, view(-1))
        if sync:
            arena.off = lay.fc_end
            sync.flush()
        dcls = lib.linear_dgrad(dpre, saved["w_fc"])
        dx32 = torch.zeros(n_tok, H, device=dev, dtype=torch.float32)
        lib.scatter_add_rows(dcls, cls_rows, dx32)
        dx = dx32 if adt == torch.float32 else dx32.to(adt)
        del dx32
        # ---- layers, last to first
        wsb = _ws(_WS_BWD, dev)
        own_b = wsb.acquire(dev, n_tok * es * (12 * H + I) + 64 * 256)
        newb = (lambda shape, dtype=adt: wsb.take(shape, dtype)) if own_b else \
               (lambda shape, dtype=adt: torch.empty(shape, device=dev, dtype=dtype))
        ws_h = newb((4 if drop.p_hidden > 0 else 3, n_tok, H))   # dz2, dx1b, dctx, (dbr)
        dpre_ws, dqkv_ws = newb((n_tok, I)), newb((n_tok, 3 * H))
        pp = [newb((n_tok, H)) for _ in range(4)]                # ping-pong pairs for (dz1, dxq)
        recb = np.zeros(n_layers, dtype=lib.LAYER_BWD_DT)
        order = np.arange(n_layers - 1, -1, -1)                  # record k <-> layer order[k]
        recb["fwd"] = rec[order]
        recb["fwd"]["tmp_h"] = ws_h[0].data_ptr()
        par = (order & 1).astype(np.int64)
        p0, p1, p2, p3 = (t.data_ptr() for t in pp)
        recb["dz1"] = np.where(par == 0, p0, p2).astype(np.uint64)
        recb["dxq"] = np.where(par == 0, p1, p3).astype(np.uint64)
        recb["dy"][0], recb["dy2"][0] = dx.data_ptr(), 0
        recb["dy"][1:], recb["dy2"][1:] = recb["dz1"][:-1], recb["dxq"][:-1]
        recb["dz2"], recb["dx1b"], recb["dctx"] = ws_h[0].data_ptr(), ws_h[1].data_ptr(), ws_h[2].data_ptr()
        recb["dbr"] = ws_h[3].data_ptr() if drop.p_hidden > 0 else 0
        recb["dpre"], recb["dqkv"] = dpre_ws.data_ptr(), dqkv_ws.data_ptr()
        g = (np.uint64(abase) + (lay.lay[order] * 4).astype(np.uint64))
        for j, nme in enumerate(("dwqkv", "dbqkv", "dw_ao", "db_ao", "dg1", "db1", "dw_i", "db_i", "dw_o", "db_o", "dg2", "db2")):
            recb[nme] = g[:, j]
        if sync is None:
            lib.bert_layers_bwd(recb)
        else:
            # a layer's gradients are all-reduced while the layers below run.  Its weight gradients are produced on the
            # library's side stream; the main stream is ordered after them once the NEXT layer has been enqueued (that
            # layer waits for every scratch buffer the side work reads), so the collective of layer k is issued one
            # layer late instead of joining the two streams after every layer (which serialised them: the 2-GPU step
            # gained nothing from the side stream)
            for k in range(n_layers):
                lib.bert_layers_bwd(recb[k:k + 1], join=False)
                if k >= 1:
                    arena.off = lay.layer_end[int(order[k - 1])]
                    sync.flush()
            lib.bert_layers_bwd(None, join=True)
            arena.off = lay.layer_end[int(order[n_layers - 1])]
            sync.flush()
        last = int(order[-1]) & 1
        dx, dx2 = pp[2 * last], pp[2 * last + 1]
        # ---- embeddings backward: y = dropout(LN(z)), z = word + pos + type
        y_emb, rstd0 = saved["emb"]
        dtypew = view(2)
        dz0, _ = lib.layernorm_bwd(dx, y_emb, eg.detach(), eb.detach(), rstd0, dy2=dx2, dgamma=view(3), dbeta=view(4),
                                   dbias=dtypew[0], p_post=drop.p_hidden, seed=drop.seed, off_post=drop.off(0))
        dword = view(0) if need[0] else None
        dposw = view(1) if need[1] else None
        if dword is not None or dposw is not None:
            lib.bert_embed_bwd(dz0, tok_ids, tok_pos, dword, dposw)
        grads = [view(i) if need[i] else None for i in range(len(params))]
        if sync:
            arena.off = lay.total
            sync.wait()
        ctx.saved = None
        if own_b:
            wsb.release()
        if getattr(ctx, "own_ws", False):
            _ws(_WS_FWD, dev).release()
            ctx.own_ws = False
        return (None, None, None, None, None) + tuple(grads)


# ==================================================================================================
# row gather with index (ID embedding, slot expansion, input slicing)
# ==================================================================================================
class GatherRowsFn(torch.autograd.Function):
    """out[i] = idx[i] >= 0 ? src[idx[i]] : 0 ; backward scatter-adds (fp32 atomics)."""

    @staticmethod
    def forward(ctx, src, idx, out_dtype):
        ctx.save_for_backward(idx)
        ctx.src_shape = src.shape
        ctx.src_dtype = src.dtype
        return lib.gather_rows(src.detach(), idx, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        d = torch.zeros(ctx.src_shape, device=dout.device, dtype=torch.float32)
        lib.scatter_add_rows(dout.contiguous(), idx, d)
        return (d if ctx.src_dtype == torch.float32 else d.to(ctx.src_dtype)), None, None


# ==================================================================================================
# SASRec user encoder
# ==================================================================================================
class SasrecFn(torch.autograd.Function):
    """H[B*L, D] = TransformerEncoder(X[B*L, D], log_mask[B, L])     (model/modules.py:89-96, post-LN blocks)

    params (flat): pos_emb, ln_w, ln_b, per block: wq, wk, wv, fc, ln1_w, ln1_b, w1, b1, w2, b2, ln2_w, ln2_b
    meta: n_blocks, n_heads, L, adt, drop (DropCtx: p_hidden = p_attn = drop_rate)
    """

    @staticmethod
    @_with_prec
    def forward(ctx, meta, X, log_mask, *params):
        n_blocks, n_heads, L, adt, drop = meta["n_blocks"], meta["n_heads"], meta["L"], meta["adt"], meta["drop"]
        posw, g0, b0 = params[:3]
        R, D = X.shape
        B = R // L
        dk = D // n_heads
        scale = 1.0 / (dk ** 0.5)
        dev = X.device
        X = X.detach().contiguous()
        h, h_pre, rstd0 = lib.layernorm_fwd(X, g0.detach(), b0.detach(), 1e-6, pos=posw.detach(), pos_period=L,
                                            p_post=drop.p_hidden, seed=drop.seed, off_post=drop.off(0))
        emb_saved = (h_pre if h_pre is not None else h, rstd0)
        blocks = []
        for l in range(n_blocks):
            (wq, wk, wv, fc, g1, b1, w1, bb1, w2, bb2, g2, b2) = params[3 + 12 * l: 3 + 12 * (l + 1)]
            if meta.get("cw") is not None:                              # persistent 16-bit shadows (ShadowSet)
                wqkv, w_fc, w_1, w_2 = meta["cw"][4 * l: 4 * l + 4]
            else:
                wqkv = _cw(meta["wqkv"][l], adt)                        # fused [3D, D] (FusedParamGroup)
                w_fc, w_1, w_2 = _cw(fc, adt), _cw(w1, adt), _cw(w2, adt)
            qkv = lib.linear_fwd(h, wqkv)
            ctxo = torch.empty(R, D, device=dev, dtype=adt)
            lib.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], ctxo, key_mask=log_mask, causal=True, n_seq=B,
                         seqlen=L, n_heads=n_heads, head_dim=dk, scale=scale, masked_add=_MASKED_ADD,
                         dropout_p=drop.p_attn, seed=drop.seed, offset=drop.off(1 + 4 * l))
            ao = lib.linear_fwd(ctxo, w_fc)
            h1, _, rstd1 = lib.layernorm_fwd(ao, g1.detach(), b1.detach(), 1e-6, residual=h, p_pre=drop.p_hidden,
                                             seed=drop.seed, off_pre=drop.off(2 + 4 * l))
            u = lib.linear_fwd(h1, w_1, bb1.detach(), epilogue=lib.EPI_RELU)
            fo = lib.linear_fwd(u, w_2, bb2.detach())
            h2, _, rstd2 = lib.layernorm_fwd(fo, g2.detach(), b2.detach(), 1e-6, residual=h1, p_pre=drop.p_hidden,
                                             seed=drop.seed, off_pre=drop.off(3 + 4 * l))
            blocks.append([h, qkv, ctxo, h1, rstd1, u, rstd2, (wqkv, w_fc, w_1, w_2)])
            h = h2
        ctx.meta = meta
        ctx.saved = dict(emb=emb_saved, blocks=blocks, h_last=h, log_mask=log_mask)
        ctx.params = params
        return h

    @staticmethod
    @_with_prec
    def backward(ctx, dH):
        meta, saved, params = ctx.meta, ctx.saved, ctx.params
        n_blocks, n_heads, L, adt, drop = meta["n_blocks"], meta["n_heads"], meta["L"], meta["adt"], meta["drop"]
        need = [p.requires_grad for p in params]
        posw, g0, b0 = params[:3]
        log_mask = saved["log_mask"]
        dx = dH.contiguous().to(adt)
        R, D = dx.shape
        B = R // L
        dk = D // n_heads
        scale = 1.0 / (dk ** 0.5)
        dev = dx.device
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        arena = _Arena(dev, [p.shape for p in params])
        _z = lambda p: arena.take(p.shape)   # noqa: E731
        h_out = saved["h_last"]
        for l in reversed(range(n_blocks)):
            (wq, wk, wv, fc, g1, b1, w1, bb1, w2, bb2, g2, b2) = params[3 + 12 * l: 3 + 12 * (l + 1)]
            h, qkv, ctxo, h1, rstd1, u, rstd2, (wqkv, w_fc, w_1, w_2) = saved["blocks"][l]
            base = 3 + 12 * l
            dwqkv = arena.take((3 * D, D))
            dg2, db2, dbb2 = _z(g2), _z(b2), _z(bb2)
            dz2, dfo = lib.layernorm_bwd(dx, h_out, g2.detach(), b2.detach(), rstd2, dgamma=dg2, dbeta=db2, dbias=dbb2,
                                         p_pre=drop.p_hidden, seed=drop.seed, off_pre=drop.off(3 + 4 * l))
            dw2 = _z(w2)
            lib.linear_wgrad(dfo, u, dw2)
            du = lib.linear_dgrad(dfo, w_2, epilogue=lib.EPI_MUL_RELU_GRAD, aux=u)
            dbb1, dw1 = _z(bb1), _z(w1)
            lib.colsum(du, dbb1)
            lib.linear_wgrad(du, h1, dw1)
            dh1_b = lib.linear_dgrad(du, w_1)
            dg1, db1 = _z(g1), _z(b1)
            dz1, dao = lib.layernorm_bwd(dz2, h1, g1.detach(), b1.detach(), rstd1, dy2=dh1_b, dgamma=dg1, dbeta=db1,
                                         p_pre=drop.p_hidden, seed=drop.seed, off_pre=drop.off(2 + 4 * l))
            dfc = _z(fc)
            lib.linear_wgrad(dao, ctxo, dfc)
            dctx = lib.linear_dgrad(dao, w_fc)
            dqkv = torch.empty_like(qkv)
            lib.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], dctx, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                         key_mask=log_mask, causal=True, n_seq=B, seqlen=L, n_heads=n_heads, head_dim=dk, scale=scale,
                         masked_add=_MASKED_ADD, dropout_p=drop.p_attn, seed=drop.seed, offset=drop.off(1 + 4 * l))
            lib.linear_wgrad(dqkv, h, dwqkv)
            dx = _dgrad_acc(dqkv, wqkv, dz1)
            h_out = h
            lay = (dwqkv[:D], dwqkv[D:2 * D], dwqkv[2 * D:], dfc, dg1, db1, dw1, dbb1, dw2, dbb2, dg2, db2)
            for j, g in enumerate(lay):
                grads[base + j] = g if need[base + j] else None
            saved["blocks"][l] = None
        y0, rstd0 = saved["emb"]
        dg0, db0 = _z(g0), _z(b0)
        dpos = _z(posw)
        dX, _ = lib.layernorm_bwd(dx, y0, g0.detach(), b0.detach(), rstd0, dgamma=dg0, dbeta=db0, dpos=dpos,
                                  pos_period=L, p_post=drop.p_hidden, seed=drop.seed, off_post=drop.off(0))
        # position table may be longer than L rows (it is exactly L in the reference, modules.py:82)
        grads[0] = dpos if need[0] else None
        grads[1] = dg0 if need[1] else None
        grads[2] = db0 if need[2] else None
        ctx.saved = None
        return (None, dX, None) + tuple(grads)


# ==================================================================================================
# in-batch debiased cross-entropy
# ==================================================================================================
class InbatchCEFn(torch.autograd.Function):
    """loss = CE over valid rows of  S = P.E^T - log_pop  with the reject mask  (model/model.py:45-67).

    P [B*L, D] (rows of the local users), E [C, D] (all score columns), member/pad from lib.inbatch_mask.
    Returns (loss, sum_cnt): loss is the mean over the local valid rows; sum_cnt = [sum of valid row losses, n_valid].
    `n_valid_override` (device scalar) replaces the local count in BOTH directions: loss = sum / override (the rank's
    share of a batch-global mean, parallel.py `global` mode).
    """

    @staticmethod
    def forward(ctx, meta, P, E, member, pad, log_pop, log_mask, B, L, col_offset, n_valid_override):
        ctx.meta = meta
        with lib.fp32_mode(meta.get("x3", True)):
            return InbatchCEFn._fwd(ctx, P, E, member, pad, log_pop, log_mask, B, L, col_offset, n_valid_override)

    @staticmethod
    def _fwd(ctx, P, E, member, pad, log_pop, log_mask, B, L, col_offset, n_valid_override):
        Pc, Ec = P.detach().contiguous(), E.detach().contiguous()
        loss, row_lse, row_loss, sum_cnt, _ = lib.inbatch_ce_fwd(Pc, Ec, member, pad, log_pop, log_mask, B, L, col_offset)
        if n_valid_override is not None:
            # global normalisation: this rank's share of the batch-global mean
            loss = (sum_cnt[0] / n_valid_override.reshape(())).reshape(())
        ctx.save_for_backward(Pc, Ec, member, pad, log_pop, log_mask, row_lse, sum_cnt)
        ctx.dims = (B, L, col_offset)
        ctx.n_valid_override = n_valid_override
        ctx.mark_non_differentiable(sum_cnt)
        return loss, sum_cnt

    @staticmethod
    def backward(ctx, dloss, _unused):
        with lib.fp32_mode(ctx.meta.get("x3", True)):
            return InbatchCEFn._bwd(ctx, dloss)

    @staticmethod
    def _bwd(ctx, dloss):
        Pc, Ec, member, pad, log_pop, log_mask, row_lse, sum_cnt = ctx.saved_tensors
        B, L, col_offset = ctx.dims
        n_valid = ctx.n_valid_override if ctx.n_valid_override is not None else sum_cnt[1:2]
        g = dloss.detach().reshape(1).to(torch.float32).contiguous()
        dS = lib.inbatch_ce_dlogits(Pc, Ec, member, pad, log_pop, log_mask, row_lse, g, n_valid.contiguous(), B, L,
                                    col_offset)
        R, D = Pc.shape
        C = Ec.shape[0]
        # dP = dS . E      (A = dS [R,C] K-major, B = E stored [C,D] = [K,N] -> MN-major)
        dP = torch.empty(R, D, device=Pc.device, dtype=Pc.dtype)
        lib.gemm(dS, Ec, dP, M=R, N=D, K=C, lda=dS.stride(0), ldb=Ec.stride(0), ldc=D, a_mn=False, b_mn=True)
        # dE = dS^T . P    (A = dS stored [R,C] = [K,M] -> MN-major, B = P stored [R,D] = [K,N] -> MN-major)
        dE = torch.empty(C, D, device=Pc.device, dtype=Pc.dtype)
        lib.gemm(dS, Pc, dE, M=C, N=D, K=R, lda=dS.stride(0), ldb=Pc.stride(0), ldc=D, a_mn=True, b_mn=True)
        return None, dP, dE, None, None, None, None, None, None, None, None
