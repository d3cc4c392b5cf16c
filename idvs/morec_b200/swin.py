"""Swin Transformer item tower on the morec_b200 kernels (forward + hand-written backward).

Replaces `GELU(SwinForImageClassification(pixel_values)[0])` of the reference's vision package
(inbatch_sasrec_e2e_vision/model/encoders.py:24-31; model surgery in inbatch_sasrec_e2e_vision/run.py:47-60) for the
HF architecture (third-party `transformers`, modeling_swin.py): patch-embed conv 4x4/4 as an im2col GEMM + LN; per
stage `depth` pre-LN blocks  x += drop_path(W-MSA/SW-MSA(LN(x)));  x += MLP(LN(x))  with 7x7 windows, relative-position
bias and the shifted-window mask; patch merging (2x2 concat -> LN -> Linear 4C->2C, no bias); final LN -> mean pool ->
classifier -> GELU.  The caller's HF module is kept for its parameters only (names / order / objects unchanged).

Tokens are rows of [n_img * H * W, C] matrices; window partition, cyclic shift and patch merging are row gathers with
index plans computed once per (resolution, window, shift) on the host.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import lib, ops
from .ops import _Arena, _cw, _with_prec

_PLAN_CACHE = {}


def _window_plan(n, Hs, Ws, ws, ss, dev):
    """perm[r_window_order] = source row (image order) after a cyclic shift by -ss; inv = inverse permutation;
    mask [nW, L, L] (0 / -100) for shifted windows (HF SwinLayer.get_attn_mask), else None."""
    key = ("win", n, Hs, Ws, ws, ss, str(dev))
    if key in _PLAN_CACHE:
        return _PLAN_CACHE[key]
    assert Hs % ws == 0 and Ws % ws == 0, "resolution must be a multiple of the window (HF pads otherwise)"
    b = np.arange(n).reshape(n, 1, 1, 1, 1)
    wh = np.arange(Hs // ws).reshape(1, -1, 1, 1, 1)
    ww = np.arange(Ws // ws).reshape(1, 1, -1, 1, 1)
    i = np.arange(ws).reshape(1, 1, 1, -1, 1)
    j = np.arange(ws).reshape(1, 1, 1, 1, -1)
    h = (wh * ws + i + ss) % Hs
    w = (ww * ws + j + ss) % Ws
    perm = (b * Hs * Ws + h * Ws + w).reshape(-1).astype(np.int32)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size, dtype=np.int32)
    mask = None
    if ss > 0:
        img = np.zeros((Hs, Ws), dtype=np.float32)
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -ss), slice(-ss, None)):
            for wsl in (slice(0, -ws), slice(-ws, -ss), slice(-ss, None)):
                img[hs, wsl] = cnt
                cnt += 1
        mw = img.reshape(Hs // ws, ws, Ws // ws, ws).transpose(0, 2, 1, 3).reshape(-1, ws * ws)
        diff = mw[:, None, :] - mw[:, :, None]
        mask = np.where(diff != 0, -100.0, 0.0).astype(np.float32)
    plan = (lib.h2d(perm, dev), lib.h2d(inv, dev), None if mask is None else lib.h2d(mask, dev))
    _PLAN_CACHE[key] = plan
    return plan


def _merge_plan(n, Hs, Ws, dev):
    key = ("merge", n, Hs, Ws, str(dev))
    if key in _PLAN_CACHE:
        return _PLAN_CACHE[key]
    assert Hs % 2 == 0 and Ws % 2 == 0
    b = np.arange(n).reshape(n, 1, 1)
    i = np.arange(Hs // 2).reshape(1, -1, 1)
    j = np.arange(Ws // 2).reshape(1, 1, -1)
    idx = []
    for (di, dj) in ((0, 0), (1, 0), (0, 1), (1, 1)):      # HF order: [0::2,0::2], [1::2,0::2], [0::2,1::2], [1::2,1::2]
        idx.append(lib.h2d((b * Hs * Ws + (2 * i + di) * Ws + (2 * j + dj)).reshape(-1).astype(np.int32), dev))
    _PLAN_CACHE[key] = idx
    return idx


def _pool_plan(n, L, dev):
    key = ("pool", n, L, str(dev))
    if key not in _PLAN_CACHE:
        _PLAN_CACHE[key] = lib.h2d((np.arange(n * L) // L).astype(np.int32), dev)
    return _PLAN_CACHE[key]


def swin_param_list(image_net):
    return [p for _, p in image_net.named_parameters()]


def _structure(image_net):
    """static description of the HF module (per call; cheap)"""
    cfg = image_net.config
    sw = image_net.swin
    ps = cfg.patch_size
    stages = []
    for layer in sw.encoder.layers:
        blocks = []
        for blk in layer.blocks:
            sa = blk.attention.self
            dp = float(getattr(blk.drop_path, "drop_prob", 0.0) or 0.0)
            blocks.append(dict(mod=blk, heads=sa.num_attention_heads, dp=dp, shift=int(blk.shift_size),
                               ws=int(blk.window_size if not isinstance(blk.window_size, (tuple, list)) else blk.window_size[0])))
        stages.append(dict(blocks=blocks, down=layer.downsample))
    return cfg, ps, stages


class SwinTowerFn(torch.autograd.Function):
    """E[n, D] = GELU(classifier(meanpool(LN(SwinEncoder(patch_embed(pixels))))))."""

    @staticmethod
    @_with_prec
    def forward(ctx, meta, pixels, *params):
        net = meta["net"]
        adt = meta["adt"]
        cfg, ps, stages = _structure(net)
        eps = cfg.layer_norm_eps
        dev = pixels.device
        pid = {id(p): i for i, p in enumerate(net.parameters())}   # params are passed in this order
        P = lambda t: t.detach()                                   # noqa: E731
        n, Cin, Hi, Wi = pixels.shape
        Hs, Ws = Hi // ps, Wi // ps
        # ---- patch embedding: im2col (index permutation by torch) + GEMM + LN
        emb = net.swin.embeddings
        wpe = emb.patch_embeddings.projection.weight
        cols = pixels.reshape(n, Cin, Hs, ps, Ws, ps).permute(0, 2, 4, 1, 3, 5).reshape(n * Hs * Ws, Cin * ps * ps)
        cols = cols.to(adt).contiguous()
        C = wpe.shape[0]
        x0 = lib.linear_fwd(cols, _cw(wpe.reshape(C, -1), adt), P(emb.patch_embeddings.projection.bias))
        x, _, rstd_e = lib.layernorm_fwd(x0, P(emb.norm.weight), P(emb.norm.bias), eps)
        saved = dict(cols=cols, x_emb=x, rstd_e=rstd_e, stages=[])
        training = meta["training"]
        for st in stages:
            srec = dict(blocks=[], Hs=Hs, Ws=Ws, C=C)
            for b in st["blocks"]:
                m = b["mod"]
                heads, ws, ss = b["heads"], b["ws"], b["shift"]
                if min(Hs, Ws) <= ws:                               # HF set_shift_and_window_size
                    ss, ws = 0, min(Hs, Ws)
                L = ws * ws
                perm, inv, mask = _window_plan(n, Hs, Ws, ws, ss, dev)
                sa, so = m.attention.self, m.attention.output
                y, _, rstd_b = lib.layernorm_fwd(x, P(m.layernorm_before.weight), P(m.layernorm_before.bias), eps)
                yw = lib.gather_rows(y, perm)
                wqkv = _cw(torch.cat([P(sa.query.weight), P(sa.key.weight), P(sa.value.weight)], 0), adt)
                bqkv = torch.cat([P(sa.query.bias), P(sa.key.bias), P(sa.value.bias)], 0) if sa.query.bias is not None else None
                qkv = lib.linear_fwd(yw, wqkv, bqkv)
                ridx = sa.relative_position_index.reshape(-1)
                bias = P(sa.relative_position_bias_table)[ridx].view(L, L, heads).permute(2, 0, 1).contiguous().float()
                ctxo = torch.empty(yw.shape[0], C, device=dev, dtype=adt)
                dh = C // heads
                lib.attn_gen_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], ctxo, bias=bias, mask=mask,
                                 n_seq=yw.shape[0] // L, seqlen=L, n_heads=heads, head_dim=dh, scale=1.0 / math.sqrt(dh))
                w_ao = _cw(so.dense.weight, adt)
                ao = lib.linear_fwd(ctxo, w_ao, P(so.dense.bias))
                gs = None
                if training and b["dp"] > 0:
                    keep = 1.0 - b["dp"]
                    gs = (torch.rand(n, device=dev) < keep).float() / keep
                x1 = lib.scale_add_rows(ao, x=x, idx=inv, group_scale=gs, rows_per_group=Hs * Ws)
                z, _, rstd_a = lib.layernorm_fwd(x1, P(m.layernorm_after.weight), P(m.layernorm_after.bias), eps)
                w_i, w_o = _cw(m.intermediate.dense.weight, adt), _cw(m.output.dense.weight, adt)
                pre = torch.empty(z.shape[0], w_i.shape[0], device=dev, dtype=adt)
                u = lib.linear_fwd(z, w_i, P(m.intermediate.dense.bias), epilogue=lib.EPI_GELU_DGELU, pre=pre)   # pre <- gelu'(z)
                mo = lib.linear_fwd(u, w_o, P(m.output.dense.bias))
                x2 = lib.scale_add_rows(mo, x=x1)
                srec["blocks"].append(dict(y=y, rstd_b=rstd_b, yw=yw, qkv=qkv, bias=bias, mask=mask, perm=perm, inv=inv,
                                           ctxo=ctxo, gs=gs, z=z, rstd_a=rstd_a, pre=pre, u=u, L=L, heads=heads,
                                           w=(wqkv, w_ao, w_i, w_o), ridx=ridx))
                x = x2
            if st["down"] is not None:
                d = st["down"]
                midx = _merge_plan(n, Hs, Ws, dev)
                n2 = midx[0].numel()
                xm = torch.empty(n2, 4 * C, device=dev, dtype=adt)
                for k in range(4):
                    lib.gather_rows(x, midx[k], out=xm[:, k * C:(k + 1) * C])
                xn, _, rstd_m = lib.layernorm_fwd(xm, P(d.norm.weight), P(d.norm.bias), eps)
                w_red = _cw(d.reduction.weight, adt)
                x = lib.linear_fwd(xn, w_red)
                srec["down"] = dict(midx=midx, xn=xn, rstd_m=rstd_m, w_red=w_red, n_prev=n * Hs * Ws)
                Hs, Ws, C = Hs // 2, Ws // 2, 2 * C
            saved["stages"].append(srec)
        # ---- head: LN -> mean pool over tokens -> classifier -> GELU
        sw = net.swin
        xf, _, rstd_f = lib.layernorm_fwd(x, P(sw.layernorm.weight), P(sw.layernorm.bias), eps)
        Lf = Hs * Ws
        pooled = lib.mean_rows(xf, n, Lf)
        wc = _cw(net.classifier.weight, adt)
        cpre = torch.empty(n, wc.shape[0], device=dev, dtype=adt)
        E = lib.linear_fwd(pooled, wc, P(net.classifier.bias), epilogue=lib.EPI_GELU, pre=cpre)
        saved.update(xf=xf, rstd_f=rstd_f, pooled=pooled, cpre=cpre, wc=wc, Lf=Lf, n=n, Cf=C)
        ctx.meta, ctx.saved, ctx.params, ctx.pid = meta, saved, params, pid
        return E

    @staticmethod
    @_with_prec
    def backward(ctx, dE):
        meta, sv, params, pid = ctx.meta, ctx.saved, ctx.params, ctx.pid
        net, adt = meta["net"], meta["adt"]
        cfg, ps, stages = _structure(net)
        dev = dE.device
        need = [p.requires_grad for p in params]
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        arena = _Arena(dev, [p.shape for p in params])

        def G(p):
            g = arena.take(p.shape)
            if need[pid[id(p)]]:
                grads[pid[id(p)]] = g
            return g

        P = lambda t: t.detach()                                   # noqa: E731
        n, Lf = sv["n"], sv["Lf"]
        dE = dE.contiguous().to(adt)
        # ---- head
        dcp = lib.act_bwd(dE, sv["cpre"], 0)
        lib.linear_wgrad(dcp, sv["pooled"], G(net.classifier.weight))
        lib.colsum(dcp, G(net.classifier.bias))
        dpool = lib.linear_dgrad(dcp, sv["wc"])
        dxf = lib.scale_add_rows(dpool, idx=_pool_plan(n, Lf, dev), alpha=1.0 / Lf)
        sw = net.swin
        dx, _ = lib.layernorm_bwd(dxf, sv["xf"], P(sw.layernorm.weight), P(sw.layernorm.bias), sv["rstd_f"],
                                  dgamma=G(sw.layernorm.weight), dbeta=G(sw.layernorm.bias))
        for st, srec in zip(reversed(stages), reversed(sv["stages"])):
            C = srec["C"]
            if st["down"] is not None:
                d, dr = st["down"], srec["down"]
                lib.linear_wgrad(dx, dr["xn"], G(d.reduction.weight))
                dxn = lib.linear_dgrad(dx, dr["w_red"])
                dxm, _ = lib.layernorm_bwd(dxn, dr["xn"], P(d.norm.weight), P(d.norm.bias), dr["rstd_m"],
                                           dgamma=G(d.norm.weight), dbeta=G(d.norm.bias))
                d32 = torch.zeros(dr["n_prev"], C, device=dev, dtype=torch.float32)
                for k in range(4):
                    lib.scatter_add_rows(dxm[:, k * C:(k + 1) * C], dr["midx"][k], d32)
                dx = d32 if adt == torch.float32 else d32.to(adt)
            for b, rec in zip(reversed(st["blocks"]), reversed(srec["blocks"])):
                m = b["mod"]
                sa, so = m.attention.self, m.attention.output
                wqkv, w_ao, w_i, w_o = rec["w"]
                L, heads = rec["L"], rec["heads"]
                Hs, Ws = srec["Hs"], srec["Ws"]
                # ---- MLP branch:  x2 = x1 + W2 gelu(W1 LN(x1))
                lib.linear_wgrad(dx, rec["u"], G(m.output.dense.weight))
                lib.colsum(dx, G(m.output.dense.bias))
                dpre = lib.linear_dgrad(dx, w_o, epilogue=lib.EPI_MUL_AUX, aux=rec["pre"])
                lib.linear_wgrad(dpre, rec["z"], G(m.intermediate.dense.weight))
                lib.colsum(dpre, G(m.intermediate.dense.bias))
                dz = lib.linear_dgrad(dpre, w_i)
                dz_in, _ = lib.layernorm_bwd(dz, rec["z"], P(m.layernorm_after.weight), P(m.layernorm_after.bias),
                                             rec["rstd_a"], dgamma=G(m.layernorm_after.weight),
                                             dbeta=G(m.layernorm_after.bias))
                dx1 = lib.scale_add_rows(dz_in, x=dx)                                  # residual + branch
                # ---- attention branch:  x1 = x + s_img * ao[inv]   =>   dao = s_img * dx1[perm]
                dao = lib.scale_add_rows(dx1, idx=rec["perm"], group_scale=rec["gs"], rows_per_group=Hs * Ws)
                lib.linear_wgrad(dao, rec["ctxo"], G(so.dense.weight))
                lib.colsum(dao, G(so.dense.bias))
                dctx = lib.linear_dgrad(dao, w_ao)
                qkv = rec["qkv"]
                dqkv = torch.empty_like(qkv)
                dbias = torch.zeros(heads, L, L, device=dev, dtype=torch.float32)
                dh = C // heads
                lib.attn_gen_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], dctx, dqkv[:, :C], dqkv[:, C:2 * C],
                                 dqkv[:, 2 * C:], dbias=dbias, bias=rec["bias"], mask=rec["mask"], n_seq=qkv.shape[0] // L,
                                 seqlen=L, n_heads=heads, head_dim=dh, scale=1.0 / math.sqrt(dh))
                gt = G(sa.relative_position_bias_table)
                gt.index_add_(0, rec["ridx"], dbias.permute(1, 2, 0).reshape(L * L, heads))
                dw = torch.zeros(3 * C, C, device=dev, dtype=torch.float32)
                lib.linear_wgrad(dqkv, rec["yw"], dw)
                G(sa.query.weight).copy_(dw[:C]); G(sa.key.weight).copy_(dw[C:2 * C]); G(sa.value.weight).copy_(dw[2 * C:])
                if sa.query.bias is not None:
                    dbq = torch.zeros(3 * C, device=dev, dtype=torch.float32)
                    lib.colsum(dqkv, dbq)
                    G(sa.query.bias).copy_(dbq[:C]); G(sa.key.bias).copy_(dbq[C:2 * C]); G(sa.value.bias).copy_(dbq[2 * C:])
                dyw = lib.linear_dgrad(dqkv, wqkv)
                dy = lib.gather_rows(dyw, rec["inv"])
                dz0, _ = lib.layernorm_bwd(dy, rec["y"], P(m.layernorm_before.weight), P(m.layernorm_before.bias),
                                           rec["rstd_b"], dgamma=G(m.layernorm_before.weight),
                                           dbeta=G(m.layernorm_before.bias))
                dx = lib.scale_add_rows(dz0, x=dx1)
        # ---- patch embedding
        emb = net.swin.embeddings
        dx0, _ = lib.layernorm_bwd(dx, sv["x_emb"], P(emb.norm.weight), P(emb.norm.bias), sv["rstd_e"],
                                   dgamma=G(emb.norm.weight), dbeta=G(emb.norm.bias))
        wpe = emb.patch_embeddings.projection.weight
        gw = G(wpe)
        lib.linear_wgrad(dx0, sv["cols"], gw.view(wpe.shape[0], -1))
        lib.colsum(dx0, G(emb.patch_embeddings.projection.bias))
        ctx.saved = None
        return (None, None) + tuple(grads)
