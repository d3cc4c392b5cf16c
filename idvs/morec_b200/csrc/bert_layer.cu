// Host-side sequencing of one post-LN BERT encoder layer (forward: 7 kernel launches, backward: 17) behind a single
// C-ABI call each, so the Python host pays one ctypes transition per layer instead of one per kernel.
//
// Replaces (per layer) HF BertLayer.forward / its autograd backward, call site model/encoders.py:68:
//   qkv = x Wqkv^T + b            (one fused projection GEMM, N = 3H)
//   ctx = softmax(q k^T / sqrt(d_h)) v        per packed sequence (pad tokens do not exist)
//   x1  = LN(x + dropout(ctx Wo^T + bo))
//   x2  = LN(x1 + dropout(gelu(x1 Wi^T + bi) Wo2^T + bo2))
// All buffers are caller-allocated device memory; nothing here allocates or synchronises.
#include "../../../include/morec_b200.h"
#include "common.cuh"

#define RUN(call)              \
    do {                       \
        int rc__ = (call);     \
        if (rc__ != MOREC_OK) return rc__; \
    } while (0)

extern "C" int morec_bert_layer_fwd(const MorecBertLayerFwd* a, void* stream) {
    MOREC_CHECK_ARG(a, "bert_layer_fwd: null args");
    const int M = a->n_tok, H = a->H, I = a->I;
    const int st = MOREC_DT_IS16(a->dtype) ? a->dtype : 0;   // storage dtype code for the non-GEMM kernels
    const int obf = st;                                      // GEMM out_dtype: same storage type as the operands
    const int es = st ? 2 : 4;
    const float scale = 1.0f / sqrtf((float)(H / a->n_heads));
    // fused QKV projection
    RUN(morec_gemm(a->x, a->wqkv, a->qkv, nullptr, a->bqkv, nullptr, M, 3 * H, H, H, H, 3 * H, 0, 0, 0, a->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    const char* q = (const char*)a->qkv;
    if (a->max_len <= 32) {
        RUN(morec_attn_fwd(q, q + (size_t)H * es, q + (size_t)2 * H * es, a->ctx, a->cu_seqlens, nullptr, 0, a->n_seq,
                           a->max_len, a->n_heads, H / a->n_heads, 3 * H, H, scale, -1e9f, a->dtype, a->p_attn, a->seed,
                           a->off_attn, stream));
    } else {   // long titles (cfg-2: T = 128): general kernel, one CTA per (sequence, head)
        RUN(morec_attn_gen_fwd(q, q + (size_t)H * es, q + (size_t)2 * H * es, a->ctx, a->cu_seqlens, nullptr, nullptr, 0,
                               a->n_seq, a->max_len, a->n_heads, H / a->n_heads, 3 * H, H, scale, a->dtype, a->p_attn,
                               a->seed, a->off_attn, stream));
    }
    RUN(morec_gemm(a->ctx, a->w_ao, a->tmp_h, nullptr, a->b_ao, nullptr, M, H, H, H, H, H, 0, 0, 0, a->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    RUN(morec_layernorm_fwd(a->tmp_h, a->x, nullptr, 0, a->g1, a->b1, a->x1, nullptr, a->rstd1, M, H, a->eps, st,
                            a->p_hidden, 0.f, a->seed, a->off_ln1, 0, stream));
    RUN(morec_gemm(a->x1, a->w_i, a->act, a->pre, a->b_i, nullptr, M, I, H, H, H, I, 0, 0, 0, a->dtype, obf, MOREC_EPI_GELU_DGELU,
                   1.f, 0, stream));
    RUN(morec_gemm(a->act, a->w_o, a->tmp_h, nullptr, a->b_o, nullptr, M, H, I, I, I, H, 0, 0, 0, a->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    RUN(morec_layernorm_fwd(a->tmp_h, a->x1, nullptr, 0, a->g2, a->b2, a->x2, nullptr, a->rstd2, M, H, a->eps, st,
                            a->p_hidden, 0.f, a->seed, a->off_ln2, 0, stream));
    return MOREC_OK;
}

extern "C" int morec_bert_layer_bwd(const MorecBertLayerBwd* a, void* stream) {
    MOREC_CHECK_ARG(a, "bert_layer_bwd: null args");
    const MorecBertLayerFwd* f = &a->fwd;
    const int M = f->n_tok, H = f->H, I = f->I;
    const int st = MOREC_DT_IS16(f->dtype) ? f->dtype : 0;
    const int obf = st;
    const int es = st ? 2 : 4;
    const float scale = 1.0f / sqrtf((float)(H / f->n_heads));
    const bool drop = f->p_hidden > 0.f;
    // ---- output LayerNorm: dy (+dy2) w.r.t. x2 -> dz2 (residual stream) and dfo (branch, dropout mask applied)
    void* dfo = drop ? a->dbr : a->dz2;
    RUN(morec_layernorm_bwd(a->dy, a->dy2, f->x2, f->g2, f->b2, f->rstd2, a->dz2, drop ? a->dbr : nullptr, a->dg2, a->db2,
                            a->db_o, nullptr, 0, M, H, st, f->p_hidden, 0.f, f->seed, f->off_ln2, 0, stream));
    // ---- FFN2: dWo2 += dfo^T act ; dpre = (dfo Wo2) * gelu'(pre)   (`pre` holds gelu'(pre-activation), saved by the forward)
    RUN(morec_gemm(dfo, f->act, a->dw_o, nullptr, nullptr, nullptr, H, I, M, H, I, I, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, stream));
    RUN(morec_gemm(dfo, f->w_o, a->dpre, nullptr, nullptr, f->pre, M, I, H, H, I, I, I, 0, 1, f->dtype, obf,
                   MOREC_EPI_MUL_AUX, 1.f, 0, stream));
    // ---- FFN1: dbi, dWi, dx1 (branch)
    RUN(morec_colsum(a->dpre, a->db_i, M, I, I, st, stream));
    RUN(morec_gemm(a->dpre, f->x1, a->dw_i, nullptr, nullptr, nullptr, I, H, M, I, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, stream));
    RUN(morec_gemm(a->dpre, f->w_i, a->dx1b, nullptr, nullptr, nullptr, M, H, I, I, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    // ---- attention-output LayerNorm: dy = dz2 + dx1b w.r.t. x1 -> dz1 and dao
    void* dao = drop ? a->dbr : a->dz1;
    RUN(morec_layernorm_bwd(a->dz2, a->dx1b, f->x1, f->g1, f->b1, f->rstd1, a->dz1, drop ? a->dbr : nullptr, a->dg1,
                            a->db1, a->db_ao, nullptr, 0, M, H, st, f->p_hidden, 0.f, f->seed, f->off_ln1, 0, stream));
    RUN(morec_gemm(dao, f->ctx, a->dw_ao, nullptr, nullptr, nullptr, H, H, M, H, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, stream));
    RUN(morec_gemm(dao, f->w_ao, a->dctx, nullptr, nullptr, nullptr, M, H, H, H, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    // ---- attention core
    const char* q = (const char*)f->qkv;
    char* dq = (char*)a->dqkv;
    if (f->max_len <= 32) {
        RUN(morec_attn_bwd(q, q + (size_t)H * es, q + (size_t)2 * H * es, a->dctx, dq, dq + (size_t)H * es,
                           dq + (size_t)2 * H * es, f->cu_seqlens, nullptr, 0, f->n_seq, f->max_len, f->n_heads,
                           H / f->n_heads, 3 * H, H, scale, -1e9f, f->dtype, f->p_attn, f->seed, f->off_attn, stream));
    } else {
        RUN(morec_attn_gen_bwd(q, q + (size_t)H * es, q + (size_t)2 * H * es, a->dctx, dq, dq + (size_t)H * es,
                               dq + (size_t)2 * H * es, nullptr, f->cu_seqlens, nullptr, nullptr, 0, f->n_seq,
                               f->max_len, f->n_heads, H / f->n_heads, 3 * H, H, scale, f->dtype, f->p_attn, f->seed,
                               f->off_attn, stream));
    }
    // ---- fused QKV projection: dWqkv, dbqkv, dx (second part of the layer-input gradient; the first is dz1)
    RUN(morec_gemm(a->dqkv, f->x, a->dwqkv, nullptr, nullptr, nullptr, 3 * H, H, M, 3 * H, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, stream));
    RUN(morec_colsum(a->dqkv, a->dbqkv, M, 3 * H, 3 * H, st, stream));
    RUN(morec_gemm(a->dqkv, f->wqkv, a->dxq, nullptr, nullptr, nullptr, M, H, 3 * H, 3 * H, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    return MOREC_OK;
}

// Whole towers per call: `layers` is a HOST array of per-layer argument records (forward: in execution order;
// backward: in execution order, i.e. last layer first).  One C-ABI transition per pass instead of one per layer -- the
// host side of a step was as long as its GPU side (bench.py `host_issue_ms_per_step`).
extern "C" int morec_bert_layers_fwd(const MorecBertLayerFwd* layers, int n_layers, void* stream) {
    MOREC_CHECK_ARG(layers && n_layers >= 0, "bert_layers_fwd: null args");
    for (int l = 0; l < n_layers; ++l) RUN(morec_bert_layer_fwd(layers + l, stream));
    return MOREC_OK;
}

extern "C" int morec_bert_layers_bwd(const MorecBertLayerBwd* layers, int n_layers, void* stream) {
    MOREC_CHECK_ARG(layers && n_layers >= 0, "bert_layers_bwd: null args");
    for (int l = 0; l < n_layers; ++l) RUN(morec_bert_layer_bwd(layers + l, stream));
    return MOREC_OK;
}
