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

#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Backward with a SIDE STREAM: the four weight-gradient GEMMs (and the two bias column sums) of a layer do not feed
// the activation-gradient chain, so they run on a second, lower-priority stream and overlap the chain's HBM-bound
// kernels (two LayerNorm backwards, attention backward) -- tensor-core work filling the SMs while the chain waits on
// memory.  Ordering: a side task waits for the event of its producer on the main stream; the main stream waits for a
// side task only where a scratch buffer is about to be overwritten (dfo/dao share `dbr` under dropout, dpre and dqkv
// are reused by the next layer).  The caller-visible contract is unchanged: when the call returns, everything has
// been enqueued and the main stream is ordered after all side work (`join`).  MOREC_SIDE_STREAM=0 disables it.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct SideCtx {
    cudaStream_t side = nullptr;
    cudaEvent_t prod[4] = {nullptr, nullptr, nullptr, nullptr};   // main -> side: producer finished (dfo, dpre, dao, dqkv)
    cudaEvent_t done[4] = {nullptr, nullptr, nullptr, nullptr};   // side -> main: consumer finished with the scratch buffer
    bool pending[4] = {false, false, false, false};               // done[k] recorded and not yet waited for
    bool ok = false;
};
SideCtx* side_ctx() {
    static thread_local SideCtx ctx[16];
    static const bool enabled = []() { const char* e = getenv("MOREC_SIDE_STREAM"); return !(e && e[0] == '0'); }();
    if (!enabled) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideCtx& c = ctx[dev];
    if (!c.ok) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
        for (int i = 0; i < 4; ++i) {
            if (cudaEventCreateWithFlags(&c.prod[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&c.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        c.ok = true;
    }
    return &c;
}
}  // namespace

static int bert_layer_bwd_impl(const MorecBertLayerBwd* a, void* stream, bool join) {
    const MorecBertLayerFwd* f = &a->fwd;
    const int M = f->n_tok, H = f->H, I = f->I;
    const int st = MOREC_DT_IS16(f->dtype) ? f->dtype : 0;
    const int obf = st;
    const int es = st ? 2 : 4;
    const float scale = 1.0f / sqrtf((float)(H / f->n_heads));
    const bool drop = f->p_hidden > 0.f;
    cudaStream_t main_s = (cudaStream_t)stream;
    SideCtx* sc = side_ctx();
    void* side = sc ? (void*)sc->side : stream;
    // fork(k): side stream waits for everything enqueued on main so far; landed(k): remember that side work reading
    // scratch k is in flight; need(k): main must not overwrite scratch k before that work has finished
    auto fork = [&](int k) -> int {
        if (!sc) return MOREC_OK;
        MOREC_CUDA(cudaEventRecord(sc->prod[k], main_s));
        MOREC_CUDA(cudaStreamWaitEvent(sc->side, sc->prod[k], 0));
        return MOREC_OK;
    };
    auto landed = [&](int k) -> int {
        if (!sc) return MOREC_OK;
        MOREC_CUDA(cudaEventRecord(sc->done[k], sc->side));
        sc->pending[k] = true;
        return MOREC_OK;
    };
    auto need = [&](int k) -> int {
        if (!sc || !sc->pending[k]) return MOREC_OK;
        MOREC_CUDA(cudaStreamWaitEvent(main_s, sc->done[k], 0));
        sc->pending[k] = false;
        return MOREC_OK;
    };
    // ---- output LayerNorm: dy (+dy2) w.r.t. x2 -> dz2 (residual stream) and dfo (branch, dropout mask applied)
    void* dfo = drop ? a->dbr : a->dz2;
    RUN(need(0)); RUN(need(2));                  // dz2 / dbr were read by the previous layer's side GEMMs (dfo, dao)
    RUN(morec_layernorm_bwd(a->dy, a->dy2, f->x2, f->g2, f->b2, f->rstd2, a->dz2, drop ? a->dbr : nullptr, a->dg2, a->db2,
                            a->db_o, nullptr, 0, M, H, st, f->p_hidden, 0.f, f->seed, f->off_ln2, 0, stream));
    // ---- FFN2: dWo2 += dfo^T act (side) ; dpre = (dfo Wo2) * gelu'(pre)   (`pre` holds gelu'(pre-activation), saved by the forward)
    RUN(fork(0));
    RUN(morec_gemm(dfo, f->act, a->dw_o, nullptr, nullptr, nullptr, H, I, M, H, I, I, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, side));
    RUN(landed(0));
    RUN(need(1));                                // dpre was read by the previous layer's side work
    RUN(morec_gemm(dfo, f->w_o, a->dpre, nullptr, nullptr, f->pre, M, I, H, H, I, I, I, 0, 1, f->dtype, obf,
                   MOREC_EPI_MUL_AUX, 1.f, 0, stream));
    // ---- FFN1: dbi, dWi (side), dx1 (branch)
    RUN(fork(1));
    RUN(morec_colsum(a->dpre, a->db_i, M, I, I, st, side));
    RUN(morec_gemm(a->dpre, f->x1, a->dw_i, nullptr, nullptr, nullptr, I, H, M, I, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, side));
    RUN(landed(1));
    RUN(morec_gemm(a->dpre, f->w_i, a->dx1b, nullptr, nullptr, nullptr, M, H, I, I, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    // ---- attention-output LayerNorm: dy = dz2 + dx1b w.r.t. x1 -> dz1 and dao
    void* dao = drop ? a->dbr : a->dz1;
    if (drop) RUN(need(0));                      // dao overwrites dbr, which the side GEMM still reads as dfo
    RUN(morec_layernorm_bwd(a->dz2, a->dx1b, f->x1, f->g1, f->b1, f->rstd1, a->dz1, drop ? a->dbr : nullptr, a->dg1,
                            a->db1, a->db_ao, nullptr, 0, M, H, st, f->p_hidden, 0.f, f->seed, f->off_ln1, 0, stream));
    RUN(fork(2));
    RUN(morec_gemm(dao, f->ctx, a->dw_ao, nullptr, nullptr, nullptr, H, H, M, H, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, side));
    RUN(landed(2));
    RUN(morec_gemm(dao, f->w_ao, a->dctx, nullptr, nullptr, nullptr, M, H, H, H, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    // ---- attention core
    const char* q = (const char*)f->qkv;
    char* dq = (char*)a->dqkv;
    RUN(need(3));                                // dqkv was read by the previous layer's side work
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
    // ---- fused QKV projection: dWqkv, dbqkv (side), dx (second part of the layer-input gradient; the first is dz1)
    RUN(fork(3));
    RUN(morec_gemm(a->dqkv, f->x, a->dwqkv, nullptr, nullptr, nullptr, 3 * H, H, M, 3 * H, H, H, 0, 1, 1, f->dtype, 0,
                   MOREC_EPI_LINEAR, 1.f, 1, side));
    RUN(morec_colsum(a->dqkv, a->dbqkv, M, 3 * H, 3 * H, st, side));
    RUN(landed(3));
    RUN(morec_gemm(a->dqkv, f->wqkv, a->dxq, nullptr, nullptr, nullptr, M, H, 3 * H, 3 * H, H, H, 0, 0, 1, f->dtype, obf,
                   MOREC_EPI_LINEAR, 1.f, 0, stream));
    if (join)
        for (int k = 0; k < 4; ++k) RUN(need(k));
    return MOREC_OK;
}

extern "C" int morec_bert_layer_bwd(const MorecBertLayerBwd* a, void* stream) {
    MOREC_CHECK_ARG(a, "bert_layer_bwd: null args");
    return bert_layer_bwd_impl(a, stream, true);
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
    for (int l = 0; l < n_layers; ++l) RUN(bert_layer_bwd_impl(layers + l, stream, l == n_layers - 1));
    return MOREC_OK;
}

extern "C" int morec_bert_layers_bwd_ex(const MorecBertLayerBwd* layers, int n_layers, int join, void* stream) {
    MOREC_CHECK_ARG((layers || n_layers == 0) && n_layers >= 0, "bert_layers_bwd_ex: null args");
    for (int l = 0; l < n_layers; ++l) RUN(bert_layer_bwd_impl(layers + l, stream, join && l == n_layers - 1));
    if (n_layers == 0 && join) {               // join only: order `stream` after whatever the side stream still holds
        SideCtx* sc = side_ctx();
        if (sc)
            for (int k = 0; k < 4; ++k)
                if (sc->pending[k]) {
                    MOREC_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, sc->done[k], 0));
                    sc->pending[k] = false;
                }
    }
    return MOREC_OK;
}
