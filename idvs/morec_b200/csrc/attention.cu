// Short-sequence multi-head attention, forward and backward, for sequences of at most 32 tokens
// (item titles: T = 30 word pieces, parameters.py:42; user histories: L <= 25..32, parameters.py:30).
//
// Replaces   HF BertSelfAttention  (call site model/encoders.py:68): softmax(QK^T/sqrt(d_h) + key-pad mask) V
//            SASRec SelfAttention  (model/modules.py:27-31, mask model/encoders.py:23-28):
//                                   softmax(QK^T/sqrt(d_k) + (k<=q & key valid ? 0 : -1e9)) V,  dropout on P
// One warp owns one (sequence, head) pair; lane i owns query row i.  K / V / Q / dO are streamed through a
// [32][64] fp32 shared-memory tile in head-dim chunks of 64 (any head_dim that is a multiple of 4 works: 16, 32, 64,
// 256, 1024), rows are read by all lanes at the same address (broadcast, conflict-free) while each lane keeps its
// own 64-wide slice of q / o / dq in registers.  The whole score row (<= 32 values) lives in registers, so the
// softmax is exact two-pass fp32.  Packed (variable-length) batches are described by cu_seqlens; fixed-length
// batches (SASRec) by seqlen + an optional [n_seq, seqlen] key-valid mask and a causal flag.
#include "../../../include/morec_b200.h"
#include "attention_common.cuh"
#include "attention_tc.cuh"

namespace morec {

constexpr int AT_WARPS = 4;
constexpr int AT_MAXL = 32;

struct AttnParams {
    const void *q, *k, *v, *o;        // o: forward output / backward: dO
    void *dq, *dk, *dv;               // backward outputs (null in forward); forward writes `out`
    void* out;
    const int* cu_seqlens;            // [n_seq+1] or null
    const float* key_mask;            // [n_seq, seqlen] (non-zero = valid key) or null
    int causal;
    int n_seq, seqlen, n_heads, head_dim, ld, ld_o;   // ld: q/k/v/dq/dk/dv row stride, ld_o: o/dO row stride
    float scale, masked_add;
    float dropout_p;
    uint64_t seed, offset;
};

__device__ __forceinline__ void seq_range(const AttnParams& p, int s, int& row0, int& len) {
    if (p.cu_seqlens) { row0 = p.cu_seqlens[s]; len = p.cu_seqlens[s + 1] - row0; }
    else { row0 = s * p.seqlen; len = p.seqlen; }
    if (len > AT_MAXL) len = AT_MAXL;
}

// additive mask for (query i, key j) of sequence s
__device__ __forceinline__ float mask_add(const AttnParams& p, int s, int i, int j) {
    bool ok = true;
    if (p.causal) ok = j <= i;
    if (p.key_mask) ok = ok && (p.key_mask[(size_t)s * p.seqlen + j] != 0.f);
    return ok ? 0.f : p.masked_add;
}

// ---------------------------------------------------------------------------------------------------------------
// Score rows live in shared memory (Srow[lane*33 + j], conflict-free) rather than in registers: the loops over keys
// then stay rolled (a fully unrolled 32-key x 64-dim body overflowed the instruction cache: ncu showed 41 % of the
// stall samples as "no_instructions").
// ---------------------------------------------------------------------------------------------------------------

// S[lane][j] = <q_lane, k_j> over the whole head dim (raw, unscaled); rows j >= len hold garbage that is never read
template <typename T>
__device__ __forceinline__ void raw_scores(const AttnParams& p, float* tile, float* S, const T* Qb, int ldq, const T* Kb,
                                           int h, int row0, int len, int lane) {
    const int len4 = (len + 3) & ~3;
    for (int dc = 0; dc < p.head_dim; dc += AT_DCH) {
        const int w = min(AT_DCH, p.head_dim - dc);
        __syncwarp();
        load_tile<T>(tile, Kb, p.ld, row0, len, h * p.head_dim + dc, w, lane);
        __syncwarp();
        float qv[AT_DCH];
        load_row<T>(qv, Qb + (size_t)(row0 + (lane < len ? lane : 0)) * ldq + h * p.head_dim + dc, w, lane < len);
#pragma unroll 1
        for (int j0 = 0; j0 < len4; j0 += 4) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            dot4keys(qv, tile, j0, w, a0, a1, a2, a3);
            float* sr = S + lane * 33 + j0;
            if (dc == 0) { sr[0] = a0; sr[1] = a1; sr[2] = a2; sr[3] = a3; }
            else { sr[0] += a0; sr[1] += a1; sr[2] += a2; sr[3] += a3; }
        }
    }
}

// in-place softmax of row `lane` of S over j < len (scale + additive mask first)
__device__ __forceinline__ void softmax_row(const AttnParams& p, float* S, int s, int len, int lane) {
    float* sr = S + lane * 33;
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < len; ++j) {
        const float x = sr[j] * p.scale + mask_add(p, s, lane, j);
        sr[j] = x;
        m = fmaxf(m, x);
    }
    float sum = 0.f;
#pragma unroll 1
    for (int j = 0; j < len; ++j) {
        const float e = __expf(sr[j] - m);
        sr[j] = e;
        sum += e;
    }
    const float inv = 1.f / sum;
#pragma unroll 1
    for (int j = 0; j < len; ++j) sr[j] *= inv;
}

// dropout keep mask bits for row (pair, i): bit j set = keep
__device__ __forceinline__ uint32_t keep_bits(const AttnParams& p, int pair, int i, int len) {
    if (!(p.dropout_p > 0.f)) return 0xffffffffu;
    const uint32_t th = (uint32_t)fminf(p.dropout_p * 4294967296.f, 4294967295.f);
    uint32_t bits = 0;
#pragma unroll 1
    for (int g = 0; g < AT_MAXL / 4; ++g) {
        if (4 * g < len) {
            const uint4 r = Philox::gen(p.seed, p.offset + ((uint64_t)pair * AT_MAXL + i) * (AT_MAXL / 4) + g);
            bits |= (uint32_t)(r.x >= th) << (4 * g);
            bits |= (uint32_t)(r.y >= th) << (4 * g + 1);
            bits |= (uint32_t)(r.z >= th) << (4 * g + 2);
            bits |= (uint32_t)(r.w >= th) << (4 * g + 3);
        }
    }
    return bits;
}

// out_lane[0..w) = sum_j coef(lane, j) * tile[j][0..w)      coef read from a [32][33] smem matrix, row- or column-wise
template <bool TRANSPOSED>
__device__ __forceinline__ void weighted_rows(const float* coef, const float* tile, int len, int lane,
                                              float (&acc)[AT_DCH]) {
#pragma unroll
    for (int t = 0; t < AT_DCH; ++t) acc[t] = 0.f;
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
        const float d = TRANSPOSED ? coef[j * 33 + lane] : coef[lane * 33 + j];
        const float4* r = reinterpret_cast<const float4*>(tile + j * AT_DCH);
#pragma unroll
        for (int t = 0; t < AT_DCH / 4; ++t) {
            const float4 x = r[t];
            acc[4 * t] += d * x.x; acc[4 * t + 1] += d * x.y; acc[4 * t + 2] += d * x.z; acc[4 * t + 3] += d * x.w;
        }
    }
}

constexpr int AT_FWD_SMEM_FLOATS = AT_WARPS * (AT_MAXL * AT_DCH + AT_MAXL * 33);

template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_fwd_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float at_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = at_smem + warp * (AT_MAXL * AT_DCH);
    float* S = at_smem + AT_WARPS * (AT_MAXL * AT_DCH) + warp * (AT_MAXL * 33);
    const int n_pairs = p.n_seq * p.n_heads;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    T* O = reinterpret_cast<T*>(p.out);
    for (int pair = blockIdx.x * AT_WARPS + warp; pair < n_pairs; pair += gridDim.x * AT_WARPS) {
        const int s = pair / p.n_heads, h = pair - s * p.n_heads;
        int row0, len;
        seq_range(p, s, row0, len);
        if (len <= 0) continue;
        raw_scores<T>(p, tile, S, Q, p.ld, K, h, row0, len, lane);
        softmax_row(p, S, s, len, lane);
        if (p.dropout_p > 0.f) {
            const uint32_t kb = keep_bits(p, pair, lane, len);
            const float sc = 1.f / (1.f - p.dropout_p);
#pragma unroll 1
            for (int j = 0; j < len; ++j) S[lane * 33 + j] = ((kb >> j) & 1u) ? S[lane * 33 + j] * sc : 0.f;
        }
        for (int dc = 0; dc < p.head_dim; dc += AT_DCH) {
            const int w = min(AT_DCH, p.head_dim - dc);
            __syncwarp();
            load_tile<T>(tile, V, p.ld, row0, len, h * p.head_dim + dc, w, lane);
            __syncwarp();
            float ov[AT_DCH];
            weighted_rows<false>(S, tile, len, lane, ov);
            if (lane < len) store_row<T>(O + (size_t)(row0 + lane) * p.ld_o + h * p.head_dim + dc, ov, w);
        }
        __syncwarp();
    }
}

constexpr int AT_BWD_SMEM_FLOATS = AT_WARPS * (AT_MAXL * AT_DCH + 2 * AT_MAXL * 33);

template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float at_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = at_smem + warp * (AT_MAXL * AT_DCH);                                   // [32][64] stream tile
    float* Pw = at_smem + AT_WARPS * (AT_MAXL * AT_DCH) + warp * (2 * AT_MAXL * 33);      // P, then P~ (dropped, rescaled)
    float* dSw = Pw + AT_MAXL * 33;                                                      // dP~, then dS * scale
    const int n_pairs = p.n_seq * p.n_heads;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    const T* dO = reinterpret_cast<const T*>(p.o);
    T* dQ = reinterpret_cast<T*>(p.dq);
    T* dK = reinterpret_cast<T*>(p.dk);
    T* dV = reinterpret_cast<T*>(p.dv);
    for (int pair = blockIdx.x * AT_WARPS + warp; pair < n_pairs; pair += gridDim.x * AT_WARPS) {
        const int s = pair / p.n_heads, h = pair - s * p.n_heads;
        int row0, len;
        seq_range(p, s, row0, len);
        if (len <= 0) continue;
        const int col_h = h * p.head_dim;
        // P = softmax(QK^T) ; dP~[i][j] = dO_i . V_j   (same inner product routine: "queries" = dO rows, "keys" = V)
        raw_scores<T>(p, tile, Pw, Q, p.ld, K, h, row0, len, lane);
        softmax_row(p, Pw, s, len, lane);
        {
            AttnParams pv = p;           // reuse raw_scores with V as the key matrix
            raw_scores<T>(pv, tile, dSw, dO, p.ld_o, V, h, row0, len, lane);
        }
        // softmax backward (dropout on P): dP = keep * dP~ / (1-p);  dS = P * (dP - sum_k P_k dP_k) * scale
        {
            const uint32_t kb = keep_bits(p, pair, lane, len);
            const float sc = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
            float* pr = Pw + lane * 33;
            float* dr = dSw + lane * 33;
            float dsum = 0.f;
            const bool act = lane < len;
#pragma unroll 1
            for (int j = 0; j < len; ++j) {
                const float keep = ((kb >> j) & 1u) ? sc : 0.f;
                const float d = act ? dr[j] * keep : 0.f;
                dr[j] = d;
                dsum += (act ? pr[j] : 0.f) * d;
            }
#pragma unroll 1
            for (int j = 0; j < len; ++j) {
                const float keep = ((kb >> j) & 1u) ? sc : 0.f;
                const float pj = act ? pr[j] : 0.f;
                dr[j] = pj * (dr[j] - dsum) * p.scale;     // dS (scaled)
                pr[j] = pj * keep;                           // P~ for dV
            }
        }
        __syncwarp();
        // dQ_i = sum_j dS_ij K_j        (lane = query i; K chunk broadcast from smem)
        for (int dc = 0; dc < p.head_dim; dc += AT_DCH) {
            const int w = min(AT_DCH, p.head_dim - dc);
            __syncwarp();
            load_tile<T>(tile, K, p.ld, row0, len, col_h + dc, w, lane);
            __syncwarp();
            float acc[AT_DCH];
            weighted_rows<false>(dSw, tile, len, lane, acc);
            if (lane < len) store_row<T>(dQ + (size_t)(row0 + lane) * p.ld + col_h + dc, acc, w);
        }
        // dK_j = sum_i dS_ij Q_i        (lane = key j; Q chunk broadcast from smem)
        for (int dc = 0; dc < p.head_dim; dc += AT_DCH) {
            const int w = min(AT_DCH, p.head_dim - dc);
            __syncwarp();
            load_tile<T>(tile, Q, p.ld, row0, len, col_h + dc, w, lane);
            __syncwarp();
            float acc[AT_DCH];
            weighted_rows<true>(dSw, tile, len, lane, acc);
            if (lane < len) store_row<T>(dK + (size_t)(row0 + lane) * p.ld + col_h + dc, acc, w);
        }
        // dV_j = sum_i P~_ij dO_i       (lane = key j; dO chunk broadcast from smem)
        for (int dc = 0; dc < p.head_dim; dc += AT_DCH) {
            const int w = min(AT_DCH, p.head_dim - dc);
            __syncwarp();
            load_tile<T>(tile, dO, p.ld_o, row0, len, col_h + dc, w, lane);
            __syncwarp();
            float acc[AT_DCH];
            weighted_rows<true>(Pw, tile, len, lane, acc);
            if (lane < len) store_row<T>(dV + (size_t)(row0 + lane) * p.ld + col_h + dc, acc, w);
        }
        __syncwarp();
    }
}

// fast precision modes (dtype 0 / 1) with 32- or 64-wide heads run on the tensor-core kernels (attention_tc.cuh)
static TcAttnParams tc_params(const AttnParams& p) {
    TcAttnParams t{};
    t.q = p.q; t.k = p.k; t.v = p.v; t.o = p.o; t.out = p.out; t.dq = p.dq; t.dk = p.dk; t.dv = p.dv;
    t.cu_seqlens = p.cu_seqlens; t.key_mask = p.key_mask; t.causal = p.causal; t.masked_add = p.masked_add;
    t.n_seq = p.n_seq; t.seqlen = p.seqlen; t.n_heads = p.n_heads; t.head_dim = p.head_dim; t.ld = p.ld; t.ld_o = p.ld_o;
    t.scale = p.scale; t.dropout_p = p.dropout_p; t.seed = p.seed; t.offset = p.offset;
    return t;
}

static int check(const AttnParams& p) {
    MOREC_CHECK_ARG(p.q && p.k && p.v, "attention: null q/k/v");
    MOREC_CHECK_ARG(p.head_dim % 4 == 0 && p.head_dim > 0, "attention: head_dim=%d must be a multiple of 4", p.head_dim);
    MOREC_CHECK_ARG(p.seqlen <= AT_MAXL && p.seqlen > 0, "attention: max sequence length %d > %d unsupported", p.seqlen, AT_MAXL);
    MOREC_CHECK_ARG(p.ld % 4 == 0 && p.ld_o % 4 == 0, "attention: row strides must be multiples of 4");
    MOREC_CHECK_ARG(!(p.key_mask && p.cu_seqlens), "attention: key_mask requires fixed-length sequences");
    return MOREC_OK;
}

}  // namespace morec

using namespace morec;

extern "C" int morec_attn_fwd(const void* q, const void* k, const void* v, void* o, const int32_t* cu_seqlens,
                              const float* key_mask, int causal, int n_seq, int seqlen, int n_heads, int head_dim,
                              int ld, int ld_o, float scale, float masked_add, int dtype, float dropout_p, uint64_t seed,
                              uint64_t offset, void* stream) {
    AttnParams p{};
    p.q = q; p.k = k; p.v = v; p.out = o; p.cu_seqlens = cu_seqlens; p.key_mask = key_mask; p.causal = causal;
    p.n_seq = n_seq; p.seqlen = seqlen; p.n_heads = n_heads; p.head_dim = head_dim; p.ld = ld; p.ld_o = ld_o; p.scale = scale;
    p.masked_add = masked_add; p.dropout_p = dropout_p; p.seed = seed; p.offset = offset;
    MOREC_CHECK_ARG(o, "attn_fwd: null output");
    if (int rc = check(p)) return rc;
    if (n_seq <= 0) return MOREC_OK;
    if (tc_attn_eligible(dtype, seqlen, head_dim, ld, ld_o)) return tc_attn_dispatch(tc_params(p), false, dtype, (cudaStream_t)stream);
    const int pairs = n_seq * n_heads;
    int blocks = (pairs + AT_WARPS - 1) / AT_WARPS;
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    constexpr size_t smem = (size_t)AT_FWD_SMEM_FLOATS * sizeof(float);
    static bool attr = false;
    if (!attr) {
        MOREC_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MOREC_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MOREC_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    MOREC_DISPATCH_T(dtype, (attn_fwd_kernel<T><<<blocks, AT_WARPS * 32, smem, (cudaStream_t)stream>>>(p)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_attn_bwd(const void* q, const void* k, const void* v, const void* d_o, void* dq, void* dk,
                              void* dv, const int32_t* cu_seqlens, const float* key_mask, int causal, int n_seq,
                              int seqlen, int n_heads, int head_dim, int ld, int ld_o, float scale, float masked_add,
                              int dtype, float dropout_p, uint64_t seed, uint64_t offset, void* stream) {
    AttnParams p{};
    p.q = q; p.k = k; p.v = v; p.o = d_o; p.dq = dq; p.dk = dk; p.dv = dv; p.cu_seqlens = cu_seqlens;
    p.key_mask = key_mask; p.causal = causal; p.n_seq = n_seq; p.seqlen = seqlen; p.n_heads = n_heads;
    p.head_dim = head_dim; p.ld = ld; p.ld_o = ld_o; p.scale = scale; p.masked_add = masked_add; p.dropout_p = dropout_p;
    p.seed = seed; p.offset = offset;
    MOREC_CHECK_ARG(d_o && dq && dk && dv, "attn_bwd: null pointer");
    if (int rc = check(p)) return rc;
    if (n_seq <= 0) return MOREC_OK;
    if (tc_attn_eligible(dtype, seqlen, head_dim, ld, ld_o)) return tc_attn_dispatch(tc_params(p), true, dtype, (cudaStream_t)stream);
    const int pairs = n_seq * n_heads;
    int blocks = (pairs + AT_WARPS - 1) / AT_WARPS;
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    constexpr size_t smem = (size_t)AT_BWD_SMEM_FLOATS * sizeof(float);
    static bool attr = false;
    if (!attr) {
        MOREC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MOREC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MOREC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    MOREC_DISPATCH_T(dtype, (attn_bwd_kernel<T><<<blocks, AT_WARPS * 32, smem, (cudaStream_t)stream>>>(p)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}
