// General multi-head attention for sequences / windows of up to 128 tokens, with an optional additive bias
// [n_heads, L, L] and an optional additive mask [n_mask, L, L] (sequence s uses mask s % n_mask):
//
//   * Swin window attention (HF SwinSelfAttention; call site inbatch_sasrec_e2e_vision/model/encoders.py:31):
//     L = 49 tokens per window, relative-position bias gathered into [heads, 49, 49], shifted-window mask (0 / -100)
//   * BERT text tower with long titles (cfg-2: T = 128 word pieces), packed by cu_seqlens, no bias
//
// One CTA of NW warps (NW = ceil(L/32): 1, 2 or 4) works on one (sequence, head) pair at a time; thread t owns query
// row t.  K / V / Q / dO are streamed through a [L][64] fp32 shared-memory tile in head-dim chunks of 64 (broadcast
// reads), the score matrix lives in shared memory ([L][L+1], conflict-free by rows and by columns), softmax is exact
// two-pass fp32.  blockIdx.y is the head, so a CTA accumulates the bias gradient of its head in shared memory over
// all the sequences it visits and flushes it with one atomicAdd per entry.
#include "../../../include/morec_b200.h"
#include "attention_common.cuh"
#include "attention_tc.cuh"

namespace morec {

struct GenAttnParams {
    const void *q, *k, *v, *o;          // o: forward output is `out`; backward: dO
    void* out;
    void *dq, *dk, *dv;
    const int* cu_seqlens;              // [n_seq+1] or null (fixed length seqlen)
    const float* bias;                  // [n_heads, seqlen, seqlen] or null
    const float* mask;                  // [n_mask, seqlen, seqlen] or null
    float* dbias;                       // [n_heads, seqlen, seqlen] accumulated (backward) or null
    int n_mask;
    int n_seq, seqlen, n_heads, head_dim, ld, ld_o;
    float scale, dropout_p;
    uint64_t seed, offset;
};

template <int NW>
struct GenCfg {
    static constexpr int LMAX = 32 * NW;
    static constexpr int LS = LMAX + 1;     // score row stride
    static constexpr int THREADS = 32 * NW;
};


// head-dim chunk width DC (32 for Swin's 32-wide heads, 64 otherwise): register slices and the smem tile stride
template <typename T, int DC>
__device__ __forceinline__ void load_row_t(float (&v)[DC], const T* row, int w, bool active) {
#pragma unroll
    for (int t = 0; t < DC / 4; ++t) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && 4 * t < w) x = ld4<T>(row + 4 * t);
        v[4 * t] = x.x; v[4 * t + 1] = x.y; v[4 * t + 2] = x.z; v[4 * t + 3] = x.w;
    }
}
template <typename T, int DC>
__device__ __forceinline__ void store_row_t(T* row, const float (&v)[DC], int w) {
#pragma unroll
    for (int t = 0; t < DC / 4; ++t)
        if (4 * t < w) st4<T>(row + 4 * t, make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]));
}
template <int DC>
__device__ __forceinline__ void dot4keys_t(const float (&v)[DC], const float* tile, int j0, int w, float& a0, float& a1,
                                           float& a2, float& a3) {
    const float4* k0 = reinterpret_cast<const float4*>(tile + (j0 + 0) * DC);
    const float4* k1 = reinterpret_cast<const float4*>(tile + (j0 + 1) * DC);
    const float4* k2 = reinterpret_cast<const float4*>(tile + (j0 + 2) * DC);
    const float4* k3 = reinterpret_cast<const float4*>(tile + (j0 + 3) * DC);
    const int nt = w >> 2;
#pragma unroll
    for (int t = 0; t < DC / 4; ++t) {
        if (t < nt) {
            const float4 x0 = k0[t], x1 = k1[t], x2 = k2[t], x3 = k3[t];
            a0 += v[4 * t] * x0.x + v[4 * t + 1] * x0.y + v[4 * t + 2] * x0.z + v[4 * t + 3] * x0.w;
            a1 += v[4 * t] * x1.x + v[4 * t + 1] * x1.y + v[4 * t + 2] * x1.z + v[4 * t + 3] * x1.w;
            a2 += v[4 * t] * x2.x + v[4 * t + 1] * x2.y + v[4 * t + 2] * x2.z + v[4 * t + 3] * x2.w;
            a3 += v[4 * t] * x3.x + v[4 * t + 1] * x3.y + v[4 * t + 2] * x3.z + v[4 * t + 3] * x3.w;
        }
    }
}

template <typename T, int NW, int DC>
__device__ __forceinline__ void load_tile_blk(float* tile, const T* base, int ld, int row0, int len, int col0, int w) {
    const int nv = w >> 2;
    for (int idx = threadIdx.x; idx < len * nv; idx += 32 * NW) {
        const int r = idx / nv, c = (idx - r * nv) << 2;
        *reinterpret_cast<float4*>(tile + r * DC + c) = ld4<T>(base + (size_t)(row0 + r) * ld + col0 + c);
    }
}

// S[row][j] (+)= <x_row, tile_j> for all keys; x = this thread's row of Xb (Q or dO)
template <typename T, int NW, int DC>
__device__ __forceinline__ void raw_scores_blk(float* tile, float* S, const T* Xb, int ldx, const T* Kb, int ldk,
                                               int head_dim, int colh, int row0, int len) {
    using G = GenCfg<NW>;
    const int row = threadIdx.x;
    const int len4 = (len + 3) & ~3;
    for (int dc = 0; dc < head_dim; dc += DC) {
        const int w = min(DC, head_dim - dc);
        __syncthreads();
        load_tile_blk<T, NW, DC>(tile, Kb, ldk, row0, len, colh + dc, w);
        __syncthreads();
        float xv[DC];
        load_row_t<T, DC>(xv, Xb + (size_t)(row0 + (row < len ? row : 0)) * ldx + colh + dc, w, row < len);
#pragma unroll 1
        for (int j0 = 0; j0 < len4; j0 += 4) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            dot4keys_t<DC>(xv, tile, j0, w, a0, a1, a2, a3);
            float* sr = S + row * G::LS + j0;
            if (dc == 0) { sr[0] = a0; sr[1] = a1; sr[2] = a2; sr[3] = a3; }
            else { sr[0] += a0; sr[1] += a1; sr[2] += a2; sr[3] += a3; }
        }
    }
}

template <int NW, bool TRANSPOSED, int DC>
__device__ __forceinline__ void weighted_rows_blk(const float* coef, const float* tile, int len, float (&acc)[DC]) {
    using G = GenCfg<NW>;
    const int row = threadIdx.x;
#pragma unroll
    for (int t = 0; t < DC; ++t) acc[t] = 0.f;
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
        const float d = TRANSPOSED ? coef[j * G::LS + row] : coef[row * G::LS + j];
        const float4* r = reinterpret_cast<const float4*>(tile + j * DC);
#pragma unroll
        for (int t = 0; t < DC / 4; ++t) {
            const float4 x = r[t];
            acc[4 * t] += d * x.x; acc[4 * t + 1] += d * x.y; acc[4 * t + 2] += d * x.z; acc[4 * t + 3] += d * x.w;
        }
    }
}

__device__ __forceinline__ bool gen_keep(const GenAttnParams& p, int pair, int i, int j, uint32_t th) {
    const uint4 r = Philox::gen(p.seed, p.offset + ((uint64_t)pair * 128 + i) * 32 + (j >> 2));
    const uint32_t w = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
    return w >= th;
}

// scale + bias + mask + softmax of this thread's row, in place
template <int NW>
__device__ __forceinline__ void softmax_row_blk(const GenAttnParams& p, float* S, int s, int h, int len) {
    using G = GenCfg<NW>;
    const int row = threadIdx.x;
    if (row >= len) return;
    float* sr = S + row * G::LS;
    const float* br = p.bias ? p.bias + ((size_t)h * p.seqlen + row) * p.seqlen : nullptr;
    const float* mr = p.mask ? p.mask + ((size_t)(s % p.n_mask) * p.seqlen + row) * p.seqlen : nullptr;
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < len; ++j) {
        float x = sr[j] * p.scale;
        if (br) x += br[j];
        if (mr) x += mr[j];
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

__device__ __forceinline__ void gen_range(const GenAttnParams& p, int s, int lmax, int& row0, int& len) {
    if (p.cu_seqlens) { row0 = p.cu_seqlens[s]; len = p.cu_seqlens[s + 1] - row0; }
    else { row0 = s * p.seqlen; len = p.seqlen; }
    if (len > lmax) len = lmax;
}

template <typename T, int NW, int DC>
__global__ void __launch_bounds__(32 * NW) attn_gen_fwd_kernel(const GenAttnParams p) {
    using G = GenCfg<NW>;
    extern __shared__ __align__(16) float ag_smem[];
    float* tile = ag_smem;                       // [LMAX][64]
    float* S = ag_smem + G::LMAX * DC;       // [LMAX][LS]
    const int row = threadIdx.x;
    const int h = blockIdx.y;
    const int colh = h * p.head_dim;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    T* O = reinterpret_cast<T*>(p.out);
    const uint32_t th = (uint32_t)fminf(p.dropout_p * 4294967296.f, 4294967295.f);
    for (int s = blockIdx.x; s < p.n_seq; s += gridDim.x) {
        int row0, len;
        gen_range(p, s, G::LMAX, row0, len);
        if (len <= 0) continue;
        raw_scores_blk<T, NW, DC>(tile, S, Q, p.ld, K, p.ld, p.head_dim, colh, row0, len);
        softmax_row_blk<NW>(p, S, s, h, len);
        if (p.dropout_p > 0.f && row < len) {
            const float sc = 1.f / (1.f - p.dropout_p);
            const int pair = s * p.n_heads + h;
#pragma unroll 1
            for (int j = 0; j < len; ++j) S[row * G::LS + j] = gen_keep(p, pair, row, j, th) ? S[row * G::LS + j] * sc : 0.f;
        }
        for (int dc = 0; dc < p.head_dim; dc += DC) {
            const int w = min(DC, p.head_dim - dc);
            __syncthreads();
            load_tile_blk<T, NW, DC>(tile, V, p.ld, row0, len, colh + dc, w);
            __syncthreads();
            float ov[DC];
            weighted_rows_blk<NW, false, DC>(S, tile, len, ov);
            if (row < len) store_row_t<T, DC>(O + (size_t)(row0 + row) * p.ld_o + colh + dc, ov, w);
        }
        __syncthreads();
    }
}

template <typename T, int NW, int DC>
__global__ void __launch_bounds__(32 * NW) attn_gen_bwd_kernel(const GenAttnParams p) {
    using G = GenCfg<NW>;
    extern __shared__ __align__(16) float ag_smem[];
    float* tile = ag_smem;                                   // [LMAX][64]
    float* Pm = ag_smem + G::LMAX * DC;                  // [LMAX][LS]  P, then P~
    float* dSm = Pm + G::LMAX * G::LS;                       // [LMAX][LS]  dP~, then dS*scale
    float* dB = dSm + G::LMAX * G::LS;                       // [seqlen*seqlen] bias-gradient accumulator (if dbias)
    const int row = threadIdx.x;
    const int h = blockIdx.y;
    const int colh = h * p.head_dim;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    const T* dO = reinterpret_cast<const T*>(p.o);
    T* dQ = reinterpret_cast<T*>(p.dq);
    T* dK = reinterpret_cast<T*>(p.dk);
    T* dV = reinterpret_cast<T*>(p.dv);
    const uint32_t th = (uint32_t)fminf(p.dropout_p * 4294967296.f, 4294967295.f);
    const int LL = p.seqlen * p.seqlen;
    if (p.dbias)
        for (int i = threadIdx.x; i < LL; i += blockDim.x) dB[i] = 0.f;
    for (int s = blockIdx.x; s < p.n_seq; s += gridDim.x) {
        int row0, len;
        gen_range(p, s, G::LMAX, row0, len);
        if (len <= 0) continue;
        raw_scores_blk<T, NW, DC>(tile, Pm, Q, p.ld, K, p.ld, p.head_dim, colh, row0, len);
        softmax_row_blk<NW>(p, Pm, s, h, len);
        raw_scores_blk<T, NW, DC>(tile, dSm, dO, p.ld_o, V, p.ld, p.head_dim, colh, row0, len);   // dP~ = dO . V^T
        {
            const float sc = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
            const int pair = s * p.n_heads + h;
            float* pr = Pm + row * G::LS;
            float* dr = dSm + row * G::LS;
            const bool act = row < len;
            float dsum = 0.f;
#pragma unroll 1
            for (int j = 0; j < len; ++j) {
                const float keep = (p.dropout_p > 0.f && act) ? (gen_keep(p, pair, row, j, th) ? sc : 0.f) : 1.f;
                const float d = act ? dr[j] * keep : 0.f;
                dr[j] = d;
                dsum += (act ? pr[j] : 0.f) * d;
            }
#pragma unroll 1
            for (int j = 0; j < len; ++j) {
                const float keep = (p.dropout_p > 0.f && act) ? (gen_keep(p, pair, row, j, th) ? sc : 0.f) : 1.f;
                const float pj = act ? pr[j] : 0.f;
                const float ds = pj * (dr[j] - dsum);           // gradient of the pre-softmax score
                if (p.dbias && act) dB[row * p.seqlen + j] += ds;
                dr[j] = ds * p.scale;
                pr[j] = pj * keep;
            }
        }
        __syncthreads();
        for (int dc = 0; dc < p.head_dim; dc += DC) {      // dQ_i = sum_j dS_ij K_j
            const int w = min(DC, p.head_dim - dc);
            __syncthreads();
            load_tile_blk<T, NW, DC>(tile, K, p.ld, row0, len, colh + dc, w);
            __syncthreads();
            float acc[DC];
            weighted_rows_blk<NW, false, DC>(dSm, tile, len, acc);
            if (row < len) store_row_t<T, DC>(dQ + (size_t)(row0 + row) * p.ld + colh + dc, acc, w);
        }
        for (int dc = 0; dc < p.head_dim; dc += DC) {      // dK_j = sum_i dS_ij Q_i
            const int w = min(DC, p.head_dim - dc);
            __syncthreads();
            load_tile_blk<T, NW, DC>(tile, Q, p.ld, row0, len, colh + dc, w);
            __syncthreads();
            float acc[DC];
            weighted_rows_blk<NW, true, DC>(dSm, tile, len, acc);
            if (row < len) store_row_t<T, DC>(dK + (size_t)(row0 + row) * p.ld + colh + dc, acc, w);
        }
        for (int dc = 0; dc < p.head_dim; dc += DC) {      // dV_j = sum_i P~_ij dO_i
            const int w = min(DC, p.head_dim - dc);
            __syncthreads();
            load_tile_blk<T, NW, DC>(tile, dO, p.ld_o, row0, len, colh + dc, w);
            __syncthreads();
            float acc[DC];
            weighted_rows_blk<NW, true, DC>(Pm, tile, len, acc);
            if (row < len) store_row_t<T, DC>(dV + (size_t)(row0 + row) * p.ld + colh + dc, acc, w);
        }
        __syncthreads();
    }
    if (p.dbias) {
        __syncthreads();
        float* g = p.dbias + (size_t)h * LL;
        for (int i = threadIdx.x; i < LL; i += blockDim.x)
            if (dB[i] != 0.f) atomicAdd(g + i, dB[i]);
    }
}

template <typename T, int NW, int DC>
static int launch_gen(const GenAttnParams& p, bool bwd, cudaStream_t stream) {
    using G = GenCfg<NW>;
    size_t smem = (size_t)(G::LMAX * DC + (bwd ? 2 : 1) * G::LMAX * G::LS) * sizeof(float);
    if (bwd && p.dbias) smem += (size_t)p.seqlen * p.seqlen * sizeof(float);
    int gx = (num_sms() * (NW == 4 ? 2 : 4) + p.n_heads - 1) / p.n_heads;
    if (gx > p.n_seq) gx = p.n_seq;
    if (gx < 1) gx = 1;
    dim3 grid(gx, p.n_heads);
    if (!bwd) {
        auto kern = attn_gen_fwd_kernel<T, NW, DC>;
        MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        kern<<<grid, 32 * NW, smem, stream>>>(p);
    } else {
        auto kern = attn_gen_bwd_kernel<T, NW, DC>;
        MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        kern<<<grid, 32 * NW, smem, stream>>>(p);
    }
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

static int dispatch_gen(const GenAttnParams& p, bool bwd, int dtype, cudaStream_t stream) {
    MOREC_CHECK_ARG(p.q && p.k && p.v, "attn_gen: null q/k/v");
    MOREC_CHECK_ARG(p.seqlen > 0 && p.seqlen <= 128, "attn_gen: sequence length %d not in [1, 128]", p.seqlen);
    MOREC_CHECK_ARG(p.head_dim > 0 && p.head_dim % 4 == 0, "attn_gen: head_dim=%d must be a multiple of 4", p.head_dim);
    MOREC_CHECK_ARG(p.ld % 4 == 0 && p.ld_o % 4 == 0, "attn_gen: row strides must be multiples of 4");
    MOREC_CHECK_ARG(!p.mask || p.n_mask > 0, "attn_gen: mask needs n_mask > 0");
    MOREC_CHECK_ARG(!(p.cu_seqlens && (p.bias || p.mask)), "attn_gen: bias / mask require fixed-length sequences");
    if (p.n_seq <= 0) return MOREC_OK;
    if (tc_attn_eligible(dtype, p.seqlen, p.head_dim, p.ld, p.ld_o)) {      // fast modes: tensor-core kernels
        TcAttnParams t{};
        t.q = p.q; t.k = p.k; t.v = p.v; t.o = p.o; t.out = p.out; t.dq = p.dq; t.dk = p.dk; t.dv = p.dv;
        t.cu_seqlens = p.cu_seqlens; t.bias = p.bias; t.mask = p.mask; t.n_mask = p.n_mask; t.dbias = p.dbias;
        t.n_seq = p.n_seq; t.seqlen = p.seqlen; t.n_heads = p.n_heads; t.head_dim = p.head_dim; t.ld = p.ld; t.ld_o = p.ld_o;
        t.scale = p.scale; t.dropout_p = p.dropout_p; t.seed = p.seed; t.offset = p.offset;
        return tc_attn_dispatch(t, bwd, dtype, stream);
    }
    const int nw = (p.seqlen + 31) / 32;
    const bool narrow = p.head_dim <= 32;         // Swin heads are 32 wide: half-width register slices / smem tile
#define MOREC_GEN(TT, NWW)                                                                         \
    return narrow ? launch_gen<TT, NWW, 32>(p, bwd, stream) : launch_gen<TT, NWW, 64>(p, bwd, stream)
    if (dtype == 1) {
        if (nw <= 1) { MOREC_GEN(__nv_bfloat16, 1); }
        if (nw == 2) { MOREC_GEN(__nv_bfloat16, 2); }
        MOREC_GEN(__nv_bfloat16, 4);
    }
    if (dtype == 3) {
        if (nw <= 1) { MOREC_GEN(__half, 1); }
        if (nw == 2) { MOREC_GEN(__half, 2); }
        MOREC_GEN(__half, 4);
    }
    if (nw <= 1) { MOREC_GEN(float, 1); }
    if (nw == 2) { MOREC_GEN(float, 2); }
    MOREC_GEN(float, 4);
#undef MOREC_GEN
}

}  // namespace morec

using namespace morec;

extern "C" int morec_attn_gen_fwd(const void* q, const void* k, const void* v, void* o, const int32_t* cu_seqlens,
                                  const float* bias, const float* mask, int n_mask, int n_seq, int seqlen, int n_heads,
                                  int head_dim, int ld, int ld_o, float scale, int dtype, float dropout_p, uint64_t seed,
                                  uint64_t offset, void* stream) {
    GenAttnParams p{};
    p.q = q; p.k = k; p.v = v; p.out = o; p.cu_seqlens = cu_seqlens; p.bias = bias; p.mask = mask; p.n_mask = n_mask;
    p.n_seq = n_seq; p.seqlen = seqlen; p.n_heads = n_heads; p.head_dim = head_dim; p.ld = ld; p.ld_o = ld_o;
    p.scale = scale; p.dropout_p = dropout_p; p.seed = seed; p.offset = offset;
    MOREC_CHECK_ARG(o, "attn_gen_fwd: null output");
    return dispatch_gen(p, false, dtype, (cudaStream_t)stream);
}

extern "C" int morec_attn_gen_bwd(const void* q, const void* k, const void* v, const void* d_o, void* dq, void* dk,
                                  void* dv, float* dbias, const int32_t* cu_seqlens, const float* bias,
                                  const float* mask, int n_mask, int n_seq, int seqlen, int n_heads, int head_dim, int ld,
                                  int ld_o, float scale, int dtype, float dropout_p, uint64_t seed, uint64_t offset,
                                  void* stream) {
    GenAttnParams p{};
    p.q = q; p.k = k; p.v = v; p.o = d_o; p.dq = dq; p.dk = dk; p.dv = dv; p.dbias = dbias; p.cu_seqlens = cu_seqlens;
    p.bias = bias; p.mask = mask; p.n_mask = n_mask; p.n_seq = n_seq; p.seqlen = seqlen; p.n_heads = n_heads;
    p.head_dim = head_dim; p.ld = ld; p.ld_o = ld_o; p.scale = scale; p.dropout_p = dropout_p; p.seed = seed;
    p.offset = offset;
    MOREC_CHECK_ARG(d_o && dq && dk && dv, "attn_gen_bwd: null pointer");
    return dispatch_gen(p, true, dtype, (cudaStream_t)stream);
}
