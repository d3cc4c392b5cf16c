// Tensor-core attention for short sequences / windows (<= 64 tokens, head_dim 32 / 64; <= 32 tokens, head_dim 256),
// forward and backward.
//
// Used by morec_attn_* and morec_attn_gen_* in the fast precision modes (C-ABI dtype 0 = "tf32", 1 = "bf16"); the
// parity mode (dtype 2) keeps the exact-fp32 SIMT kernels.  Replaces the same reference sites:
//   HF BertSelfAttention (call site model/encoders.py:68), SASRec SelfAttention (model/modules.py:27-31),
//   HF SwinSelfAttention (call site inbatch_sasrec_e2e_vision/model/encoders.py:31).
//
// The per-(sequence, head) problem is tiny (S = Q K^T is at most 64 x 64 x 64), far below one tcgen05 tile
// (M = 128 would be 1/4 .. 1/16 filled by block-diagonal work), and the op is bound by streaming q/k/v/dO through the
// SM once.  So the matrix products run on warp-level mma.sync.m16n8k8 (TF32 inputs, fp32 accumulate; bf16 storage is
// exactly representable in TF32) and the design effort goes into data movement:
//   * one CTA of LP/16 warps per (sequence, head); warp w owns the 16-query stripe [16w, 16w+16) of S, all keys;
//   * Q, K, V (, dO) are staged ONCE into shared memory as TF32 (row stride D+4 floats: every fragment load below is
//     bank-conflict free), zero-filled to LP rows;
//   * the softmax runs on the accumulator fragments in registers (row statistics by quad shuffles); P (and dS in the
//     backward) feed the next mma straight from the accumulator registers: the C-fragment holds columns (2t, 2t+1)
//     of row g, the A-fragment wants (t, t+4), so the k index of that product is permuted (k' = t <-> column 2t,
//     k' = t+4 <-> column 2t+1) on BOTH operands, which a sum over k does not notice;
//   * backward: the transposed products (dK = dS^T Q, dV = P~^T dO) read dS / P~ from a [LP][LP+4] shared tile written
//     once from the accumulator layout; warp w then owns the 16-KEY stripe.
// Dropout keep bits: one Philox4x32-7 block per (pair, warp, n-tile PAIR, lane), 16 bits per decision = exactly the eight
// accumulator elements of that thread's two 16x8 tiles, identical in forward and backward.
#pragma once
#include "attention_common.cuh"

namespace morec {

struct TcAttnParams {
    const void *q, *k, *v, *o;          // o: backward dO
    void* out;                          // forward output
    void *dq, *dk, *dv;
    const int* cu_seqlens;              // [n_seq+1] or null
    const float* key_mask;              // [n_seq, seqlen] non-zero = valid key, or null
    int causal;
    float masked_add;                   // added where causal / key_mask reject (reference: -1e9)
    const float* bias;                  // [n_heads, seqlen, seqlen] or null
    const float* mask;                  // [n_mask, seqlen, seqlen] or null
    int n_mask;
    float* dbias;                       // [n_heads, seqlen, seqlen] accumulated, or null
    int n_seq, seqlen, n_heads, head_dim, ld, ld_o;
    float scale, dropout_p;
    uint64_t seed, offset;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int D, int LP>
struct TcCfg {
    static constexpr int NW = LP / 16;          // warps per CTA = 16-row stripes
    static constexpr int THREADS = 32 * NW;
    static constexpr int SQ = D + 4;            // operand tile row stride (floats), == 4 mod 32
    static constexpr int SS = LP + 4;           // score tile row stride
    static constexpr int NT = LP / 8;           // key n-tiles of a stripe of S
    static constexpr int DC = D > 64 ? 64 : D;  // output stripes are produced DC columns at a time
    static constexpr int MT = DC / 8;           // n-tiles of a [16, DC] output stripe
    static constexpr int TILE = LP * SQ;
};

// rows [row0, row0+len) x cols [col0, col0+D) of a [*, ld] matrix -> dst[LP][SQ] as TF32 bit patterns; rows >= len zero
template <typename T, int D, int LP>
__device__ __forceinline__ void tc_load_tile(float* dst, const T* base, int ld, int row0, int len, int col0) {
    using C = TcCfg<D, LP>;
    constexpr int NV = D / 4;
    constexpr int ITER = LP * NV / C::THREADS;           // exact: LP * NV is a multiple of THREADS
    constexpr int BATCH = ITER < 8 ? ITER : 8;           // global loads issued back to back (memory-level parallelism)
#pragma unroll 1
    for (int i0 = 0; i0 < ITER; i0 += BATCH) {
        float4 x[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int idx = threadIdx.x + (i0 + i) * C::THREADS;
            const int r = idx / NV, c = (idx - r * NV) << 2;
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < len) x[i] = ld4<T>(base + (size_t)(row0 + r) * ld + col0 + c);
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int idx = threadIdx.x + (i0 + i) * C::THREADS;
            const int r = idx / NV, c = (idx - r * NV) << 2;
            uint4 u;
            u.x = to_tf32(x[i].x); u.y = to_tf32(x[i].y); u.z = to_tf32(x[i].z); u.w = to_tf32(x[i].w);
            *reinterpret_cast<uint4*>(dst + r * C::SQ + c) = u;
        }
    }
}

// acc[n] (16 x 8 tiles, n < NT) = X[m0 .. m0+16, :] . Y[8n .. 8n+8, :]^T   over the D columns (both row operands)
template <int D, int LP>
__device__ __forceinline__ void tc_rows_dot_rows(const float* X, const float* Y, int m0, int ncols, float (&acc)[LP / 8][4],
                                                 int g, int t) {
    using C = TcCfg<D, LP>;
    const uint32_t* Xu = reinterpret_cast<const uint32_t*>(X);
    const uint32_t* Yu = reinterpret_cast<const uint32_t*>(Y);
#pragma unroll
    for (int n = 0; n < C::NT; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < D / 8; ++kk) {
        const uint32_t a0 = Xu[(m0 + g) * C::SQ + kk * 8 + t];
        const uint32_t a1 = Xu[(m0 + g + 8) * C::SQ + kk * 8 + t];
        const uint32_t a2 = Xu[(m0 + g) * C::SQ + kk * 8 + t + 4];
        const uint32_t a3 = Xu[(m0 + g + 8) * C::SQ + kk * 8 + t + 4];
#pragma unroll
        for (int n = 0; n < C::NT; ++n) {
            if (n * 8 < ncols) {
                const uint32_t b0 = Yu[(n * 8 + g) * C::SQ + kk * 8 + t];
                const uint32_t b1 = Yu[(n * 8 + g) * C::SQ + kk * 8 + t + 4];
                mma_tf32(acc[n], a0, a1, a2, a3, b0, b1);
            }
        }
    }
}

// out[m] (16 x 8 tiles, m < MT) = sum over key steps ks of  A_regs[ks] (accumulator layout, permuted k) . Y[8ks.., 8m..]
template <int D, int LP>
__device__ __forceinline__ void tc_regs_dot_cols(const float (&a)[LP / 8][4], const float* Y, int c0, int nrows,
                                                 float (&out)[TcCfg<D, LP>::MT][4], int g, int t) {
    using C = TcCfg<D, LP>;
    const uint32_t* Yu = reinterpret_cast<const uint32_t*>(Y);
#pragma unroll
    for (int m = 0; m < C::MT; ++m) { out[m][0] = out[m][1] = out[m][2] = out[m][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < C::NT; ++ks) {
        if (ks * 8 < nrows) {
            const uint32_t a0 = to_tf32(a[ks][0]), a1 = to_tf32(a[ks][2]), a2 = to_tf32(a[ks][1]), a3 = to_tf32(a[ks][3]);
            const uint32_t* r0 = Yu + (ks * 8 + 2 * t) * C::SQ + c0 + g;
#pragma unroll
            for (int m = 0; m < C::MT; ++m) mma_tf32(out[m], a0, a1, a2, a3, r0[m * 8], r0[C::SQ + m * 8]);
        }
    }
}

// out[m] = sum over query steps ks of  Sm^T[j0 .. j0+16, 8ks ..] . Y[8ks .., 8m ..]      (Sm: [LP][SS], transposed read)
template <int D, int LP>
__device__ __forceinline__ void tc_smT_dot_cols(const float* Sm, const float* Y, int c0, int j0, int nrows,
                                                float (&out)[TcCfg<D, LP>::MT][4], int g, int t) {
    using C = TcCfg<D, LP>;
    const uint32_t* Su = reinterpret_cast<const uint32_t*>(Sm);
    const uint32_t* Yu = reinterpret_cast<const uint32_t*>(Y);
#pragma unroll
    for (int m = 0; m < C::MT; ++m) { out[m][0] = out[m][1] = out[m][2] = out[m][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < C::NT; ++ks) {
        if (ks * 8 < nrows) {
            const uint32_t* s0 = Su + (ks * 8 + 2 * t) * C::SS + j0 + g;
            const uint32_t a0 = s0[0], a1 = s0[8], a2 = s0[C::SS], a3 = s0[C::SS + 8];
            const uint32_t* r0 = Yu + (ks * 8 + 2 * t) * C::SQ + c0 + g;
#pragma unroll
            for (int m = 0; m < C::MT; ++m) mma_tf32(out[m], a0, a1, a2, a3, r0[m * 8], r0[C::SQ + m * 8]);
        }
    }
}

// store a [16, DC] accumulator stripe: rows m0+g and m0+g+8 (if < len), columns col0 + 8m + 2t, +1
template <typename T, int DC>
__device__ __forceinline__ void tc_store_stripe(T* base, int ld, int row0, int len, int col0, int m0,
                                                const float (&o)[DC / 8][4], int g, int t) {
    const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
    for (int m = 0; m < DC / 8; ++m) {
        const int c = col0 + m * 8 + 2 * t;
        if constexpr (sizeof(T) == 4) {
            if (r0 < len) *reinterpret_cast<float2*>(base + (size_t)(row0 + r0) * ld + c) = make_float2(o[m][0], o[m][1]);
            if (r1 < len) *reinterpret_cast<float2*>(base + (size_t)(row0 + r1) * ld + c) = make_float2(o[m][2], o[m][3]);
        } else {
            constexpr bool f16 = __is_same(T, __half);
            if (r0 < len) *reinterpret_cast<uint32_t*>(base + (size_t)(row0 + r0) * ld + c) = pack2_16(o[m][0], o[m][1], f16);
            if (r1 < len) *reinterpret_cast<uint32_t*>(base + (size_t)(row0 + r1) * ld + c) = pack2_16(o[m][2], o[m][3], f16);
        }
    }
}

__device__ __forceinline__ void tc_range(const TcAttnParams& p, int s, int lmax, int& row0, int& len) {
    if (p.cu_seqlens) { row0 = p.cu_seqlens[s]; len = p.cu_seqlens[s + 1] - row0; }
    else { row0 = s * p.seqlen; len = p.seqlen; }
    if (len > lmax) len = lmax;
}

// scale + bias + mask + causal / key-valid rejection, then softmax over the keys, on a warp's accumulator stripe.
// On return acc holds P (rows >= len and columns >= len are exactly 0).
template <int LP>
__device__ __forceinline__ void tc_softmax_stripe(const TcAttnParams& p, float (&acc)[LP / 8][4], int s, int h, int m0,
                                                  int len, int g, int t) {
    constexpr int NT = LP / 8;
    const int L = p.seqlen;
    const float* bias = p.bias ? p.bias + (size_t)h * L * L : nullptr;
    const float* mask = p.mask ? p.mask + (size_t)(s % p.n_mask) * L * L : nullptr;
    const float* kmask = p.key_mask ? p.key_mask + (size_t)s * L : nullptr;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row = m0 + g + ((e >> 1) << 3), col = n * 8 + 2 * t + (e & 1);
            float x = -INFINITY;
            if (col < len && row < len) {
                x = acc[n][e] * p.scale;
                if (bias) x += bias[row * L + col];
                if (mask) x += mask[row * L + col];
                bool ok = true;
                if (p.causal) ok = col <= row;
                if (kmask) ok = ok && (kmask[col] != 0.f);
                if (!ok) x += p.masked_add;
            }
            acc[n][e] = x;
            mx[e >> 1] = fmaxf(mx[e >> 1], x);
        }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        if (mx[r] == -INFINITY) mx[r] = 0.f;            // padding row: every entry is -inf -> P = 0
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float pe = __expf(acc[n][e] - mx[e >> 1]);
            acc[n][e] = pe;
            sum[e >> 1] += pe;
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
        sum[r] = sum[r] > 0.f ? 1.f / sum[r] : 0.f;
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        acc[n][0] *= sum[0]; acc[n][1] *= sum[0]; acc[n][2] *= sum[1]; acc[n][3] *= sum[1];
    }
}

// keep multipliers (0 or 1/(1-p)) of the four accumulator elements of n-tiles 2*np and 2*np+1 of this warp's stripe:
// one Philox4x32-7 block per (pair, warp, n-tile pair, lane), 16 bits per decision (common.cuh)
__device__ __forceinline__ void tc_keep8(const TcAttnParams& p, int pair, int nw, int nt, int warp, int np, int lane,
                                         uint32_t th16, float sc, float (&k0)[4], float (&k1)[4]) {
    const uint4 r = Philox7::gen(p.seed, p.offset + (((uint64_t)pair * nw + warp) * (nt / 2) + np) * 32 + lane);
    k0[0] = (r.x & 0xffffu) >= th16 ? sc : 0.f;  k1[0] = (r.x >> 16) >= th16 ? sc : 0.f;
    k0[1] = (r.y & 0xffffu) >= th16 ? sc : 0.f;  k1[1] = (r.y >> 16) >= th16 ? sc : 0.f;
    k0[2] = (r.z & 0xffffu) >= th16 ? sc : 0.f;  k1[2] = (r.z >> 16) >= th16 ? sc : 0.f;
    k0[3] = (r.w & 0xffffu) >= th16 ? sc : 0.f;  k1[3] = (r.w >> 16) >= th16 ? sc : 0.f;
}
// all keep multipliers of a stripe (1 everywhere when dropout is off; tiles at or beyond `len` are left at 1)
template <int NT>
__device__ __forceinline__ void tc_keep_stripe(const TcAttnParams& p, int pair, int nw, int warp, int lane, int len,
                                               float (&keep)[NT][4]) {
#pragma unroll
    for (int n = 0; n < NT; ++n) keep[n][0] = keep[n][1] = keep[n][2] = keep[n][3] = 1.f;
    if (!(p.dropout_p > 0.f)) return;
    const uint32_t th16 = drop_thresh16(p.dropout_p);
    const float sc = 1.f / (1.f - p.dropout_p);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np)
        if (np * 16 < len) tc_keep8(p, pair, nw, NT, warp, np, lane, th16, sc, keep[2 * np], keep[2 * np + 1]);
}

template <typename T, int D, int LP>
__global__ void __launch_bounds__(TcCfg<D, LP>::THREADS) attn_tc_fwd_kernel(const TcAttnParams p) {
    using C = TcCfg<D, LP>;
    extern __shared__ __align__(16) float tc_smem[];
    float* Qs = tc_smem;
    float* Ks = Qs + C::TILE;
    float* Vs = Ks + C::TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, colh = h * D;
    const int m0 = warp * 16;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    T* O = reinterpret_cast<T*>(p.out);
    for (int s = blockIdx.x; s < p.n_seq; s += gridDim.x) {
        int row0, len;
        tc_range(p, s, LP, row0, len);
        if (len <= 0) continue;
        __syncthreads();                                   // previous pair's readers are done with the tiles
        tc_load_tile<T, D, LP>(Qs, Q, p.ld, row0, len, colh);
        tc_load_tile<T, D, LP>(Ks, K, p.ld, row0, len, colh);
        tc_load_tile<T, D, LP>(Vs, V, p.ld, row0, len, colh);
        __syncthreads();
        if (m0 >= len) continue;                           // stripe of padding rows (warp-uniform; barriers are at the loop top)
        float acc[C::NT][4];
        tc_rows_dot_rows<D, LP>(Qs, Ks, m0, len, acc, g, t);
        tc_softmax_stripe<LP>(p, acc, s, h, m0, len, g, t);
        if (p.dropout_p > 0.f) {
            float keep[C::NT][4];
            tc_keep_stripe<C::NT>(p, s * p.n_heads + h, C::NW, warp, lane, len, keep);
#pragma unroll
            for (int n = 0; n < C::NT; ++n) {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[n][e] *= keep[n][e];
            }
        }
#pragma unroll 1
        for (int c0 = 0; c0 < D; c0 += C::DC) {
            float o[C::MT][4];
            tc_regs_dot_cols<D, LP>(acc, Vs, c0, len, o, g, t);
            tc_store_stripe<T, C::DC>(O, p.ld_o, row0, len, colh + c0, m0, o, g, t);
        }
    }
}

template <typename T, int D, int LP>
__global__ void __launch_bounds__(TcCfg<D, LP>::THREADS) attn_tc_bwd_kernel(const TcAttnParams p) {
    using C = TcCfg<D, LP>;
    extern __shared__ __align__(16) float tc_smem[];
    float* Qs = tc_smem;
    float* Ks = Qs + C::TILE;
    float* Vs = Ks + C::TILE;
    float* Gs = Vs + C::TILE;                  // dO
    float* Pm = Gs + C::TILE;                  // [LP][SS]  P~ (dropped, rescaled)
    float* dSm = Pm + LP * C::SS;              // [LP][SS]  dS * scale
    float* dB = dSm + LP * C::SS;              // [seqlen*seqlen] bias-gradient accumulator (if dbias)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = blockIdx.y, colh = h * D;
    const int m0 = warp * 16;
    const T* Q = reinterpret_cast<const T*>(p.q);
    const T* K = reinterpret_cast<const T*>(p.k);
    const T* V = reinterpret_cast<const T*>(p.v);
    const T* dO = reinterpret_cast<const T*>(p.o);
    T* dQ = reinterpret_cast<T*>(p.dq);
    T* dK = reinterpret_cast<T*>(p.dk);
    T* dV = reinterpret_cast<T*>(p.dv);
    const int LL = p.seqlen * p.seqlen;
    if (p.dbias)
        for (int i = threadIdx.x; i < LL; i += C::THREADS) dB[i] = 0.f;
    for (int s = blockIdx.x; s < p.n_seq; s += gridDim.x) {
        int row0, len;
        tc_range(p, s, LP, row0, len);
        if (len <= 0) continue;
        __syncthreads();
        tc_load_tile<T, D, LP>(Qs, Q, p.ld, row0, len, colh);
        tc_load_tile<T, D, LP>(Ks, K, p.ld, row0, len, colh);
        tc_load_tile<T, D, LP>(Vs, V, p.ld, row0, len, colh);
        tc_load_tile<T, D, LP>(Gs, dO, p.ld_o, row0, len, colh);
        __syncthreads();
        if (m0 < len) {                                    // ---- query-stripe phase
            float pr[C::NT][4], dp[C::NT][4];
            tc_rows_dot_rows<D, LP>(Qs, Ks, m0, len, pr, g, t);
            tc_softmax_stripe<LP>(p, pr, s, h, m0, len, g, t);
            tc_rows_dot_rows<D, LP>(Gs, Vs, m0, len, dp, g, t);          // dP~ = dO . V^T
            float keepm[C::NT][4];
            tc_keep_stripe<C::NT>(p, s * p.n_heads + h, C::NW, warp, lane, len, keepm);
            float dsum[2] = {0.f, 0.f};
#pragma unroll
            for (int n = 0; n < C::NT; ++n) {
                if (n * 8 < len) {
                    const float (&keep)[4] = keepm[n];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        dp[n][e] *= keep[e];                             // dP
                        dsum[e >> 1] += pr[n][e] * dp[n][e];
                    }
                    // P~ for dV
                    float* w0 = Pm + (m0 + g) * C::SS + n * 8 + 2 * t;
                    float* w1 = Pm + (m0 + g + 8) * C::SS + n * 8 + 2 * t;
                    uint2 u0, u1;
                    u0.x = to_tf32(pr[n][0] * keep[0]); u0.y = to_tf32(pr[n][1] * keep[1]);
                    u1.x = to_tf32(pr[n][2] * keep[2]); u1.y = to_tf32(pr[n][3] * keep[3]);
                    *reinterpret_cast<uint2*>(w0) = u0;
                    *reinterpret_cast<uint2*>(w1) = u1;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
                dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
            }
#pragma unroll
            for (int n = 0; n < C::NT; ++n) {
                if (n * 8 < len) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float ds = pr[n][e] * (dp[n][e] - dsum[e >> 1]);     // gradient of the pre-softmax score
                        if (p.dbias) {
                            const int row = m0 + g + ((e >> 1) << 3), col = n * 8 + 2 * t + (e & 1);
                            if (row < len && col < len) dB[row * p.seqlen + col] += ds;
                        }
                        dp[n][e] = ds * p.scale;
                    }
                    float* w0 = dSm + (m0 + g) * C::SS + n * 8 + 2 * t;
                    float* w1 = dSm + (m0 + g + 8) * C::SS + n * 8 + 2 * t;
                    uint2 u0, u1;
                    u0.x = to_tf32(dp[n][0]); u0.y = to_tf32(dp[n][1]);
                    u1.x = to_tf32(dp[n][2]); u1.y = to_tf32(dp[n][3]);
                    *reinterpret_cast<uint2*>(w0) = u0;
                    *reinterpret_cast<uint2*>(w1) = u1;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) dp[n][e] = 0.f;
                }
            }
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += C::DC) {
                float dq[C::MT][4];
                tc_regs_dot_cols<D, LP>(dp, Ks, c0, len, dq, g, t);      // dQ = dS . K
                tc_store_stripe<T, C::DC>(dQ, p.ld, row0, len, colh + c0, m0, dq, g, t);
            }
        }
        __syncthreads();
        if (m0 < len) {                                    // ---- key-stripe phase (j0 = m0)
            // query rows read below: [0, 8*ceil(len/8)); rows >= len inside an active stripe hold P = dS = 0
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += C::DC) {
                float acc[C::MT][4];
                tc_smT_dot_cols<D, LP>(dSm, Qs, c0, m0, len, acc, g, t); // dK = dS^T . Q
                tc_store_stripe<T, C::DC>(dK, p.ld, row0, len, colh + c0, m0, acc, g, t);
                tc_smT_dot_cols<D, LP>(Pm, Gs, c0, m0, len, acc, g, t);  // dV = P~^T . dO
                tc_store_stripe<T, C::DC>(dV, p.ld, row0, len, colh + c0, m0, acc, g, t);
            }
        }
    }
    if (p.dbias) {
        __syncthreads();
        float* gb = p.dbias + (size_t)h * LL;
        for (int i = threadIdx.x; i < LL; i += C::THREADS)
            if (dB[i] != 0.f) atomicAdd(gb + i, dB[i]);
    }
}

template <typename T, int D, int LP>
static int tc_launch(const TcAttnParams& p, bool bwd, cudaStream_t stream) {
    using C = TcCfg<D, LP>;
    size_t smem = (size_t)(bwd ? 4 * C::TILE + 2 * LP * C::SS : 3 * C::TILE) * sizeof(float);
    if (bwd && p.dbias) smem += (size_t)p.seqlen * p.seqlen * sizeof(float);
    int gx = p.n_seq;
    if (bwd && p.dbias) {                       // few CTAs per head: each flushes its bias-gradient tile once
        gx = (num_sms() * 3 + p.n_heads - 1) / p.n_heads;
        if (gx > p.n_seq) gx = p.n_seq;
    }
    if (gx < 1) gx = 1;
    dim3 grid(gx, p.n_heads);
    if (!bwd) {
        auto kern = attn_tc_fwd_kernel<T, D, LP>;
        static size_t set_f = 0;                         // per instantiation: raise the opt-in limit only when it grows
        if (smem > set_f) {
            MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_f = smem;
        }
        kern<<<grid, C::THREADS, smem, stream>>>(p);
    } else {
        auto kern = attn_tc_bwd_kernel<T, D, LP>;
        static size_t set_b = 0;
        if (smem > set_b) {
            MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_b = smem;
        }
        kern<<<grid, C::THREADS, smem, stream>>>(p);
    }
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

// true when the tensor-core kernels cover this problem (fast modes only: C-ABI dtype 0 or 1)
inline bool tc_attn_eligible(int dtype, int seqlen, int head_dim, int ld, int ld_o) {
    if (!((dtype == 0 || dtype == 1 || dtype == 3) && seqlen > 0 && ld % 4 == 0 && ld_o % 4 == 0)) return false;
    if (head_dim == 32 || head_dim == 64) return seqlen <= 64;
    return head_dim == 256 && seqlen <= 32;          // SASRec user tower: D = 512, 2 heads (parameters.py:28)
}

int tc_attn_dispatch(const TcAttnParams& p, bool bwd, int dtype, cudaStream_t stream);

}  // namespace morec
