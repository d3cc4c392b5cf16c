// bf16-native tensor-core attention (mma.sync.m16n8k16, fp32 accumulate) for the "bf16" mode: <= 64 tokens per
// sequence / window, head_dim 32 or 64.  Same decomposition as attention_tc.cuh (one CTA of LP/16 warps per (sequence,
// head), warp w owns a 16-query stripe, softmax on the accumulator fragments, P / dS fed to the next product from
// registers) with the data movement rebuilt around 16-bit operands:
//   * Q, K, V (, dO) go global -> shared with 16-byte cp.async (zero-fill for rows >= len): no register staging, no
//     conversion instructions; tiles are bf16 with row stride D+8 elements (ldmatrix-conflict-free);
//   * every operand fragment is ONE ldmatrix.x4 (plain for row operands, .trans for the "k-row" operands V / K / Q / dO
//     of the P.V-shaped products and for the transposed reads of dS / P~ in the key-stripe phase) instead of 4-6 scalar
//     LDS: the TF32 kernel issued ~1900 instructions per warp per pair and ran at 38 % issue utilisation with 14 warps
//     per SM (ncu); this one needs half the mma count, ~1/8 of the shared loads and half the shared memory;
//   * P~ and dS are rounded to bf16 for the second product (as in FlashAttention); the softmax, its backward and the
//     accumulators are fp32.
#pragma once
#include "attention_tc.cuh"

namespace morec {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
// T = __nv_bfloat16 ("bf16" mode) or __half ("fp16" mode: the reference's own autocast arithmetic, run.py:242)
template <typename T>
__device__ __forceinline__ void mma16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (sizeof(T) == 2 && !__is_same(T, __half)) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}
template <typename T>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    return pack2_16(lo, hi, __is_same(T, __half));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int D, int LP>
struct Tc16Cfg {
    static constexpr int NW = LP / 16;
    static constexpr int THREADS = 32 * NW;
    static constexpr int SB = D + 8;            // operand tile row stride (bf16 elements): rows 16-byte aligned, 8 rows
                                                // of an ldmatrix 8x8 land in 8 distinct 16-byte bank groups
    static constexpr int SSB = LP + 8;          // score tile row stride (bf16 elements)
    static constexpr int NT = LP / 8;
    static constexpr int MT = D / 8;
    static constexpr int TILE = LP * SB;        // elements
};

// rows [row0, row0+len) x cols [col0, col0+D) -> dst[LP][SB] (bf16), rows >= len zero-filled; asynchronous
template <typename T, int D, int LP>
__device__ __forceinline__ void tc16_load_tile(T* dst, const T* base, int ld, int row0, int len,
                                               int col0) {
    using C = Tc16Cfg<D, LP>;
    constexpr int NV = D / 8;                                    // 16-byte chunks per row
    constexpr int ITER = LP * NV / C::THREADS;
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int idx = threadIdx.x + i * C::THREADS;
        const int r = idx / NV, c = (idx - r * NV) << 3;
        const bool ok = r < len;
        const T* src = ok ? base + (size_t)(row0 + r) * ld + col0 + c : base;
        cp_async16(smem_u32(dst + r * C::SB + c), src, ok ? 16u : 0u);
    }
}

// acc[n] (n < NT) = X[m0 .. m0+16, :] . Y[8n .. 8n+8, :]^T over the D columns (both operands stored by rows)
template <typename T, int D, int LP>
__device__ __forceinline__ void tc16_rows_dot_rows(const T* X, const T* Y, int m0, int ncols,
                                                   float (&acc)[LP / 8][4], int lane) {
    using C = Tc16Cfg<D, LP>;
#pragma unroll
    for (int n = 0; n < C::NT; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
    const uint32_t xa = smem_u32(X + (m0 + (lane & 15)) * C::SB + ((lane >> 4) << 3));
    const uint32_t ya = smem_u32(Y + ((lane & 7) + ((lane >> 4) << 3)) * C::SB + (((lane >> 3) & 1) << 3));
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t a[4];
        ldsm_x4(a, xa + kk * 32);
#pragma unroll
        for (int j = 0; j < C::NT / 2; ++j) {
            if (j * 16 < ncols) {
                uint32_t b[4];
                ldsm_x4(b, ya + (j * 16 * C::SB + kk * 16) * 2);
                mma16<T>(acc[2 * j], a, b[0], b[1]);
                mma16<T>(acc[2 * j + 1], a, b[2], b[3]);
            }
        }
    }
}

// out[m] (m < MT) = sum over 16-key steps of  A (accumulator registers, rounded to bf16) . Y[16ks .., 8m ..]
template <typename T, int D, int LP>
__device__ __forceinline__ void tc16_regs_dot_cols(const float (&p)[LP / 8][4], const T* Y, int nrows,
                                                   float (&out)[D / 8][4], int lane) {
    using C = Tc16Cfg<D, LP>;
#pragma unroll
    for (int m = 0; m < C::MT; ++m) { out[m][0] = out[m][1] = out[m][2] = out[m][3] = 0.f; }
    const uint32_t ya = smem_u32(Y + ((lane & 7) + (((lane >> 3) & 1) << 3)) * C::SB + ((lane >> 4) << 3));
#pragma unroll
    for (int ks = 0; ks < LP / 16; ++ks) {
        if (ks * 16 < nrows) {
            uint32_t a[4];
            a[0] = pack16<T>(p[2 * ks][0], p[2 * ks][1]);
            a[1] = pack16<T>(p[2 * ks][2], p[2 * ks][3]);
            a[2] = pack16<T>(p[2 * ks + 1][0], p[2 * ks + 1][1]);
            a[3] = pack16<T>(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
            for (int mp = 0; mp < C::MT / 2; ++mp) {
                uint32_t b[4];
                ldsm_x4_t(b, ya + (ks * 16 * C::SB + mp * 16) * 2);
                mma16<T>(out[2 * mp], a, b[0], b[1]);
                mma16<T>(out[2 * mp + 1], a, b[2], b[3]);
            }
        }
    }
}

// out[m] = sum over 16-query steps of  Sm^T[j0 .. j0+16, 16ks ..] . Y[16ks .., 8m ..]     (Sm: [LP][SSB] bf16)
template <typename T, int D, int LP>
__device__ __forceinline__ void tc16_smT_dot_cols(const T* Sm, const T* Y, int j0, int nrows,
                                                  float (&out)[D / 8][4], int lane) {
    using C = Tc16Cfg<D, LP>;
#pragma unroll
    for (int m = 0; m < C::MT; ++m) { out[m][0] = out[m][1] = out[m][2] = out[m][3] = 0.f; }
    const uint32_t sa = smem_u32(Sm + ((lane & 7) + (((lane >> 4) & 1) << 3)) * C::SSB + j0 + (((lane >> 3) & 1) << 3));
    const uint32_t ya = smem_u32(Y + ((lane & 7) + (((lane >> 3) & 1) << 3)) * C::SB + ((lane >> 4) << 3));
#pragma unroll
    for (int ks = 0; ks < LP / 16; ++ks) {
        if (ks * 16 < nrows) {
            uint32_t a[4];
            ldsm_x4_t(a, sa + (ks * 16 * C::SSB) * 2);
#pragma unroll
            for (int mp = 0; mp < C::MT / 2; ++mp) {
                uint32_t b[4];
                ldsm_x4_t(b, ya + (ks * 16 * C::SB + mp * 16) * 2);
                mma16<T>(out[2 * mp], a, b[0], b[1]);
                mma16<T>(out[2 * mp + 1], a, b[2], b[3]);
            }
        }
    }
}

template <typename T, int D, int LP>
__global__ void __launch_bounds__(Tc16Cfg<D, LP>::THREADS) attn_tc16_fwd_kernel(const TcAttnParams p) {
    pdl_wait();
    pdl_trigger();
    using C = Tc16Cfg<D, LP>;
    extern __shared__ __align__(16) uint8_t tc16_smem[];
    T* Qs = reinterpret_cast<T*>(tc16_smem);
    T* Ks = Qs + C::TILE;
    T* Vs = Ks + C::TILE;
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
        tc16_load_tile<T, D, LP>(Qs, Q, p.ld, row0, len, colh);
        tc16_load_tile<T, D, LP>(Ks, K, p.ld, row0, len, colh);
        tc16_load_tile<T, D, LP>(Vs, V, p.ld, row0, len, colh);
        cp_async_wait_all();
        __syncthreads();
        if (m0 >= len) continue;                           // stripe of padding rows (warp-uniform)
        float acc[C::NT][4];
        tc16_rows_dot_rows<T, D, LP>(Qs, Ks, m0, len, acc, lane);
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
        float o[C::MT][4];
        tc16_regs_dot_cols<T, D, LP>(acc, Vs, len, o, lane);
        tc_store_stripe<T, D>(O, p.ld_o, row0, len, colh, m0, o, g, t);
    }
}

template <typename T, int D, int LP>
__global__ void __launch_bounds__(Tc16Cfg<D, LP>::THREADS) attn_tc16_bwd_kernel(const TcAttnParams p) {
    pdl_wait();
    pdl_trigger();
    using C = Tc16Cfg<D, LP>;
    extern __shared__ __align__(16) uint8_t tc16_smem[];
    T* Qs = reinterpret_cast<T*>(tc16_smem);
    T* Ks = Qs + C::TILE;
    T* Vs = Ks + C::TILE;
    T* Gs = Vs + C::TILE;                      // dO
    T* Pm = Gs + C::TILE;                      // [LP][SSB]  P~ (dropped, rescaled)
    T* dSm = Pm + LP * C::SSB;                 // [LP][SSB]  dS * scale
    float* dB = reinterpret_cast<float*>(dSm + LP * C::SSB);   // [seqlen*seqlen] bias-gradient accumulator (if dbias)
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
        tc16_load_tile<T, D, LP>(Qs, Q, p.ld, row0, len, colh);
        tc16_load_tile<T, D, LP>(Ks, K, p.ld, row0, len, colh);
        tc16_load_tile<T, D, LP>(Vs, V, p.ld, row0, len, colh);
        tc16_load_tile<T, D, LP>(Gs, dO, p.ld_o, row0, len, colh);
        cp_async_wait_all();
        __syncthreads();
        if (m0 < len) {                                    // ---- query-stripe phase
            float pr[C::NT][4], dp[C::NT][4];
            tc16_rows_dot_rows<T, D, LP>(Qs, Ks, m0, len, pr, lane);
            tc_softmax_stripe<LP>(p, pr, s, h, m0, len, g, t);
            tc16_rows_dot_rows<T, D, LP>(Gs, Vs, m0, len, dp, lane);        // dP~ = dO . V^T
            float keepm[C::NT][4];
            tc_keep_stripe<C::NT>(p, s * p.n_heads + h, C::NW, warp, lane, len, keepm);
            float dsum[2] = {0.f, 0.f};
#pragma unroll
            for (int n = 0; n < C::NT; ++n) {
                const float (&keep)[4] = keepm[n];
                if (n * 8 < len) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        dp[n][e] *= keep[e];                             // dP
                        dsum[e >> 1] += pr[n][e] * dp[n][e];
                    }
                }
                // P~ for dV (zero beyond len: pr is exactly 0 there); every 16-key block the key phase reads is written
                *reinterpret_cast<uint32_t*>(Pm + (m0 + g) * C::SSB + n * 8 + 2 * t) =
                    pack16<T>(pr[n][0] * keep[0], pr[n][1] * keep[1]);
                *reinterpret_cast<uint32_t*>(Pm + (m0 + g + 8) * C::SSB + n * 8 + 2 * t) =
                    pack16<T>(pr[n][2] * keep[2], pr[n][3] * keep[3]);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
                dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
            }
#pragma unroll
            for (int n = 0; n < C::NT; ++n) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float ds = pr[n][e] * (dp[n][e] - dsum[e >> 1]);         // gradient of the pre-softmax score
                    if (p.dbias) {
                        const int row = m0 + g + ((e >> 1) << 3), col = n * 8 + 2 * t + (e & 1);
                        if (row < len && col < len) dB[row * p.seqlen + col] += ds;
                    }
                    dp[n][e] = ds * p.scale;
                }
                *reinterpret_cast<uint32_t*>(dSm + (m0 + g) * C::SSB + n * 8 + 2 * t) = pack16<T>(dp[n][0], dp[n][1]);
                *reinterpret_cast<uint32_t*>(dSm + (m0 + g + 8) * C::SSB + n * 8 + 2 * t) = pack16<T>(dp[n][2], dp[n][3]);
            }
            float dq[C::MT][4];
            tc16_regs_dot_cols<T, D, LP>(dp, Ks, len, dq, lane);            // dQ = dS . K
            tc_store_stripe<T, D>(dQ, p.ld, row0, len, colh, m0, dq, g, t);
        }
        __syncthreads();
        if (m0 < len) {                                    // ---- key-stripe phase (j0 = m0)
            // query rows read: [0, 16*ceil(len/16)), all inside stripes that ran the phase above; rows >= len hold 0
            float acc[C::MT][4];
            tc16_smT_dot_cols<T, D, LP>(dSm, Qs, m0, len, acc, lane);       // dK = dS^T . Q
            tc_store_stripe<T, D>(dK, p.ld, row0, len, colh, m0, acc, g, t);
            tc16_smT_dot_cols<T, D, LP>(Pm, Gs, m0, len, acc, lane);        // dV = P~^T . dO
            tc_store_stripe<T, D>(dV, p.ld, row0, len, colh, m0, acc, g, t);
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
static int tc16_launch(const TcAttnParams& p, bool bwd, cudaStream_t stream) {
    using C = Tc16Cfg<D, LP>;
    size_t smem = (size_t)(bwd ? 4 * C::TILE + 2 * LP * C::SSB : 3 * C::TILE) * sizeof(T);
    if (bwd && p.dbias) smem += (size_t)p.seqlen * p.seqlen * sizeof(float);
    int gx = p.n_seq;
    if (bwd && p.dbias) {                       // few CTAs per head: each flushes its bias-gradient tile once
        gx = (num_sms() * 4 + p.n_heads - 1) / p.n_heads;
        if (gx > p.n_seq) gx = p.n_seq;
    }
    if (gx < 1) gx = 1;
    dim3 grid(gx, p.n_heads);
    if (!bwd) {
        auto kern = attn_tc16_fwd_kernel<T, D, LP>;
        static size_t set_f = 0;                         // per instantiation: raise the opt-in limit only when it grows
        if (smem > set_f) {
            MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_f = smem;
        }
        MOREC_CUDA(launch_pdl(kern, grid, dim3(C::THREADS), smem, stream, p));
    } else {
        auto kern = attn_tc16_bwd_kernel<T, D, LP>;
        static size_t set_b = 0;
        if (smem > set_b) {
            MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_b = smem;
        }
        MOREC_CUDA(launch_pdl(kern, grid, dim3(C::THREADS), smem, stream, p));
    }
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

// bf16 storage, 32/64-wide heads, 16-byte aligned rows: the cp.async / ldmatrix kernels apply
inline bool tc16_eligible(const TcAttnParams& p, int dtype, bool bwd) {
    if (!MOREC_DT_IS16(dtype) || !(p.head_dim == 32 || p.head_dim == 64) || p.seqlen > 64) return false;
    if (p.ld % 8 || p.ld_o % 8) return false;
    auto al = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
    if (!al(p.q) || !al(p.k) || !al(p.v)) return false;
    if (bwd && !al(p.o)) return false;
    return true;
}

}  // namespace morec
