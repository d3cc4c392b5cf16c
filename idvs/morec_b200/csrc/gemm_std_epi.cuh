// Standard epilogue family of the tcgen05 GEMM: bias / GELU / ReLU / activation-gradient, specialised at COMPILE time
// on the epilogue mode (and on the math kind for the erf flavour).  One kernel per mode keeps the epilogue loop a few
// hundred instructions long; the former runtime `switch` over all modes produced a 35,000-instruction kernel whose
// GELU paths ran out of the instruction cache (measured: 176 us for the GELU forward and 163 us for the GELU-gradient
// dgrad against 54-61 us for the same GEMM with a plain epilogue).
#pragma once
#include "gemm2_tcgen05.cuh"

#include "../../../include/morec_b200.h"

namespace morec {

struct StdEpiParams {
    int mode;           // MOREC_EPI_* (informational: the kernel is specialised on it)
    float alpha;
    const float* bias;   // [N] fp32 or null
    const void* aux;     // [M, ldaux] or null (dtype = aux_bf16 ? bf16 : fp32)
    int ldaux;
    int aux_bf16;        // side input is 16-bit
    int aux_f16;         // ... and IEEE half rather than bfloat16
};

template <int MODE>
struct StdEpi {
    using Params = StdEpiParams;
    static constexpr bool kAuxMode = MODE == MOREC_EPI_MUL_GELU_GRAD || MODE == MOREC_EPI_MUL_RELU_GRAD || MODE == MOREC_EPI_MUL_AUX;
    static constexpr int kStreams = (MODE == MOREC_EPI_GELU || MODE == MOREC_EPI_GELU_DGELU) ? 2 : 1;
    static constexpr int kGroups = 2;      // CTA-pair kernel: two epilogue warp groups, half of the tile's columns each

    __device__ __forceinline__ static void load_aux(const Params& ep, float (&a)[32], int row, int col0, int M, int N) {
        if (row < M) {
            if (!ep.aux_bf16) {
                const float* ap = reinterpret_cast<const float*>(ep.aux) + (size_t)row * ep.ldaux + col0;
                if (col0 + 32 <= N) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(ap) + j);
                        a[4 * j] = t.x; a[4 * j + 1] = t.y; a[4 * j + 2] = t.z; a[4 * j + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (col0 + j < N) ? __ldg(ap + j) : 0.f;
                }
            } else {
                const uint16_t* ap = reinterpret_cast<const uint16_t*>(ep.aux) + (size_t)row * ep.ldaux + col0;
                const bool f16 = ep.aux_f16 != 0;
                if (col0 + 32 <= N && (ep.ldaux & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.aux) & 15) == 0) {
                    // 4 x 16-byte loads per row (the scalar form issued 32 two-byte loads, each touching 32 sectors)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 u = __ldg(reinterpret_cast<const uint4*>(ap) + j);
                        const uint32_t* h = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2_16(h[e], f16);
                            a[8 * j + 2 * e] = f.x;
                            a[8 * j + 2 * e + 1] = f.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (col0 + j < N) ? unpack1_16(__ldg(ap + j), f16) : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = 0.f;
        }
    }

    // bf16 side input of one 32-column chunk, kept PACKED (4 x 16 B): issued one chunk ahead of its use so the load
    // latency overlaps the previous chunk's arithmetic (each epilogue warp is alone on its scheduler)
    __device__ __forceinline__ static bool aux_raw_ok(const Params& ep, const TileSched& s) {
        return kAuxMode && ep.aux != nullptr && ep.aux_bf16 && (ep.ldaux & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.aux) & 15) == 0 && (s.N & 31) == 0;
    }
    __device__ __forceinline__ static void load_aux_raw(const Params& ep, uint4 (&r)[4], int row, int col0, int M) {
        if (row < M) {
            const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.aux) +
                                                             (size_t)row * ep.ldaux + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = __ldg(ap + j);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // unpack the 32 packed 16-bit side-input values of a chunk (F16: IEEE half, else bfloat16)
    template <bool F16>
    __device__ __forceinline__ static void aux_unpack(const uint4 (&r)[4], float (&a)[32]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(&r[j >> 2])[j & 3];
            const float2 f = unpack2_16(w, F16);
            a[2 * j] = f.x;
            a[2 * j + 1] = f.y;
        }
    }

    // this warp's chunk range [c_lo, c_end) of the tile at n0 (column group cg of ncg)
    template <int BLOCK_N>
    __device__ __forceinline__ static void chunk_range(const TileSched& s, int n0, int cg, int ncg, int& c_lo, int& c_end) {
        c_end = (s.N - n0 + 31) / 32;
        if (c_end > BLOCK_N / 32) c_end = BLOCK_N / 32;
        if (s.out_bf16) c_end = (c_end + 1) & ~1;
        c_lo = cg * (BLOCK_N / 32) / ncg;
        const int c_hi = (cg + 1) * (BLOCK_N / 32) / ncg;
        if (c_end > c_hi) c_end = c_hi;
    }

    // Side input by TMA (CTA-pair kernel, bf16): before the wait for the tile's accumulator, stage this warp's
    // 32 rows x (its columns) of aux into its own staging sub-buffers -- the very boxes (and swizzle) the result will be
    // stored from, so the epilogue reads aux from shared memory and overwrites it in place.  Per-thread global loads
    // of a row-per-thread layout touched 32 cache lines per instruction and were exposed to their full latency
    // (measured: 93 us for dgrad * aux against 52 us for the plain dgrad).
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void pre_tile(const Params& ep, const CUtensorMap& tmAux, EpiStore& st, int m0, int q,
                                                    int n0, const TileSched& s, int cg, int ncg) {
        st.aux_groups = 0;
        if constexpr (kAuxMode) {
            if (!s.aux_tma || st.aux_bar == 0) return;
            const int row0 = m0 + q * 32;
            if (row0 >= s.M) return;
            int c_lo, c_end;
            chunk_range<BLOCK_N>(s, n0, cg, ncg, c_lo, c_end);
            const int ng = (c_end - c_lo + 1) / 2;                // 64-column (two-chunk) groups = sub-buffers
            if (ng <= 0 || ng > st.nsub) return;
            if (st.lane == 0) {
                tma_store_wait_read<0>();                         // both staging halves are free again
                mbar_expect_tx(st.aux_bar, (uint32_t)ng * kEpiBufBytes);
                for (int i = 0; i < ng; ++i)
                    tma_load_2d(&tmAux, st.aux_bar, smem_u32(st.bufs + ((st.grp + i) & 1) * kEpiBufBytes),
                                n0 + (c_lo + 2 * i) * 32, row0);
            }
            __syncwarp();
            st.aux_groups = ng;
        }
    }
    // packed side input of chunk c from the staging sub-buffer the result of this chunk will be written to
    __device__ __forceinline__ static void load_aux_smem(const EpiStore& st, uint4 (&r)[4], int c) {
        const uint8_t* rowp = st.bufs + (st.grp & 1) * kEpiBufBytes + st.lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = *reinterpret_cast<const uint4*>(rowp + ((((c & 1) << 2) + j) ^ (st.lane & 7)) * 16);
    }

    // one tile ahead: pull this thread's row of the side input (aux) for columns [n0, n0 + BLOCK_N) into L2
    template <int BLOCK_N>
    __device__ __forceinline__ static void prefetch(const Params& ep, int row, int n0, const TileSched& s) {
        if (!kAuxMode || ep.aux == nullptr || row >= s.M) return;
        const int es = ep.aux_bf16 ? 2 : 4;
        const char* base = reinterpret_cast<const char*>(ep.aux) + ((size_t)row * ep.ldaux + n0) * es;
        int bytes = (s.N - n0 < BLOCK_N ? s.N - n0 : BLOCK_N) * es;
        for (int off = 0; off < bytes; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }

    // One 32-column chunk of this thread's row: activation / gradient math on x, then staging + (grouped) TMA store.
    template <bool FAST, bool O16, bool F16>
    __device__ __forceinline__ static void chunk(const Params& ep, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                                                 EpiStore& st, float (&x)[32], int c, int row, int row0, int n0,
                                                 const TileSched& s, const uint4 (&araw)[4], bool have_raw) {
        const int col0 = n0 + c * 32;
        constexpr int ns = kStreams;
        if constexpr (MODE == MOREC_EPI_GELU) {
            // pre-activation to C2 first (kept for the backward), then the activation to C
            st.template put_c<ns, O16, F16>(&tmC2, x, c, 1);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = FAST ? gelu_fast(x[j]) : gelu_erf(x[j]);
        } else if constexpr (MODE == MOREC_EPI_GELU_DGELU) {
            // activation derivative to C2 (the backward multiplies by it), activation to C; both from one erfc / exp
            float d[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if constexpr (FAST) {
                    float q, e;
                    gelu_fast_parts(x[j], q, e);
                    const float cdf = x[j] < 0.f ? q : 1.0f - q;
                    d[j] = fmaf(x[j], 0.39894228040143267794f * e, cdf);
                    x[j] = fmaf(-fabsf(x[j]), q, fmaxf(x[j], 0.f));
                } else {
                    d[j] = gelu_erf_grad(x[j]);
                    x[j] = gelu_erf(x[j]);
                }
            }
            st.template put_c<ns, O16, F16>(&tmC2, d, c, 1);
        } else if constexpr (MODE == MOREC_EPI_GELU_NOSAVE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = FAST ? gelu_fast(x[j]) : gelu_erf(x[j]);
        } else if constexpr (MODE == MOREC_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        } else if constexpr (kAuxMode) {
            float a[32];
            if (have_raw) { if (ep.aux_f16) aux_unpack<true>(araw, a); else aux_unpack<false>(araw, a); }
            else load_aux(ep, a, row, col0, s.M, s.N);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if constexpr (MODE == MOREC_EPI_MUL_AUX) x[j] *= a[j];
                else if constexpr (MODE == MOREC_EPI_MUL_GELU_GRAD) x[j] *= FAST ? gelu_fast_grad(a[j]) : gelu_erf_grad(a[j]);
                else x[j] = a[j] > 0.f ? x[j] : 0.f;
            }
        }
        st.template put_c<ns, O16, F16>(&tmC, x, c, 0);
        st.template end_chunk_c<ns, O16>(c, n0, row0, s.accumulate != 0);
    }

    // KIND 2 (3xTF32, the parity mode) keeps erff; the fast modes use the polynomial erf (common.cuh).
    // Specialised on the output element type at the tile level (the flags are warp-uniform kernel parameters).
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void tile(const Params& ep, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                                                uint32_t taddr, EpiStore& st, int m0, int q, int n0, int split,
                                                const TileSched& s, int cg, int ncg) {
        if (s.out_bf16) {
            if (st.f16) tile_t<KIND, BLOCK_N, true, true>(ep, tmC, tmC2, taddr, st, m0, q, n0, split, s, cg, ncg);
            else tile_t<KIND, BLOCK_N, true, false>(ep, tmC, tmC2, taddr, st, m0, q, n0, split, s, cg, ncg);
        } else {
            tile_t<KIND, BLOCK_N, false, false>(ep, tmC, tmC2, taddr, st, m0, q, n0, split, s, cg, ncg);
        }
    }

    template <int KIND, int BLOCK_N, bool O16, bool F16>
    __device__ __forceinline__ static void tile_t(const Params& ep, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                                                  uint32_t taddr, EpiStore& st, int m0, int q, int n0, int split,
                                                  const TileSched& s, int cg, int ncg) {
        constexpr bool FAST = KIND != 2;
        const int row0 = m0 + q * 32;
        if (row0 >= s.M) return;   // warp-uniform
        const int row = row0 + st.lane;
        int c_lo, c_end;
        chunk_range<BLOCK_N>(s, n0, cg, ncg, c_lo, c_end);
        if (c_lo >= c_end) return;
        st.c_end = c_end;
        // bias: lane j holds the bias of column j of the chunk (ONE load per lane and chunk, fetched a chunk ahead and
        // broadcast by shuffle) -- the former 8 x LDG.128 per thread and chunk sat on the critical path of every chunk
        // (ncu: 10 % of the GELU epilogue's samples were long-scoreboard stalls on the first bias add)
        const float* bias = (ep.bias != nullptr && split == 0) ? ep.bias : nullptr;
        auto bias_of = [&](int c) -> float {
            const int col = n0 + c * 32 + st.lane;
            return (bias != nullptr && col < s.N) ? __ldg(bias + col) : 0.f;
        };
        const bool tma_aux = st.aux_groups > 0;
        if (tma_aux) {                                  // side input staged by pre_tile(): wait for it to land
            mbar_wait(st.aux_bar, st.aux_phase);
            st.aux_phase ^= 1;
        }
        const bool direct = !tma_aux && aux_raw_ok(ep, s);
        const bool raw = tma_aux || direct;
        uint4 ra[4], rb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ra[j] = rb[j] = make_uint4(0u, 0u, 0u, 0u);
        uint32_t v[32];
        if (direct) load_aux_raw(ep, ra, row, n0 + c_lo * 32, s.M);
        tmem_ld32(taddr + c_lo * 32, v);
        float b_next = bias_of(c_lo);
        // software pipeline with ONE copy of the chunk code: the accumulator registers are consumed (scaled, biased)
        // first, then the TMEM load of chunk c+1 is issued into the same registers and is in flight during the math
#pragma unroll 1
        for (int c = c_lo; c < c_end; ++c) {
            tc_wait_ld();
            const float bc = b_next;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaf(__uint_as_float(v[j]), ep.alpha, __shfl_sync(0xffffffffu, bc, j));
            if (c + 1 < c_end) {
                tmem_ld32(taddr + (c + 1) * 32, v);
                b_next = bias_of(c + 1);
                if (direct) load_aux_raw(ep, rb, row, n0 + (c + 1) * 32, s.M);
            }
            if (tma_aux) load_aux_smem(st, ra, c);
            chunk<FAST, O16, F16>(ep, tmC, tmC2, st, x, c, row, row0, n0, s, ra, raw);
            if (direct) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ra[j] = rb[j];
            }
        }
    }
};

// one translation unit per epilogue mode (gemm_std_m<MODE>.cu) instantiates the kernels of that mode
#define MOREC_DECLARE_STD_GEMM(MODE) int gemm_std_run_##MODE(const GemmArgs& g, const StdEpiParams& ep, cudaStream_t stream);
MOREC_DECLARE_STD_GEMM(0)
MOREC_DECLARE_STD_GEMM(1)
MOREC_DECLARE_STD_GEMM(2)
MOREC_DECLARE_STD_GEMM(3)
MOREC_DECLARE_STD_GEMM(4)
MOREC_DECLARE_STD_GEMM(5)
MOREC_DECLARE_STD_GEMM(6)
MOREC_DECLARE_STD_GEMM(7)
#define MOREC_DEFINE_STD_GEMM(MODE)                                                                     \
    namespace morec {                                                                                   \
    int gemm_std_run_##MODE(const GemmArgs& g, const StdEpiParams& ep, cudaStream_t stream) {           \
        return gemm_dispatch_auto<StdEpi<MODE>>(g, ep, stream);                                         \
    }                                                                                                   \
    }

}  // namespace morec
