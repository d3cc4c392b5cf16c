// Standard epilogue family of the tcgen05 GEMM + the C-ABI entry point morec_gemm (see include/morec_b200.h).
#include "gemm2_tcgen05.cuh"

#include <mutex>

#include "../../../include/morec_b200.h"

namespace morec {

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is resolved at run time through cudart: the library links only libcudart, so it loads
// (and exports every symbol) on a CPU-only box and fails loudly at the first compute call instead.
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
    (void)cudaGetLastError();
}

int make_tmap_2d(CUtensorMap* map, const void* base, bool is_bf16, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle32) {
    std::call_once(g_encode_once, resolve_encode);
    if (!g_encode) {
        set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver): the CUDA path cannot run here");
        return MOREC_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15)) {
        set_last_error("TMA operand must be 16-byte aligned (base %p, pitch %llu)", base,
                       (unsigned long long)row_pitch_bytes);
        return MOREC_ERR_ARG;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = g_encode(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed rc=%d (inner=%llu outer=%llu pitch=%llu box=%ux%u)", (int)r,
                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_pitch_bytes,
                       box_inner, box_outer);
        return MOREC_ERR_CUDA;
    }
    return MOREC_OK;
}

// ------------------------------------------------------------------------------------------------
// standard epilogues
// ------------------------------------------------------------------------------------------------
struct StdEpi {
    struct Params {
        int mode;
        float alpha;
        const float* bias;   // [N] fp32 or null
        const void* aux;     // [M, ldaux] or null (dtype = aux_bf16 ? bf16 : fp32)
        int ldaux;
        int aux_bf16;
        int fast;            // 1: polynomial erf (|err| <= 1.5e-7) in the GELU epilogues (fast modes); 0: erff (parity mode)
    };

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
                const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)row * ep.ldaux + col0;
                if (col0 + 32 <= N && (ep.ldaux & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.aux) & 15) == 0) {
                    // 4 x 16-byte loads per row (the scalar form issued 32 two-byte loads, each touching 32 sectors)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 u = __ldg(reinterpret_cast<const uint4*>(ap) + j);
                        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            a[8 * j + 2 * e] = __low2float(h[e]);
                            a[8 * j + 2 * e + 1] = __high2float(h[e]);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (col0 + j < N) ? __bfloat162float(ap[j]) : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = 0.f;
        }
    }

    __device__ __forceinline__ static void chunk(const Params& ep, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                                                 EpiStore& st, const uint32_t (&v)[32], int c, int row, int row0,
                                                 int n0, bool add_bias, const TileSched& s) {
        const int col0 = n0 + c * 32;
        const bool obf = s.out_bf16 != 0;
        const int ns = ep.mode == MOREC_EPI_GELU ? 2 : 1;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * ep.alpha;
        if (add_bias) {
            if (col0 + 32 <= s.N) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j);
                    x[4 * j] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < s.N) x[j] += __ldg(ep.bias + col0 + j);
            }
        }
        switch (ep.mode) {
            case MOREC_EPI_GELU: {
                // pre-activation to C2 first (kept for the backward), then the activation to C
                st.put(&tmC2, x, c, obf, 1, 2);
                if (ep.fast) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_fast(x[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
                }
                break;
            }
            case MOREC_EPI_GELU_NOSAVE: {
                if (ep.fast) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_fast(x[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
                }
                break;
            }
            case MOREC_EPI_RELU: {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                break;
            }
            case MOREC_EPI_MUL_GELU_GRAD: {
                float a[32];
                load_aux(ep, a, row, col0, s.M, s.N);
                if (ep.fast) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] *= gelu_fast_grad(a[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] *= gelu_erf_grad(a[j]);
                }
                break;
            }
            case MOREC_EPI_MUL_RELU_GRAD: {
                float a[32];
                load_aux(ep, a, row, col0, s.M, s.N);
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = a[j] > 0.f ? x[j] : 0.f;
                break;
            }
            default:
                break;
        }
        st.put(&tmC, x, c, obf, 0, ns);
        st.end_chunk(c, n0, row0, obf, s.accumulate != 0, ns);
    }

    template <int BLOCK_N>
    __device__ __forceinline__ static void tile(const Params& ep, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                                                uint32_t taddr, EpiStore& st, int m0, int q, int n0, int split,
                                                const TileSched& s) {
        const int row0 = m0 + q * 32;
        if (row0 >= s.M) return;   // warp-uniform
        const int row = row0 + st.lane;
        int c_end = (s.N - n0 + 31) / 32;
        if (c_end > BLOCK_N / 32) c_end = BLOCK_N / 32;
        if (s.out_bf16) c_end = (c_end + 1) & ~1;
        st.c_end = c_end;
        const bool add_bias = ep.bias != nullptr && split == 0;
        // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is processed
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < c_end; c += 2) {
            tc_wait_ld();
            if (c + 1 < c_end) tmem_ld32(taddr + (c + 1) * 32, vb);
            chunk(ep, tmC, tmC2, st, va, c, row, row0, n0, add_bias, s);
            if (c + 1 < c_end) {
                tc_wait_ld();
                if (c + 2 < c_end) tmem_ld32(taddr + (c + 2) * 32, va);
                chunk(ep, tmC, tmC2, st, vb, c + 1, row, row0, n0, add_bias, s);
            }
        }
    }
};

}  // namespace morec

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int morec_gemm(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux,
                          int M, int N, int K, int lda, int ldb, int ldc, int ldaux, int a_mn_major, int b_mn_major,
                          int dtype, int out_bf16, int epilogue, float alpha, int accumulate, void* stream) {
    using namespace morec;
    MOREC_CHECK_ARG(A && B && C, "morec_gemm: null operand");
    MOREC_CHECK_ARG(epilogue >= MOREC_EPI_LINEAR && epilogue <= MOREC_EPI_MUL_RELU_GRAD, "morec_gemm: bad epilogue %d",
                    epilogue);
    MOREC_CHECK_ARG(epilogue != MOREC_EPI_GELU || C2, "morec_gemm: EPI_GELU needs C2 (pre-activation output)");
    MOREC_CHECK_ARG((epilogue != MOREC_EPI_MUL_GELU_GRAD && epilogue != MOREC_EPI_MUL_RELU_GRAD) || aux,
                    "morec_gemm: activation-gradient epilogue needs aux");
    MOREC_CHECK_ARG(!(accumulate && epilogue != MOREC_EPI_LINEAR), "morec_gemm: accumulate only with EPI_LINEAR");
    GemmArgs g;
    g.A = A; g.B = B; g.C = C; g.C2 = C2;
    g.M = M; g.N = N; g.K = K;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.a_mn = a_mn_major; g.b_mn = b_mn_major;
    g.dtype = dtype; g.out_bf16 = out_bf16;
    g.accumulate = accumulate; g.allow_split_k = accumulate;
    StdEpi::Params ep;
    ep.mode = epilogue; ep.alpha = alpha; ep.bias = bias; ep.aux = aux; ep.ldaux = ldaux;
    ep.aux_bf16 = (dtype == 1);
    ep.fast = (dtype != 2);
    return gemm_dispatch_auto<StdEpi>(g, ep, (cudaStream_t)stream);
}
