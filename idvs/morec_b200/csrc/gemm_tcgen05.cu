// TMA descriptor construction + the C-ABI entry point morec_gemm (see include/morec_b200.h); the kernels live in
// gemm_std_m<MODE>.cu, one translation unit per epilogue mode.
#include "gemm_std_epi.cuh"

#include <stdlib.h>

#include <atomic>
#include <mutex>

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
// dynamic tile scheduling of the CTA-pair kernel: counter pool
// ------------------------------------------------------------------------------------------------
constexpr int kSchedSlots = 512;
__device__ int g_sched_ctr[kSchedSlots * 2];          // zero-initialised; every kernel leaves its slot zeroed again

int* gemm_sched_slot() {
    // OPT-IN (MOREC_GEMM_DYN=1).  Measured on B200: free when a GEMM runs alone (12037x3072x768: 51.4 us either way),
    // but no gain where it was expected to help -- with the 2-GPU gradient all-reduce alongside, GEMM time per step
    // stayed at 7.76 ms (static: 7.75; 7.1-7.2 on one GPU), and the single-GPU step went 10.74 -> 10.86 ms.  The
    // slowdown under NCCL is therefore not late-starting CTA pairs.
    static const bool enabled = []() { const char* e = getenv("MOREC_GEMM_DYN"); return e && e[0] == '1'; }();
    if (!enabled) return nullptr;
    static int* base[16] = {nullptr};
    static std::atomic<unsigned> seq{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!base[dev]) {
        void* ptr = nullptr;
        if (cudaGetSymbolAddress(&ptr, g_sched_ctr) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        base[dev] = static_cast<int*>(ptr);
    }
    return base[dev] + 2 * (seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
}

}  // namespace morec

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int morec_gemm(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux,
                          int M, int N, int K, int lda, int ldb, int ldc, int ldaux, int a_mn_major, int b_mn_major,
                          int dtype, int out_dtype, int epilogue, float alpha, int accumulate, void* stream) {
    using namespace morec;
    MOREC_CHECK_ARG(A && B && C, "morec_gemm: null operand");
    MOREC_CHECK_ARG(epilogue >= MOREC_EPI_LINEAR && epilogue <= MOREC_EPI_MUL_AUX, "morec_gemm: bad epilogue %d",
                    epilogue);
    MOREC_CHECK_ARG((epilogue != MOREC_EPI_GELU && epilogue != MOREC_EPI_GELU_DGELU) || C2,
                    "morec_gemm: EPI_GELU / EPI_GELU_DGELU need C2 (second output)");
    MOREC_CHECK_ARG((epilogue != MOREC_EPI_MUL_GELU_GRAD && epilogue != MOREC_EPI_MUL_RELU_GRAD && epilogue != MOREC_EPI_MUL_AUX) || aux,
                    "morec_gemm: activation-gradient epilogue needs aux");
    MOREC_CHECK_ARG(!(accumulate && epilogue != MOREC_EPI_LINEAR), "morec_gemm: accumulate only with EPI_LINEAR");
    GemmArgs g;
    g.A = A; g.B = B; g.C = C; g.C2 = C2;
    g.M = M; g.N = N; g.K = K;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.a_mn = a_mn_major; g.b_mn = b_mn_major;
    MOREC_CHECK_ARG(out_dtype == 0 || out_dtype == 1 || out_dtype == 3, "morec_gemm: out_dtype must be 0 (fp32), 1 (bf16) or 3 (fp16)");
    g.dtype = dtype; g.out_bf16 = out_dtype != 0; g.out_f16 = out_dtype == 3;
    g.accumulate = accumulate; g.allow_split_k = accumulate;
    g.aux = aux; g.ldaux = ldaux;
    StdEpiParams ep;
    ep.mode = epilogue; ep.alpha = alpha; ep.bias = bias; ep.aux = aux; ep.ldaux = ldaux;
    ep.aux_bf16 = MOREC_DT_IS16(dtype);
    ep.aux_f16 = dtype == 3;
    cudaStream_t st = (cudaStream_t)stream;
    switch (epilogue) {
        case MOREC_EPI_LINEAR: return gemm_std_run_0(g, ep, st);
        case MOREC_EPI_GELU: return gemm_std_run_1(g, ep, st);
        case MOREC_EPI_GELU_NOSAVE: return gemm_std_run_2(g, ep, st);
        case MOREC_EPI_RELU: return gemm_std_run_3(g, ep, st);
        case MOREC_EPI_MUL_GELU_GRAD: return gemm_std_run_4(g, ep, st);
        case MOREC_EPI_MUL_RELU_GRAD: return gemm_std_run_5(g, ep, st);
        case MOREC_EPI_GELU_DGELU: return gemm_std_run_6(g, ep, st);
        default: return gemm_std_run_7(g, ep, st);
    }
}
