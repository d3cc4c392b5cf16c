// Where does a CTA pair of the tcgen05 GEMM spend its cycles?  Compiles the CTA-pair kernel with MOREC_GEMM_TRACE and
// prints, for CTA 0, the clock64() stamps of the producer (slot free), the MMA issuer (operands landed / K-block
// issued / accumulator stage free) and epilogue warp 4 (accumulator ready / tile stored).  Diagnostic only.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMOREC_GEMM_TRACE \
//        tools/gemm_trace.cu -o gpurun_out/gemm_trace -Lidvs/morec_b200 -lmorec_b200 -Xlinker -rpath=idvs/morec_b200
//   gpurun_out/gemm_trace M N K [epilogue: 0 linear | 6 gelu+dgelu | 7 mul_aux]
#include "../idvs/morec_b200/csrc/gemm_std_epi.cuh"

#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace morec;

template <int MODE>
static int run(const GemmArgs& g, const StdEpiParams& ep) { return gemm2_launch<1, StdEpi<MODE>>(g, ep, 0); }

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 12037, N = argc > 2 ? atoi(argv[2]) : 3072, K = argc > 3 ? atoi(argv[3]) : 768;
    const int mode = argc > 4 ? atoi(argv[4]) : 0;
    __half *A, *B, *C, *C2, *aux;
    float* bias;
    cudaMalloc(&A, (size_t)M * K * 2); cudaMalloc(&B, (size_t)N * K * 2); cudaMalloc(&C, (size_t)M * N * 2);
    cudaMalloc(&C2, (size_t)M * N * 2); cudaMalloc(&aux, (size_t)M * N * 2); cudaMalloc(&bias, (size_t)N * 4);
    cudaMemset(A, 0x11, (size_t)M * K * 2); cudaMemset(B, 0x11, (size_t)N * K * 2); cudaMemset(aux, 0x3c, (size_t)M * N * 2);
    cudaMemset(bias, 0, (size_t)N * 4);
    GemmArgs g{};
    g.A = A; g.B = B; g.C = C; g.C2 = (mode == 6) ? C2 : nullptr;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N;
    g.dtype = 3; g.out_bf16 = 1; g.out_f16 = 1;
    g.aux = mode == 7 ? aux : nullptr; g.ldaux = N;
    StdEpiParams ep{};
    ep.mode = mode; ep.alpha = 1.f; ep.bias = bias; ep.aux = g.aux; ep.ldaux = N; ep.aux_bf16 = 1; ep.aux_f16 = 1;
    auto launch = [&]() { return mode == 6 ? run<6>(g, ep) : mode == 7 ? run<7>(g, ep) : run<0>(g, ep); };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) if (launch()) { printf("launch failed: %s\n", morec_last_error()); return 1; }
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("M=%d N=%d K=%d mode=%d: %.2f us/launch  %.1f TFLOP/s  (%s)\n", M, N, K, mode, ms * 1e3 / reps,
           2.0 * M * N * K / (ms * 1e-3 / reps) / 1e12, cudaGetErrorString(cudaGetLastError()));
    static long long h[6][4096];
    cudaMemcpyFromSymbol(h, g_gemm_trace, sizeof(h));
    const int num_kb = (K + 63) / 64;
    const int tiles = ((M + 255) / 256) * ((N + 255) / 256);
    const int my_tiles = (tiles + 73) / 74;    // CTA pair 0 always has the maximum
    const long long t0 = h[0][0];
    printf("CTA0: %d tiles x %d K-blocks; cycles relative to the producer's first slot\n", my_tiles, num_kb);
    printf("%4s %9s %9s %9s %9s %9s | %8s %8s %8s\n", "tile", "acc_free", "kb0_ready", "last_issue", "epi_start", "epi_end",
           "mma_span", "epi_span", "kb_avg");
    for (int t = 0; t < my_tiles && t < 64; ++t) {
        const int k0 = t * num_kb, k1 = k0 + num_kb - 1;
        printf("%4d %9lld %9lld %9lld %9lld %9lld | %8lld %8lld %8.0f\n", t, h[3][t] - t0, h[1][k0] - t0, h[2][k1] - t0,
               h[4][t] - t0, h[5][t] - t0, h[2][k1] - h[1][k0], h[5][t] - h[4][t], (double)(h[2][k1] - h[1][k0]) / num_kb);
    }
    // steady-state K-block cadence of the middle tile: issue-to-issue and producer slot-free intervals
    const int tm = my_tiles / 2;
    printf("tile %d K-block stamps (ready, issued, producer slot-free), deltas to previous issue:\n", tm);
    for (int kb = 0; kb < num_kb && kb < 48; ++kb) {
        const int i = tm * num_kb + kb;
        printf("  kb %2d ready %9lld issued %9lld (+%5lld)  prod %9lld\n", kb, h[1][i] - t0, h[2][i] - t0,
               i > 0 ? h[2][i] - h[2][i - 1] : 0, h[0][i] - t0);
    }
    return 0;
}
