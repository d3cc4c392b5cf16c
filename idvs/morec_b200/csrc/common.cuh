// Shared device/host helpers for the morec_b200 C-ABI library (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace morec {

// ---------------------------------------------------------------- error convention (C-ABI: int rc)
void set_last_error(const char* fmt, ...);
#define MOREC_OK 0
#define MOREC_ERR_ARG (-1)
#define MOREC_ERR_CUDA (-2)
#define MOREC_ERR_UNSUPPORTED (-3)

#define MOREC_CHECK_ARG(cond, ...)                               \
    do {                                                         \
        if (!(cond)) {                                           \
            ::morec::set_last_error(__VA_ARGS__);                \
            return MOREC_ERR_ARG;                                \
        }                                                        \
    } while (0)

#define MOREC_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ::morec::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,           \
                                    cudaGetErrorString(e__));                              \
            return MOREC_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

#define MOREC_LAUNCH_CHECK() MOREC_CUDA(cudaGetLastError())

int num_sms();

// ---- programmatic dependent launch (on by default; MOREC_PDL=0 disables) ---------------------------------------
// The kernels of the BERT tower sequence (GEMMs, LayerNorm, 16-bit attention, column sums: ~290 of a step's ~320
// launches) begin with pdl_wait() -- griddepcontrol.wait: block until every prerequisite grid has completed and its
// memory is visible; a no-op for a normal launch -- followed by pdl_trigger().  Launched through launch_pdl() with the
// programmatic-stream-serialization attribute, the NEXT kernel of the stream is set up (and its CTAs made resident as
// SMs free up) while the current one still runs, so the launch latency between dependent kernels is hidden; all
// reads and writes of a kernel still happen after its predecessor has finished.  Measured: 10.61 -> 10.46 ms per
// BERT-base step (A/B twice on one box).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Storage dtype codes of the C ABI: 0 = fp32, 1 = bf16, 3 = fp16 (2 = fp32 storage with 3xTF32 GEMM math; it only
// differs from 0 inside morec_gemm).  MOREC_DISPATCH_T runs the statement with `T` bound to the element type.
#define MOREC_DT_IS16(dt) ((dt) == 1 || (dt) == 3)
#define MOREC_DISPATCH_T(dtype, ...)                                     \
    do {                                                                 \
        if ((dtype) == 1) { using T = __nv_bfloat16; __VA_ARGS__; }      \
        else if ((dtype) == 3) { using T = __half; __VA_ARGS__; }        \
        else { using T = float; __VA_ARGS__; }                           \
    } while (0)

// 2 x fp32 <-> one packed pair of 16-bit floats; f16 selects IEEE half instead of bfloat16 (warp-uniform runtime flag:
// the tcgen05 epilogues serve both 16-bit storage types with one instantiation)
__device__ __forceinline__ uint32_t pack2_16(float a, float b, bool f16) {
    if (f16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2_16(uint32_t u, bool f16) {
    if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ float unpack1_16(uint16_t u, bool f16) {
    if (f16) return __half2float(*reinterpret_cast<const __half*>(&u));
    return __uint_as_float((uint32_t)u << 16);
}

// 4 consecutive elements <-> float4 for every storage type
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <>
__device__ __forceinline__ float4 ld4<__half>(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void st4<__half>(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
template <>
__device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
        "elect.sync R|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int KIND>  // 0 = tf32, 1 = f16/bf16
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Fast erf-GELU for the fast precision modes: q(x) = 0.5 * erfc(|x|/sqrt2) from the Abramowitz-Stegun 7.1.26 rational
// form (|error| <= 1.5e-7 absolute; measured 3.3e-7 on gelu, 2.9e-7 on gelu' in fp32), 2 MUFU + 9 FMA-pipe instructions
// instead of erff's ~40: the tensor-core GEMM epilogues are instruction-issue bound on this math.  The exponential is
// taken in base 2 (log2 e folded into the argument scale), gelu(x) = max(x,0) - |x| q(x) needs no select, and
// phi(x) = exp(-x^2/2)/sqrt(2 pi) shares the exponential.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_fast_parts(float x, float& q, float& e) {
    const float u = fabsf(x) * 0.8493218f;                       // u^2 = (x^2/2) * log2(e)
    const float t = rcp_approx(fmaf(0.27273747f, u, 1.0f));      // 1 / (1 + 0.3275911 |x|/sqrt2)
    e = ex2_approx(-u * u);                                      // exp(-x^2/2)
    float poly = fmaf(t, 0.5307027145f, -0.7265760135f);         // 0.5 * (a5 t + a4)
    poly = fmaf(t, poly, 0.7107068705f);
    poly = fmaf(t, poly, -0.142248368f);
    poly = fmaf(t, poly, 0.127414796f);
    q = poly * t * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
    float q, e;
    gelu_fast_parts(x, q, e);
    return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
    float q, e;
    gelu_fast_parts(x, q, e);
    const float cdf = x < 0.f ? q : 1.0f - q;
    return fmaf(x, 0.39894228040143267794f * e, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Counter-based RNG for dropout (Philox4x32-10).  Same (seed, offset) in fwd and bwd regenerates the mask.
struct Philox {
    __device__ __forceinline__ static uint4 gen(uint64_t seed, uint64_t idx) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = 0x5bd1e995u, c3 = 0x9e3779b9u;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// Dropout masks of the hot kernels (LayerNorm, tensor-core attention): Philox4x32-7 (the shortest variant that passes
// BigCrush) and SIXTEEN bits per decision, so one block of 128 random bits serves 8 elements.  With 10 rounds and one
// block per 4 elements the generator was ~60 % of the instructions of the LayerNorm forward (ncu: 965 warp
// instructions per 768-wide row) and cost 6-17 us per launch.  keep <=> r16 >= round(p * 65536); P(keep) differs from
// 1 - p by < 8e-6.
struct Philox7 {
    __device__ __forceinline__ static uint4 gen(uint64_t seed, uint64_t idx) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = 0x5bd1e995u, c3 = 0x9e3779b9u;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ uint32_t drop_thresh16(float p) { return (uint32_t)fminf(p * 65536.f + 0.5f, 65535.f); }
// 16-bit lane `half` (0 = low, 1 = high) of word w
__device__ __forceinline__ uint32_t rnd16(uint32_t w, int half) { return half ? (w >> 16) : (w & 0xffffu); }

// keep-probability test for element `idx` (one 32-bit lane of the Philox block idx/4)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thresh /* p * 2^32 */) {
    const uint4 r = Philox::gen(seed, idx >> 2);
    const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
    return w >= thresh;
}

}  // namespace morec
