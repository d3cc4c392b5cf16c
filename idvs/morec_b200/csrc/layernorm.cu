// Fused (dropout ->) residual/position add -> LayerNorm (-> dropout) forward and backward, one warp per row.
//
// Replaces the eager sequences   LayerNorm(residual + dropout(x))            modules.py:16-17, 62-63; HF BertSelfOutput
//                                dropout(LayerNorm(x + position_embedding))  modules.py:89-93; HF BertEmbeddings
// HBM-bound: forward reads x (+residual), writes y; backward reads dy (+dy2), y, writes dz (+dx_branch).
// The backward reconstructs xhat from the saved output: xhat = (y - beta) / gamma  (gamma must be non-zero),
// so neither the LN input nor the mean is kept.  Column reductions (dgamma, dbeta, dbias, dpos) are accumulated in
// registers per CTA and flushed with one atomicAdd per column per CTA: the target buffers must be zero-initialised
// (or hold a gradient to accumulate into).
#include "../../../include/morec_b200.h"
#include "common.cuh"

namespace morec {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 16;   // float4 per lane -> H <= 2048 (kernels are templated on the per-lane vector count NV)

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float4 drop4(float4 v, uint64_t seed, uint64_t idx4, uint32_t thresh, float scale) {
    // idx4 = element index / 4 : one Philox block covers the 4 lanes of a float4
    const uint4 r = Philox::gen(seed, idx4);
    v.x = r.x >= thresh ? v.x * scale : 0.f;
    v.y = r.y >= thresh ? v.y * scale : 0.f;
    v.z = r.z >= thresh ? v.z * scale : 0.f;
    v.w = r.w >= thresh ? v.w * scale : 0.f;
    return v;
}

struct LnFwdParams {
    const void* x; const void* residual; const float* pos; int pos_period;
    const float* gamma; const float* beta;
    void* y; void* y_pre; float* rstd;
    int M, H; float eps;
    float p_pre, p_post; uint64_t seed, off_pre, off_post;
};

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const LnFwdParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = p.H >> 2;   // float4 per row
    const uint32_t th_pre = (uint32_t)fminf(p.p_pre * 4294967296.f, 4294967295.f);
    const uint32_t th_post = (uint32_t)fminf(p.p_post * 4294967296.f, 4294967295.f);
    const float sc_pre = p.p_pre > 0.f ? 1.f / (1.f - p.p_pre) : 1.f;
    const float sc_post = p.p_post > 0.f ? 1.f / (1.f - p.p_post) : 1.f;
    for (int row = blockIdx.x * LN_WARPS + warp; row < p.M; row += gridDim.x * LN_WARPS) {
        const T* xr = reinterpret_cast<const T*>(p.x) + (size_t)row * p.H;
        const T* rr = p.residual ? reinterpret_cast<const T*>(p.residual) + (size_t)row * p.H : nullptr;
        const float* pr = p.pos ? p.pos + (size_t)(row % p.pos_period) * p.H : nullptr;
        float4 z[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                float4 v = load4<T>(xr + 4 * i);
                if (p.p_pre > 0.f) v = drop4(v, p.seed, p.off_pre + (uint64_t)row * nv + i, th_pre, sc_pre);
                if (rr) { const float4 r = load4<T>(rr + 4 * i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
                if (pr) { const float4 r = *reinterpret_cast<const float4*>(pr + 4 * i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
                z[j] = v;
                s += v.x + v.y + v.z + v.w;
            }
        }
        const float mean = warp_sum(s) / p.H;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                const float a = z[j].x - mean, b = z[j].y - mean, c = z[j].z - mean, d = z[j].w - mean;
                q += a * a + b * b + c * c + d * d;
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / p.H + p.eps);
        if (lane == 0 && p.rstd) p.rstd[row] = rstd;
        T* yr = reinterpret_cast<T*>(p.y) + (size_t)row * p.H;
        T* ypr = p.y_pre ? reinterpret_cast<T*>(p.y_pre) + (size_t)row * p.H : nullptr;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                const float4 g = *reinterpret_cast<const float4*>(p.gamma + 4 * i);
                const float4 b = *reinterpret_cast<const float4*>(p.beta + 4 * i);
                float4 o;
                o.x = (z[j].x - mean) * rstd * g.x + b.x;
                o.y = (z[j].y - mean) * rstd * g.y + b.y;
                o.z = (z[j].z - mean) * rstd * g.z + b.z;
                o.w = (z[j].w - mean) * rstd * g.w + b.w;
                if (ypr) store4<T>(ypr + 4 * i, o);
                if (p.p_post > 0.f) o = drop4(o, p.seed, p.off_post + (uint64_t)row * nv + i, th_post, sc_post);
                store4<T>(yr + 4 * i, o);
            }
        }
    }
}

struct LnBwdParams {
    const void* dy; const void* dy2; const void* y; const float* gamma; const float* beta; const float* rstd;
    void* dz; void* dx_branch;
    float* dgamma; float* dbeta; float* dbias; float* dpos; int pos_period;
    int M, H;
    float p_pre, p_post; uint64_t seed, off_pre, off_post;
};

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const LnBwdParams p) {
    extern __shared__ float red[];   // [LN_WARPS][3][H] for the column reductions
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = p.H >> 2;
    const uint32_t th_pre = (uint32_t)fminf(p.p_pre * 4294967296.f, 4294967295.f);
    const uint32_t th_post = (uint32_t)fminf(p.p_post * 4294967296.f, 4294967295.f);
    const float sc_pre = p.p_pre > 0.f ? 1.f / (1.f - p.p_pre) : 1.f;
    const float sc_post = p.p_post > 0.f ? 1.f / (1.f - p.p_post) : 1.f;
    float4 ag[NV], ab[NV], ax[NV];   // dgamma, dbeta, dbias partials
    float4 gam[NV], bet[NV], igam[NV];   // this lane's columns of gamma, beta, 1/gamma (loop invariant)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        ag[j] = ab[j] = ax[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i = lane + 32 * j;
        if (i < nv) {
            gam[j] = *reinterpret_cast<const float4*>(p.gamma + 4 * i);
            bet[j] = *reinterpret_cast<const float4*>(p.beta + 4 * i);
            igam[j] = make_float4(1.f / gam[j].x, 1.f / gam[j].y, 1.f / gam[j].z, 1.f / gam[j].w);
        } else {
            gam[j] = bet[j] = igam[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    for (int row = blockIdx.x * LN_WARPS + warp; row < p.M; row += gridDim.x * LN_WARPS) {
        const T* dyr = reinterpret_cast<const T*>(p.dy) + (size_t)row * p.H;
        const T* dy2r = p.dy2 ? reinterpret_cast<const T*>(p.dy2) + (size_t)row * p.H : nullptr;
        const T* yr = reinterpret_cast<const T*>(p.y) + (size_t)row * p.H;
        const float rstd = p.rstd[row];
        float4 g[NV], xh[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                float4 d = load4<T>(dyr + 4 * i);
                if (dy2r) { const float4 e = load4<T>(dy2r + 4 * i); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
                if (p.p_post > 0.f) d = drop4(d, p.seed, p.off_post + (uint64_t)row * nv + i, th_post, sc_post);
                const float4 yv = load4<T>(yr + 4 * i);
                const float4 ga = gam[j], be = bet[j], ig = igam[j];
                float4 h;
                h.x = (yv.x - be.x) * ig.x; h.y = (yv.y - be.y) * ig.y; h.z = (yv.z - be.z) * ig.z; h.w = (yv.w - be.w) * ig.w;
                ag[j].x += d.x * h.x; ag[j].y += d.y * h.y; ag[j].z += d.z * h.z; ag[j].w += d.w * h.w;
                ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
                d.x *= ga.x; d.y *= ga.y; d.z *= ga.z; d.w *= ga.w;
                g[j] = d; xh[j] = h;
                s1 += d.x + d.y + d.z + d.w;
                s2 += d.x * h.x + d.y * h.y + d.z * h.z + d.w * h.w;
            }
        }
        const float m1 = warp_sum(s1) / p.H, m2 = warp_sum(s2) / p.H;
        T* dzr = reinterpret_cast<T*>(p.dz) + (size_t)row * p.H;
        T* dxr = p.dx_branch ? reinterpret_cast<T*>(p.dx_branch) + (size_t)row * p.H : nullptr;
        float* dpr = p.dpos ? p.dpos + (size_t)(row % p.pos_period) * p.H : nullptr;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < nv) {
                float4 o;
                o.x = rstd * (g[j].x - m1 - xh[j].x * m2);
                o.y = rstd * (g[j].y - m1 - xh[j].y * m2);
                o.z = rstd * (g[j].z - m1 - xh[j].z * m2);
                o.w = rstd * (g[j].w - m1 - xh[j].w * m2);
                store4<T>(dzr + 4 * i, o);
                if (dpr) {
                    atomicAdd(dpr + 4 * i, o.x); atomicAdd(dpr + 4 * i + 1, o.y);
                    atomicAdd(dpr + 4 * i + 2, o.z); atomicAdd(dpr + 4 * i + 3, o.w);
                }
                float4 b = o;
                if (p.p_pre > 0.f) b = drop4(o, p.seed, p.off_pre + (uint64_t)row * nv + i, th_pre, sc_pre);
                if (dxr) store4<T>(dxr + 4 * i, b);
                ax[j].x += b.x; ax[j].y += b.y; ax[j].z += b.z; ax[j].w += b.w;
            }
        }
    }
    // ---- column reductions: warps -> smem -> one atomicAdd per column per CTA
    float* rg = red + (size_t)warp * 3 * p.H;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = lane + 32 * j;
        if (i < nv) {
            *reinterpret_cast<float4*>(rg + 4 * i) = ag[j];
            *reinterpret_cast<float4*>(rg + p.H + 4 * i) = ab[j];
            *reinterpret_cast<float4*>(rg + 2 * p.H + 4 * i) = ax[j];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * p.H; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[(size_t)w * 3 * p.H + c];
        const int which = c / p.H, col = c % p.H;
        float* dst = which == 0 ? p.dgamma : which == 1 ? p.dbeta : p.dbias;
        if (dst) atomicAdd(dst + col, s);
    }
}

}  // namespace morec

using namespace morec;

extern "C" int morec_layernorm_fwd(const void* x, const void* residual, const float* pos, int pos_period,
                                   const float* gamma, const float* beta, void* y, void* y_pre, float* rstd, int M,
                                   int H, float eps, int dtype, float p_pre, float p_post, uint64_t seed,
                                   uint64_t off_pre, uint64_t off_post, void* stream) {
    MOREC_CHECK_ARG(x && gamma && beta && y, "layernorm_fwd: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && H <= LN_MAXV * 128 && H > 0, "layernorm_fwd: H=%d must be a multiple of 4, <= %d", H,
                    LN_MAXV * 128);
    MOREC_CHECK_ARG(!(p_post > 0.f) || y_pre, "layernorm_fwd: post-dropout needs y_pre for the backward");
    if (M <= 0) return MOREC_OK;
    LnFwdParams p{x, residual, pos, pos_period > 0 ? pos_period : 1, gamma, beta, y, y_pre, rstd, M, H, eps,
                  p_pre, p_post, seed, off_pre, off_post};
    int blocks = (M + LN_WARPS - 1) / LN_WARPS;
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const int nvl = (H / 4 + 31) / 32;
#define LN_FWD(NVV)                                                                                         \
    do {                                                                                                    \
        if (dtype == 0) ln_fwd_kernel<float, NVV><<<blocks, LN_WARPS * 32, 0, (cudaStream_t)stream>>>(p);    \
        else ln_fwd_kernel<__nv_bfloat16, NVV><<<blocks, LN_WARPS * 32, 0, (cudaStream_t)stream>>>(p);       \
    } while (0)
    if (nvl <= 2) LN_FWD(2); else if (nvl <= 4) LN_FWD(4); else if (nvl <= 6) LN_FWD(6); else if (nvl <= 8) LN_FWD(8); else LN_FWD(16);
#undef LN_FWD
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_layernorm_bwd(const void* dy, const void* dy2, const void* y, const float* gamma,
                                   const float* beta, const float* rstd, void* dz, void* dx_branch, float* dgamma,
                                   float* dbeta, float* dbias, float* dpos, int pos_period, int M, int H, int dtype,
                                   float p_pre, float p_post, uint64_t seed, uint64_t off_pre, uint64_t off_post,
                                   void* stream) {
    MOREC_CHECK_ARG(dy && y && gamma && beta && rstd && dz, "layernorm_bwd: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && H <= LN_MAXV * 128 && H > 0, "layernorm_bwd: H=%d unsupported", H);
    MOREC_CHECK_ARG(!(p_pre > 0.f) || dx_branch, "layernorm_bwd: pre-dropout needs dx_branch");
    if (M <= 0) return MOREC_OK;
    LnBwdParams p{dy, dy2, y, gamma, beta, rstd, dz, dx_branch, dgamma, dbeta, dbias, dpos,
                  pos_period > 0 ? pos_period : 1, M, H, p_pre, p_post, seed, off_pre, off_post};
    int blocks = (M + LN_WARPS - 1) / LN_WARPS;
    const int cap = num_sms() * 2;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)LN_WARPS * 3 * H * sizeof(float);
    const int nvl = (H / 4 + 31) / 32;
#define LN_BWD(TT, NVV)                                                                                              \
    do {                                                                                                             \
        static bool set_ = false;                                                                                    \
        if (!set_) {                                                                                                 \
            MOREC_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<TT, NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                            200 * 1024));                                                            \
            set_ = true;                                                                                             \
        }                                                                                                            \
        ln_bwd_kernel<TT, NVV><<<blocks, LN_WARPS * 32, smem, (cudaStream_t)stream>>>(p);                            \
    } while (0)
#define LN_BWD_T(NVV)                                         \
    do {                                                      \
        if (dtype == 0) LN_BWD(float, NVV);                   \
        else LN_BWD(__nv_bfloat16, NVV);                      \
    } while (0)
    if (nvl <= 2) LN_BWD_T(2); else if (nvl <= 4) LN_BWD_T(4); else if (nvl <= 6) LN_BWD_T(6); else if (nvl <= 8) LN_BWD_T(8); else LN_BWD_T(16);
#undef LN_BWD_T
#undef LN_BWD
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}
