// Fused (dropout ->) residual/position add -> LayerNorm (-> dropout) forward (one warp per row) and backward (one
// thread per float4 column, LN_R rows per block step).
//
// Replaces the eager sequences   LayerNorm(residual + dropout(x))            modules.py:16-17, 62-63; HF BertSelfOutput
//                                dropout(LayerNorm(x + position_embedding))  modules.py:89-93; HF BertEmbeddings
// HBM-bound: forward reads x (+residual), writes y; backward reads dy (+dy2), y, writes dz (+dx_branch).
// The backward reconstructs xhat from the saved output: xhat = (y - beta) / gamma  (gamma must be non-zero),
// so neither the LN input nor the mean is kept.  Column reductions (dgamma, dbeta, dbias, dpos) are accumulated in
// registers per thread and flushed with one (vector) atomicAdd per column group per CTA: the target buffers must be zero-initialised
// (or hold a gradient to accumulate into).
#include "../../../include/morec_b200.h"
#include "common.cuh"

namespace morec {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 16;   // float4 per lane -> H <= 2048 (kernels are templated on the per-lane vector count NV)

// Dropout of one float4 (4 consecutive columns) of row `row`: rows are paired, the Philox block of (row >> 1, float4
// column i) carries the decisions of both rows -- row & 1 selects the 16-bit half of each word.
__device__ __forceinline__ float4 drop4_bits(float4 v, uint4 r, int half, uint32_t th16, float scale) {
    v.x = rnd16(r.x, half) >= th16 ? v.x * scale : 0.f;
    v.y = rnd16(r.y, half) >= th16 ? v.y * scale : 0.f;
    v.z = rnd16(r.z, half) >= th16 ? v.z * scale : 0.f;
    v.w = rnd16(r.w, half) >= th16 ? v.w * scale : 0.f;
    return v;
}
__device__ __forceinline__ uint4 drop_block(uint64_t seed, uint64_t off, int row, int nv, int i) {
    return Philox7::gen(seed, off + (uint64_t)(row >> 1) * nv + i);
}

struct LnFwdParams {
    const void* x; const void* residual; const float* pos; int pos_period;
    const float* gamma; const float* beta;
    void* y; void* y_pre; float* rstd;
    int M, H; float eps;
    float p_pre, p_post; uint64_t seed, off_pre, off_post;
};

// One warp per row.  (Letting a warp normalise both rows of a dropout pair halves the Philox work but doubles the live
// registers: measured 23.6 us against 21.3 us for this layout at 12,037 x 768 bf16.)
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const LnFwdParams p) {
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = p.H >> 2;   // float4 per row
    const uint32_t th_pre = drop_thresh16(p.p_pre), th_post = drop_thresh16(p.p_post);
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
                float4 v = ld4<T>(xr + 4 * i);
                if (p.p_pre > 0.f) v = drop4_bits(v, drop_block(p.seed, p.off_pre, row, nv, i), row & 1, th_pre, sc_pre);
                if (rr) { const float4 r = ld4<T>(rr + 4 * i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
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
                if (ypr) st4<T>(ypr + 4 * i, o);
                if (p.p_post > 0.f) o = drop4_bits(o, drop_block(p.seed, p.off_post, row, nv, i), row & 1, th_post, sc_post);
                st4<T>(yr + 4 * i, o);
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

// Backward: thread c owns float4 column c of EVERY row its CTA visits (blockDim = H/4 rounded up to a warp), so the
// three column reductions cost 12 registers per thread instead of 3*H/32 per lane (the former warp-per-row layout
// needed 230 registers at H = 768 and ran 8 warps per SM: 72 us against a 24 us HBM floor).  Rows are processed LN_R at
// a time; their 2*LN_R row statistics are reduced with one warp shuffle tree each plus ONE __syncthreads per group
// (partials double-buffered in shared memory).
constexpr int LN_R = 4;

template <typename T, bool BIG>     // BIG: more than 256 threads (H > 1024); otherwise registers are capped for 3 CTAs of 256
__global__ void __launch_bounds__(BIG ? 512 : 256, BIG ? 1 : 3) ln_bwd_kernel(const LnBwdParams p) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float red[2][16][2 * LN_R];
    const int c = threadIdx.x, warp = c >> 5, lane = c & 31;
    const int nwarps = blockDim.x >> 5;
    const int nv = p.H >> 2;
    const bool act = c < nv;
    const float invH = 1.f / (float)p.H;
    const uint32_t th_pre = drop_thresh16(p.p_pre), th_post = drop_thresh16(p.p_post);
    const float sc_pre = p.p_pre > 0.f ? 1.f / (1.f - p.p_pre) : 1.f;
    const float sc_post = p.p_post > 0.f ? 1.f / (1.f - p.p_post) : 1.f;
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, ax = ag;      // dgamma, dbeta, dbias partials of column c
    float4 gam = ag, bet = ag, ig = ag;
    if (act) {
        gam = *reinterpret_cast<const float4*>(p.gamma + 4 * c);
        bet = *reinterpret_cast<const float4*>(p.beta + 4 * c);
        // xhat is rebuilt as (y - beta) / gamma; a gamma of exactly 0 carries no xhat information (its column then
        // contributes nothing to dx either, dy * gamma = 0): treat xhat as 0 there instead of producing inf / NaN
        ig = make_float4(gam.x != 0.f ? 1.f / gam.x : 0.f, gam.y != 0.f ? 1.f / gam.y : 0.f,
                         gam.z != 0.f ? 1.f / gam.z : 0.f, gam.w != 0.f ? 1.f / gam.w : 0.f);
    }
    const T* dyb = reinterpret_cast<const T*>(p.dy);
    const T* dy2b = reinterpret_cast<const T*>(p.dy2);
    const T* yb = reinterpret_cast<const T*>(p.y);
    T* dzb = reinterpret_cast<T*>(p.dz);
    T* dxb = reinterpret_cast<T*>(p.dx_branch);
    int buf = 0;
    for (int row0 = blockIdx.x * LN_R; row0 < p.M; row0 += gridDim.x * LN_R, buf ^= 1) {
        float4 d[LN_R], yv[LN_R];
        float rs[LN_R];
        // ---- all global loads of the row group first
#pragma unroll
        for (int r = 0; r < LN_R; ++r) {
            const int row = row0 + r;
            d[r] = yv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            rs[r] = 0.f;
            if (act && row < p.M) {
                const size_t o = (size_t)row * p.H + 4 * c;
                d[r] = ld4<T>(dyb + o);
                if (dy2b) { const float4 e = ld4<T>(dy2b + o); d[r].x += e.x; d[r].y += e.y; d[r].z += e.z; d[r].w += e.w; }
                yv[r] = ld4<T>(yb + o);
                rs[r] = p.rstd[row];
            }
        }
        float st[2 * LN_R];
        uint4 rb[LN_R / 2];                               // one Philox block per row PAIR (row0 is a multiple of LN_R)
        if (p.p_post > 0.f) {
#pragma unroll
            for (int r = 0; r < LN_R / 2; ++r) rb[r] = drop_block(p.seed, p.off_post, row0 + 2 * r, nv, c);
        }
#pragma unroll
        for (int r = 0; r < LN_R; ++r) {
            const int row = row0 + r;
            float4 dd = d[r], h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act && row < p.M) {
                if (p.p_post > 0.f) dd = drop4_bits(dd, rb[r >> 1], r & 1, th_post, sc_post);
                h.x = (yv[r].x - bet.x) * ig.x; h.y = (yv[r].y - bet.y) * ig.y;
                h.z = (yv[r].z - bet.z) * ig.z; h.w = (yv[r].w - bet.w) * ig.w;
                ag.x += dd.x * h.x; ag.y += dd.y * h.y; ag.z += dd.z * h.z; ag.w += dd.w * h.w;
                ab.x += dd.x; ab.y += dd.y; ab.z += dd.z; ab.w += dd.w;
                dd.x *= gam.x; dd.y *= gam.y; dd.z *= gam.z; dd.w *= gam.w;
            }
            d[r] = dd; yv[r] = h;                       // d <- dy*gamma, yv <- xhat
            st[2 * r] = warp_sum(dd.x + dd.y + dd.z + dd.w);
            st[2 * r + 1] = warp_sum(dd.x * h.x + dd.y * h.y + dd.z * h.z + dd.w * h.w);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 2 * LN_R; i += 4)
                *reinterpret_cast<float4*>(&red[buf][warp][i]) = make_float4(st[i], st[i + 1], st[i + 2], st[i + 3]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2 * LN_R; ++i) st[i] = 0.f;
        for (int w = 0; w < nwarps; ++w) {
#pragma unroll
            for (int i = 0; i < 2 * LN_R; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&red[buf][w][i]);
                st[i] += v.x; st[i + 1] += v.y; st[i + 2] += v.z; st[i + 3] += v.w;
            }
        }
        if (p.p_pre > 0.f) {
#pragma unroll
            for (int r = 0; r < LN_R / 2; ++r) rb[r] = drop_block(p.seed, p.off_pre, row0 + 2 * r, nv, c);
        }
#pragma unroll
        for (int r = 0; r < LN_R; ++r) {
            const int row = row0 + r;
            if (act && row < p.M) {
                const float m1 = st[2 * r] * invH, m2 = st[2 * r + 1] * invH, rstd = rs[r];
                const size_t o = (size_t)row * p.H + 4 * c;
                float4 z;
                z.x = rstd * (d[r].x - m1 - yv[r].x * m2);
                z.y = rstd * (d[r].y - m1 - yv[r].y * m2);
                z.z = rstd * (d[r].z - m1 - yv[r].z * m2);
                z.w = rstd * (d[r].w - m1 - yv[r].w * m2);
                st4<T>(dzb + o, z);
                if (p.dpos) atomicAdd(reinterpret_cast<float4*>(p.dpos + (size_t)(row % p.pos_period) * p.H + 4 * c), z);
                float4 b = z;
                if (p.p_pre > 0.f) b = drop4_bits(z, rb[r >> 1], r & 1, th_pre, sc_pre);
                if (dxb) st4<T>(dxb + o, b);
                ax.x += b.x; ax.y += b.y; ax.z += b.z; ax.w += b.w;
            }
        }
    }
    if (act) {          // one vector atomic per column group per CTA
        if (p.dgamma) atomicAdd(reinterpret_cast<float4*>(p.dgamma + 4 * c), ag);
        if (p.dbeta) atomicAdd(reinterpret_cast<float4*>(p.dbeta + 4 * c), ab);
        if (p.dbias) atomicAdd(reinterpret_cast<float4*>(p.dbias + 4 * c), ax);
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
    // Grid: up to 8 CTAs per SM although only 4 are resident (62 registers): measured on B200 at 12,037 x 768, the
    // resulting two waves of short-lived CTAs (18.9 us) beat an occupancy-sized persistent grid (21.4 us) and register
    // caps for 5 / 6 resident CTAs (22.4 / 24.9 us, spills) -- the block scheduler's refill balances better than a
    // fixed row stride.
    int blocks = (M + LN_WARPS - 1) / LN_WARPS;
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const int nvl = (H / 4 + 31) / 32;
#define LN_FWD(NVV)                                                                                         \
    do {                                                                                                    \
        MOREC_DISPATCH_T(dtype, MOREC_CUDA(launch_pdl(ln_fwd_kernel<T, NVV>, dim3(blocks), dim3(LN_WARPS * 32), 0, (cudaStream_t)stream, p))); \
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
    const int threads = ((H / 4 + 31) / 32) * 32;                 // <= 512
    // persistent grid = resident CTAs (occupancy query): 4 CTAs/SM were launched where 3 fit, and the second partial
    // wave cost a third of the kernel (ncu: 1.33 waves)
    static int occ_f32[17] = {0}, occ_16[17] = {0};          // (bf16 and fp16 instantiations have the same footprint)
    int& occ = (MOREC_DT_IS16(dtype) ? occ_16 : occ_f32)[threads / 32];
    const bool big = threads > 256;
    if (occ == 0) {
        cudaError_t e;
        if (MOREC_DT_IS16(dtype)) e = big ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<__nv_bfloat16, true>, threads, 0)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<__nv_bfloat16, false>, threads, 0);
        else e = big ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<float, true>, threads, 0)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<float, false>, threads, 0);
        MOREC_CUDA(e);
        if (occ < 1) occ = 1;
    }
    int blocks = (M + LN_R - 1) / LN_R;
    const int cap = num_sms() * occ;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (big) MOREC_DISPATCH_T(dtype, MOREC_CUDA(launch_pdl(ln_bwd_kernel<T, true>, dim3(blocks), dim3(threads), 0, st, p)));
    else MOREC_DISPATCH_T(dtype, MOREC_CUDA(launch_pdl(ln_bwd_kernel<T, false>, dim3(blocks), dim3(threads), 0, st, p)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}
