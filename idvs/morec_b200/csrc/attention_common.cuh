// Device helpers shared by the short-sequence (attention.cu) and general (attention_gen.cu) attention kernels.
#pragma once
#include "common.cuh"

namespace morec {

constexpr int AT_DCH = 64;      // head-dim chunk streamed through shared memory

// cooperative (one warp) load of rows [row0, row0+len) x cols [col0, col0+w) into tile[32][AT_DCH] (fp32), 4 elements
// per lane per iteration (w % 4 == 0, rows 16-byte aligned)
template <typename T>
__device__ __forceinline__ void load_tile(float* tile, const T* base, int ld, int row0, int len, int col0, int w,
                                          int lane) {
    const int nv = w >> 2;
    for (int idx = lane; idx < len * nv; idx += 32) {
        const int r = idx / nv, c = (idx - r * nv) << 2;
        *reinterpret_cast<float4*>(tile + r * AT_DCH + c) = ld4<T>(base + (size_t)(row0 + r) * ld + col0 + c);
    }
}

// per-lane store of a 64-wide (or narrower) register slice to a row
template <typename T>
__device__ __forceinline__ void store_row(T* row, const float (&v)[AT_DCH], int w) {
#pragma unroll
    for (int t = 0; t < AT_DCH / 4; ++t)
        if (4 * t < w) st4<T>(row + 4 * t, make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]));
}

// per-lane load of one 64-wide (or narrower) slice of a row into registers
template <typename T>
__device__ __forceinline__ void load_row(float (&v)[AT_DCH], const T* row, int w, bool active) {
#pragma unroll
    for (int t = 0; t < AT_DCH / 4; ++t) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && 4 * t < w) x = ld4<T>(row + 4 * t);
        v[4 * t] = x.x; v[4 * t + 1] = x.y; v[4 * t + 2] = x.z; v[4 * t + 3] = x.w;
    }
}

// acc[jj] += <v, tile[j0 + jj][0..w)> for 4 keys at once (4 independent FMA chains)
__device__ __forceinline__ void dot4keys(const float (&v)[AT_DCH], const float* tile, int j0, int w, float& a0, float& a1,
                                         float& a2, float& a3) {
    const float4* k0 = reinterpret_cast<const float4*>(tile + (j0 + 0) * AT_DCH);
    const float4* k1 = reinterpret_cast<const float4*>(tile + (j0 + 1) * AT_DCH);
    const float4* k2 = reinterpret_cast<const float4*>(tile + (j0 + 2) * AT_DCH);
    const float4* k3 = reinterpret_cast<const float4*>(tile + (j0 + 3) * AT_DCH);
    const int nt = w >> 2;
#pragma unroll
    for (int t = 0; t < AT_DCH / 4; ++t) {
        if (t < nt) {
            const float4 x0 = k0[t], x1 = k1[t], x2 = k2[t], x3 = k3[t];
            a0 += v[4 * t] * x0.x + v[4 * t + 1] * x0.y + v[4 * t + 2] * x0.z + v[4 * t + 3] * x0.w;
            a1 += v[4 * t] * x1.x + v[4 * t + 1] * x1.y + v[4 * t + 2] * x1.z + v[4 * t + 3] * x1.w;
            a2 += v[4 * t] * x2.x + v[4 * t + 1] * x2.y + v[4 * t + 2] * x2.z + v[4 * t + 3] * x2.w;
            a3 += v[4 * t] * x3.x + v[4 * t + 1] * x3.y + v[4 * t + 2] * x3.z + v[4 * t + 3] * x3.w;
        }
    }
}


}  // namespace morec
