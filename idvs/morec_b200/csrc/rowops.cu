// Row-wise gather / scatter / reduction kernels of the MoRec step (all HBM-bound, vectorised, grid-stride).
//
//   morec_bert_embed_fwd/bwd   HF BertEmbeddings gathers: word[id] + position[pos] + token_type[0]   (encoders.py:68)
//   morec_gather_rows          CLS pooling h[:,0] (encoders.py:69), unique-item -> slot expansion, nn.Embedding
//                              forward of the ID tower (model.py:37), input_embs[:, :-1] slicing (model.py:39-41)
//   morec_scatter_add_rows     the matching backward (embedding_dense_backward / index_select backward)
//   morec_colsum               bias gradients
//   morec_adamw_multi          multi-tensor AdamW with fused grad unscale + found-inf (run.py:159-162, 245-247)
#include "../../../include/morec_b200.h"
#include "common.cuh"

namespace morec {

// ---------------------------------------------------------------------------------------------- BERT embeddings
template <typename T>
__global__ void bert_embed_fwd_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ pos,
                                      const float* __restrict__ word, const float* __restrict__ posemb,
                                      const float* __restrict__ type0, T* __restrict__ out, int n_tok, int H) {
    const int nv = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_tok * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / nv), c = (int)(i - (size_t)t * nv) * 4;
        const float4 a = *reinterpret_cast<const float4*>(word + (size_t)ids[t] * H + c);
        const float4 b = *reinterpret_cast<const float4*>(posemb + (size_t)pos[t] * H + c);
        const float4 d = *reinterpret_cast<const float4*>(type0 + c);
        st4<T>(out + (size_t)t * H + c, make_float4(a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w));
    }
}

template <typename T>
__global__ void bert_embed_bwd_kernel(const T* __restrict__ dz, const int64_t* __restrict__ ids,
                                      const int32_t* __restrict__ pos, float* __restrict__ dword,
                                      float* __restrict__ dposemb, int n_tok, int H) {
    const int nv = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_tok * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / nv), c = (int)(i - (size_t)t * nv) * 4;
        const float4 g = ld4<T>(dz + (size_t)t * H + c);
        if (dword) {
            float* w = dword + (size_t)ids[t] * H + c;
            atomicAdd(w, g.x); atomicAdd(w + 1, g.y); atomicAdd(w + 2, g.z); atomicAdd(w + 3, g.w);
        }
        if (dposemb) {
            float* q = dposemb + (size_t)pos[t] * H + c;
            atomicAdd(q, g.x); atomicAdd(q + 1, g.y); atomicAdd(q + 2, g.z); atomicAdd(q + 3, g.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------- token packing plan
// The text tower runs on PACKED tokens (only real word pieces of non-pad items).  The host needs just one number per
// item to lay the batch out -- its count of real tokens (n x int32 instead of the n x T attention mask) -- and a
// prefix sum; which word piece goes where is resolved on the device.  Replaces the token-level index arithmetic the
// host used to do between the size-determining sync and the first layer (the GPU idles there).
__global__ void mask_row_lens_kernel(const int64_t* __restrict__ text, int ld, int T, int n, int32_t* __restrict__ lens) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int64_t* m = text + (size_t)warp * ld + T;          // attention-mask half of the row (run.py:93-98)
    int cnt = 0;
    for (int c = lane; c < T; c += 32) cnt += m[c] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) lens[warp] = cnt;
}

// one warp per encoded item s: its kept word pieces (attention mask != 0), in column order, go to rows
// [cu[s], cu[s+1]) of tok_ids / tok_pos (tok_pos = original column, so position embeddings equal HF's)
__global__ void pack_tokens_kernel(const int64_t* __restrict__ text, int ld, int T, const int32_t* __restrict__ enc_rows,
                                   const int32_t* __restrict__ cu, int n_enc, int64_t* __restrict__ tok_ids,
                                   int32_t* __restrict__ tok_pos) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_enc) return;
    const int64_t* row = text + (size_t)enc_rows[warp] * ld;
    int out = cu[warp];
    for (int c0 = 0; c0 < T; c0 += 32) {
        const int c = c0 + lane;
        const bool keep = c < T && row[T + c] != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int k = out + __popc(bal & ((1u << lane) - 1u));
            tok_ids[k] = row[c];
            tok_pos[k] = c;
        }
        out += __popc(bal);
    }
}

// ---------------------------------------------------------------------------------------------- gather / scatter
template <typename TS, typename TD>
__global__ void gather_rows_kernel(const TS* __restrict__ src, const int32_t* __restrict__ idx, TD* __restrict__ dst,
                                   int n, int H, int ld_src, int ld_dst) {
    const int nv = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / nv), c = (int)(i - (size_t)r * nv) * 4;
        const int s = idx[r];
        const float4 v = s >= 0 ? ld4<TS>(src + (size_t)s * ld_src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        st4<TD>(dst + (size_t)r * ld_dst + c, v);
    }
}

template <typename TS>
__global__ void scatter_add_rows_kernel(const TS* __restrict__ src, const int32_t* __restrict__ idx,
                                        float* __restrict__ dst, int n, int H, int ld_src, int ld_dst) {
    const int nv = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / nv), c = (int)(i - (size_t)r * nv) * 4;
        const int d = idx[r];
        if (d < 0) continue;
        const float4 v = ld4<TS>(src + (size_t)r * ld_src + c);
        float* o = dst + (size_t)d * ld_dst + c;
        atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w);
    }
}

// ---------------------------------------------------------------------------------------------- scaled row add / pooling
// out[r] = (x ? x[r] : 0) + alpha * (gscale ? gscale[r / rows_per_group] : 1) * y[idx ? idx[r] : r]
template <typename T>
__global__ void scale_add_rows_kernel(const T* __restrict__ x, const T* __restrict__ y, const int32_t* __restrict__ idx,
                                      const float* __restrict__ gscale, int rows_per_group, float alpha,
                                      T* __restrict__ out, int n, int H, int ld_y) {
    const int nv = H >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / nv), c = (int)(i - (size_t)r * nv) * 4;
        const int sr = idx ? idx[r] : r;
        const float sc = alpha * (gscale ? gscale[r / rows_per_group] : 1.f);
        float4 v = sr >= 0 ? ld4<T>(y + (size_t)sr * ld_y + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        if (x) { const float4 a = ld4<T>(x + (size_t)r * H + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        st4<T>(out + (size_t)r * H + c, v);
    }
}

// out[g] = mean over the rows_per_group consecutive rows of group g
template <typename T>
__global__ void mean_rows_kernel(const T* __restrict__ x, T* __restrict__ out, int n_groups, int rows_per_group, int H) {
    const int nv = H >> 2;
    const float inv = 1.f / rows_per_group;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_groups * nv;
         i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / nv), c = (int)(i - (size_t)g * nv) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < rows_per_group; ++r) {
            const float4 v = ld4<T>(x + ((size_t)g * rows_per_group + r) * H + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        st4<T>(out + (size_t)g * H + c, acc);
    }
}

// ---------------------------------------------------------------------------------------------- column sum
// out[n] += sum_m x[m, n]; CTA = 32 x 8 threads: each thread owns 4 columns, 8 row groups; grid-stride over rows.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N,
                                                     int ld, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 4;
    const int r0 = blockIdx.y * rows_per_cta;
    const int r1 = min(M, r0 + rows_per_cta);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < N) {
        for (int r = r0 + ty; r < r1; r += 8) {
            const float4 v = ld4<T>(x + (size_t)r * ld + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < N) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { acc.x += red[k][tx].x; acc.y += red[k][tx].y; acc.z += red[k][tx].z; acc.w += red[k][tx].w; }
        atomicAdd(out + c, acc.x); atomicAdd(out + c + 1, acc.y); atomicAdd(out + c + 2, acc.z); atomicAdd(out + c + 3, acc.w);
    }
}

// ---------------------------------------------------------------------------------------------- activation bwd
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ aux, T* __restrict__ out, size_t n4,
                               int mode) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 g = ld4<T>(dy + 4 * i), a = ld4<T>(aux + 4 * i);
        float4 o;
        if (mode == 0) {
            o.x = g.x * gelu_erf_grad(a.x); o.y = g.y * gelu_erf_grad(a.y);
            o.z = g.z * gelu_erf_grad(a.z); o.w = g.w * gelu_erf_grad(a.w);
        } else {
            o.x = a.x > 0.f ? g.x : 0.f; o.y = a.y > 0.f ? g.y : 0.f;
            o.z = a.z > 0.f ? g.z : 0.f; o.w = a.w > 0.f ? g.w : 0.f;
        }
        st4<T>(out + 4 * i, o);
    }
}

// ---------------------------------------------------------------------------------------------- casts
template <typename T>
__global__ void cast_f32_to_16_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        st4<T>(dst + 4 * i, *reinterpret_cast<const float4*>(src + 4 * i));
}

// many tensors in ONE launch (bf16 weight shadows of a whole tower: 49 casts for BERT-base): same chunk table as AdamW
struct CastTensor {
    const float* src; void* dst; long long n;
};
constexpr int CAST_CHUNK = 16384;
__device__ __forceinline__ int find_tensor(const int* __restrict__ chunk_start, int n_tensors, int chunk);
template <typename T>
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastTensor* __restrict__ tensors,
                                                         const int* __restrict__ chunk_start, int n_tensors, int n_chunks) {
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const int t = find_tensor(chunk_start, n_tensors, ci);
        const CastTensor ch = tensors[t];
        const long long e0 = (long long)(ci - chunk_start[t]) * CAST_CHUNK;
        const long long e1 = min(ch.n, e0 + CAST_CHUNK);
        for (long long i = e0 + threadIdx.x * 4; i < e1; i += blockDim.x * 4) {
            T* dst = reinterpret_cast<T*>(ch.dst);
            if (i + 4 <= e1) {
                st4<T>(dst + i, *reinterpret_cast<const float4*>(ch.src + i));
            } else {
                for (long long k = i; k < e1; ++k) stf<T>(dst + k, ch.src[k]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- AdamW (multi-tensor)
// One entry per parameter tensor; CTAs walk fixed-size chunks and find their tensor by binary search in the
// exclusive prefix sum of per-tensor chunk counts (chunk_start[n_tensors + 1]).
struct AdamTensor {
    float* p; const float* g; float* m; float* v; void* p16;   // p16: optional 16-bit copy of the updated parameter
    int n; float lr, wd, beta1, beta2, eps;
};
constexpr int ADAM_CHUNK = 16384;

__device__ __forceinline__ int find_tensor(const int* __restrict__ chunk_start, int n_tensors, int chunk) {
    int lo = 0, hi = n_tensors;   // chunk_start[lo] <= chunk < chunk_start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_start[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename T16>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamTensor* __restrict__ tensors,
                                                          const int* __restrict__ chunk_start, int n_tensors,
                                                          int n_chunks, const float* __restrict__ step_dev,
                                                          const float* __restrict__ grad_scale,
                                                          const float* __restrict__ found_inf) {
    // found_inf != 0 -> skip the whole step (GradScaler semantics)
    if (found_inf && *found_inf != 0.f) return;
    const float gs = grad_scale ? 1.f / *grad_scale : 1.f;
    const float step_no = *step_dev;             // already advanced by adam_step_kernel
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const int t = find_tensor(chunk_start, n_tensors, ci);
        const AdamTensor ch = tensors[t];
        const int e0 = (ci - chunk_start[t]) * ADAM_CHUNK;
        const int e1 = min(ch.n, e0 + ADAM_CHUNK);
        const float beta1 = ch.beta1, beta2 = ch.beta2, eps = ch.eps;
        const float bc1 = 1.f - powf(beta1, step_no), bc2 = 1.f - powf(beta2, step_no);
        const float rs2 = rsqrtf(bc2);
        const float decay = 1.f - ch.lr * ch.wd, step = ch.lr / bc1;
        for (int i = e0 + threadIdx.x * 4; i < e1; i += blockDim.x * 4) {
            if (i + 4 <= e1) {
                float4 p = *reinterpret_cast<float4*>(ch.p + i);
                float4 g = *reinterpret_cast<const float4*>(ch.g + i);
                float4 m = *reinterpret_cast<float4*>(ch.m + i);
                float4 v = *reinterpret_cast<float4*>(ch.v + i);
                float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float gg = gp[k] * gs;
                    pp[k] *= decay;
                    mp[k] = beta1 * mp[k] + (1.f - beta1) * gg;
                    vp[k] = beta2 * vp[k] + (1.f - beta2) * gg * gg;
                    pp[k] -= step * (mp[k] / (sqrtf(vp[k]) * rs2 + eps));
                }
                *reinterpret_cast<float4*>(ch.p + i) = p;
                *reinterpret_cast<float4*>(ch.m + i) = m;
                *reinterpret_cast<float4*>(ch.v + i) = v;
                if (ch.p16) st4<T16>(reinterpret_cast<T16*>(ch.p16) + i, p);
            } else {
                for (int k = i; k < e1; ++k) {
                    const float gg = ch.g[k] * gs;
                    float pk = ch.p[k] * decay;
                    const float mk = beta1 * ch.m[k] + (1.f - beta1) * gg;
                    const float vk = beta2 * ch.v[k] + (1.f - beta2) * gg * gg;
                    pk -= step * (mk / (sqrtf(vk) * rs2 + eps));
                    ch.p[k] = pk; ch.m[k] = mk; ch.v[k] = vk;
                    if (ch.p16) stf<T16>(reinterpret_cast<T16*>(ch.p16) + k, pk);
                }
            }
        }
    }
}

// step += 1 unless an overflow was found (runs after grad_check_kernel, before adamw_multi_kernel)
__global__ void adam_step_kernel(float* __restrict__ step, const float* __restrict__ found_inf) {
    if (!(found_inf && *found_inf != 0.f)) *step += 1.f;
}

// found_inf = 1 if any gradient is inf/nan
__global__ void __launch_bounds__(256) grad_check_kernel(const AdamTensor* __restrict__ tensors,
                                                         const int* __restrict__ chunk_start, int n_tensors,
                                                         int n_chunks, float* __restrict__ found_inf) {
    bool bad = false;
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const int t = find_tensor(chunk_start, n_tensors, ci);
        const AdamTensor ch = tensors[t];
        const int e0 = (ci - chunk_start[t]) * ADAM_CHUNK;
        const int e1 = min(ch.n, e0 + ADAM_CHUNK);
        // float4 loads, four independent ones in flight per thread (the scalar loop with its dependent `bad |=` chain
        // read at 2.1 TB/s: 211 us for the 440 MB gradient set); gradients are 16-byte aligned as in adamw_multi_kernel
        int i = e0 + threadIdx.x * 4;
        for (; i + 3 * 1024 + 4 <= e1; i += 4 * 1024) {
            const float4 a = *reinterpret_cast<const float4*>(ch.g + i);
            const float4 b = *reinterpret_cast<const float4*>(ch.g + i + 1024);
            const float4 c = *reinterpret_cast<const float4*>(ch.g + i + 2048);
            const float4 d = *reinterpret_cast<const float4*>(ch.g + i + 3072);
#define MOREC_BAD4(q) (!(fabsf(q.x) <= 3.0e38f) | !(fabsf(q.y) <= 3.0e38f) | !(fabsf(q.z) <= 3.0e38f) | !(fabsf(q.w) <= 3.0e38f))
            bad |= MOREC_BAD4(a) | MOREC_BAD4(b) | MOREC_BAD4(c) | MOREC_BAD4(d);
#undef MOREC_BAD4
        }
        for (; i < e1; i += 1024)
            for (int k = i; k < min(i + 4, e1); ++k) bad |= !(fabsf(ch.g[k]) <= 3.0e38f);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) *found_inf = 1.f;
}

// SM clock probe: spins ~20 us and reports (elapsed SM cycles, elapsed ns) so the host can derive the SM clock
// under load without touching NVML inside a timed region.
__global__ void clock_probe_kernel(unsigned long long* out) {
    unsigned long long t0, t1, c0, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 20000ull);
    c1 = clock64();
    out[0] = c1 - c0;
    out[1] = t1 - t0;
}

static int grid_for(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    const size_t cap = (size_t)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace morec

using namespace morec;

extern "C" int morec_bert_embed_fwd(const int64_t* ids, const int32_t* pos, const float* word, const float* posemb,
                                    const float* type0, void* out, int n_tok, int H, int dtype, void* stream) {
    MOREC_CHECK_ARG(ids && pos && word && posemb && type0 && out, "bert_embed_fwd: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0, "bert_embed_fwd: H %% 4 != 0");
    if (n_tok <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n_tok * (H / 4), 256);
    MOREC_DISPATCH_T(dtype, (bert_embed_fwd_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>(ids, pos, word, posemb, type0, (T*)out, n_tok, H)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_bert_embed_bwd(const void* dz, const int64_t* ids, const int32_t* pos, float* dword,
                                    float* dposemb, int n_tok, int H, int dtype, void* stream) {
    MOREC_CHECK_ARG(dz && ids && pos, "bert_embed_bwd: null pointer");
    if (n_tok <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n_tok * (H / 4), 256);
    MOREC_DISPATCH_T(dtype, (bert_embed_bwd_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)dz, ids, pos, dword, dposemb, n_tok, H)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_gather_rows(const void* src, const int32_t* idx, void* dst, int n, int H, int ld_src, int ld_dst,
                                 int src_dtype, int dst_dtype, void* stream) {
    MOREC_CHECK_ARG(src && idx && dst, "gather_rows: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, "gather_rows: H/ld must be multiples of 4");
    if (n <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n * (H / 4), 256);
    cudaStream_t st = (cudaStream_t)stream;
    MOREC_CHECK_ARG(!(MOREC_DT_IS16(src_dtype) && MOREC_DT_IS16(dst_dtype) && src_dtype != dst_dtype),
                    "gather_rows: bf16 <-> fp16 conversion is not supported");
    if (src_dtype == 0 || src_dtype == 2) {
        MOREC_DISPATCH_T(dst_dtype, (gather_rows_kernel<float, T><<<g, 256, 0, st>>>((const float*)src, idx, (T*)dst, n, H, ld_src, ld_dst)));
    } else if (dst_dtype == 0 || dst_dtype == 2) {
        MOREC_DISPATCH_T(src_dtype, (gather_rows_kernel<T, float><<<g, 256, 0, st>>>((const T*)src, idx, (float*)dst, n, H, ld_src, ld_dst)));
    } else {
        MOREC_DISPATCH_T(src_dtype, (gather_rows_kernel<T, T><<<g, 256, 0, st>>>((const T*)src, idx, (T*)dst, n, H, ld_src, ld_dst)));
    }
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_mask_row_lens(const int64_t* text, int ld, int T, int n, int32_t* lens, void* stream) {
    MOREC_CHECK_ARG(text && lens, "mask_row_lens: null pointer");
    MOREC_CHECK_ARG(T > 0 && ld >= 2 * T, "mask_row_lens: rows must hold T ids followed by T mask entries");
    if (n <= 0) return MOREC_OK;
    mask_row_lens_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(text, ld, T, n, lens);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_pack_tokens(const int64_t* text, int ld, int T, const int32_t* enc_rows, const int32_t* cu,
                                 int n_enc, int64_t* tok_ids, int32_t* tok_pos, void* stream) {
    MOREC_CHECK_ARG(text && enc_rows && cu && tok_ids && tok_pos, "pack_tokens: null pointer");
    MOREC_CHECK_ARG(T > 0 && ld >= 2 * T, "pack_tokens: rows must hold T ids followed by T mask entries");
    if (n_enc <= 0) return MOREC_OK;
    pack_tokens_kernel<<<(n_enc + 7) / 8, 256, 0, (cudaStream_t)stream>>>(text, ld, T, enc_rows, cu, n_enc, tok_ids, tok_pos);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_scatter_add_rows(const void* src, const int32_t* idx, float* dst, int n, int H, int ld_src,
                                      int ld_dst, int src_dtype, void* stream) {
    MOREC_CHECK_ARG(src && idx && dst, "scatter_add_rows: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && ld_src % 4 == 0, "scatter_add_rows: H/ld must be multiples of 4");
    if (n <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n * (H / 4), 256);
    MOREC_DISPATCH_T(src_dtype, (scatter_add_rows_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)src, idx, dst, n, H, ld_src, ld_dst)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_scale_add_rows(const void* x, const void* y, const int32_t* idx, const float* group_scale,
                                    int rows_per_group, float alpha, void* out, int n, int H, int ld_y, int dtype,
                                    void* stream) {
    MOREC_CHECK_ARG(y && out, "scale_add_rows: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && ld_y % 4 == 0, "scale_add_rows: H/ld must be multiples of 4");
    if (n <= 0) return MOREC_OK;
    if (rows_per_group <= 0) rows_per_group = 1;
    const int g = grid_for((size_t)n * (H / 4), 256);
    MOREC_DISPATCH_T(dtype, (scale_add_rows_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)y, idx, group_scale, rows_per_group, alpha, (T*)out, n, H, ld_y)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_mean_rows(const void* x, void* out, int n_groups, int rows_per_group, int H, int dtype, void* stream) {
    MOREC_CHECK_ARG(x && out, "mean_rows: null pointer");
    MOREC_CHECK_ARG(H % 4 == 0 && rows_per_group > 0, "mean_rows: bad shape");
    if (n_groups <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n_groups * (H / 4), 256);
    MOREC_DISPATCH_T(dtype, (mean_rows_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, n_groups, rows_per_group, H)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_colsum(const void* x, float* out, int M, int N, int ld, int dtype, void* stream) {
    MOREC_CHECK_ARG(x && out, "colsum: null pointer");
    MOREC_CHECK_ARG(N % 4 == 0 && ld % 4 == 0, "colsum: N/ld must be multiples of 4");
    if (M <= 0) return MOREC_OK;
    const int gx = (N / 4 + 31) / 32;
    int gy = (num_sms() * 4 + gx - 1) / gx;
    int rows_per = (M + gy - 1) / gy;
    if (rows_per < 64) rows_per = 64;
    gy = (M + rows_per - 1) / rows_per;
    dim3 grid(gx, gy);
    MOREC_DISPATCH_T(dtype, MOREC_CUDA(launch_pdl(colsum_kernel<T>, grid, dim3(256), 0, (cudaStream_t)stream, (const T*)x, out, M, N, ld, rows_per)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_act_bwd(const void* dy, const void* aux, void* out, int64_t n, int mode, int dtype, void* stream) {
    MOREC_CHECK_ARG(dy && aux && out, "act_bwd: null pointer");
    MOREC_CHECK_ARG(n % 4 == 0, "act_bwd: n %% 4 != 0");
    if (n <= 0) return MOREC_OK;
    const int g = grid_for((size_t)n / 4, 256);
    MOREC_DISPATCH_T(dtype, (act_bwd_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)dy, (const T*)aux, (T*)out, (size_t)n / 4, mode)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_cast_f32_to_16(const float* src, void* dst, int64_t n, int dst_dtype, void* stream) {
    MOREC_CHECK_ARG(src && dst, "cast: null pointer");
    MOREC_CHECK_ARG(n % 4 == 0, "cast: n %% 4 != 0");
    MOREC_CHECK_ARG(MOREC_DT_IS16(dst_dtype), "cast: dst_dtype must be 1 (bf16) or 3 (fp16)");
    if (n <= 0) return MOREC_OK;
    MOREC_DISPATCH_T(dst_dtype, (cast_f32_to_16_kernel<T><<<grid_for((size_t)n / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, (size_t)n / 4)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_cast_chunk_elems(void) { return CAST_CHUNK; }

extern "C" int morec_cast_f32_to_16_multi(const void* tensors, const int32_t* chunk_start, int n_tensors, int n_chunks,
                                          int dst_dtype, void* stream) {
    MOREC_CHECK_ARG(tensors && chunk_start, "cast_multi: null table");
    static_assert(sizeof(CastTensor) == sizeof(MorecCastTensor), "ABI struct mismatch");
    if (n_chunks <= 0 || n_tensors <= 0) return MOREC_OK;
    const int grid = n_chunks < num_sms() * 8 ? n_chunks : num_sms() * 8;
    MOREC_CHECK_ARG(MOREC_DT_IS16(dst_dtype), "cast_multi: dst_dtype must be 1 (bf16) or 3 (fp16)");
    MOREC_DISPATCH_T(dst_dtype, (cast_multi_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const CastTensor*)tensors, chunk_start, n_tensors, n_chunks)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

// table: device buffer = n_tensors MorecAdamTensor records followed by (n_tensors + 1) int32 chunk offsets
extern "C" int morec_adamw_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int morec_adamw_multi(const void* tensors, const int32_t* chunk_start, int n_tensors, int n_chunks,
                                 float* step, const float* grad_scale, float* found_inf, int check_finite,
                                 int p16_dtype, void* stream) {
    MOREC_CHECK_ARG(tensors && chunk_start && step, "adamw_multi: null table / step");
    static_assert(sizeof(AdamTensor) == sizeof(MorecAdamTensor), "ABI struct mismatch");
    if (n_chunks <= 0 || n_tensors <= 0) return MOREC_OK;
    const int grid = n_chunks < num_sms() * 8 ? n_chunks : num_sms() * 8;
    if (check_finite && found_inf) {
        grad_check_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const AdamTensor*)tensors, chunk_start, n_tensors,
                                                                 n_chunks, found_inf);
        MOREC_LAUNCH_CHECK();
    }
    adam_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, found_inf);
    MOREC_LAUNCH_CHECK();
    if (p16_dtype == 3)
        adamw_multi_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const AdamTensor*)tensors, chunk_start, n_tensors,
                                                                          n_chunks, step, grad_scale, found_inf);
    else
        adamw_multi_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const AdamTensor*)tensors, chunk_start, n_tensors,
                                                                                 n_chunks, step, grad_scale, found_inf);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_clock_probe(uint64_t* out_cycles_ns, void* stream) {
    MOREC_CHECK_ARG(out_cycles_ns, "clock_probe: null output");
    clock_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)out_cycles_ns);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}
