// Full-catalogue evaluation: rank of the held-out item among all catalogue items, fused with the scoring GEMM.
//
// Replaces data_utils/metrics.py:77-107 (eval_model) + :49-57 (metrics_topK) of the reference, which per user
//   scores = prec_emb . item_embeddings^T ; scores[history] = -inf ; scores = scores[1:] ;
//   order = argsort(scores, descending) ; rank = position of the target in `order`          (one argsort over N items
//   and one dense N-long one-hot label PER USER, data_utils/dataset.py:60-61).
// Closed form used here (identical for tie-free scores; ties resolve like a STABLE descending sort):
//   rank(u) = 1 + #{ c in [1, N] : c not in history(u), c != target(u),
//                    s[u,c] > s[u,target]  or  (s[u,c] == s[u,target] and c < target(u)) }
// The [U, N+1] score matrix never leaves TMEM / registers: the GEMM epilogue compares each accumulator with the row's
// target score and adds the per-tile count to count[u] (int32 atomics).  History membership is a bit matrix
// hist_bits[U, ceil((N+1)/32)] built by morec_eval_hist_bits (column 0 = the pad item is always masked: `scores[1:]`).
// The target score comes from a small pre-pass GEMM over the gathered target rows (same mainloop, same K order), and
// the big pass reports the score it saw in the target column (tgt_seen) so the caller can verify bit-equality.
#include "../../../include/morec_b200.h"
#include "gemm2_tcgen05.cuh"

namespace morec {

// one thread per history entry: set bit (u, item); thread u of the first U threads also masks column 0
__global__ void eval_hist_bits_kernel(const int32_t* __restrict__ hist_ptr, const int64_t* __restrict__ hist_items,
                                      int U, int n_cols, int Wc, uint32_t* __restrict__ bits) {
    const int total = hist_ptr[U];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + U; i += gridDim.x * blockDim.x) {
        if (i < U) {
            atomicOr(bits + (size_t)i * Wc, 1u);
            continue;
        }
        const int e = i - U;
        int lo = 0, hi = U;                   // user of entry e: hist_ptr[lo] <= e < hist_ptr[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (hist_ptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int64_t it = hist_items[e];
        if (it >= 0 && it < n_cols) atomicOr(bits + (size_t)lo * Wc + (it >> 5), 1u << (it & 31));
    }
}

struct RankParams {
    const uint32_t* hist_bits;   // [U, Wc]
    int Wc;
    const float* tgt_score;      // [U]
    const int32_t* tgt;          // [U] target column (item id)
    int* count;                  // [U] += items ranked before the target
    float* tgt_seen;             // [U] score of the target column as computed by THIS pass (may be null)
};

struct RankEpi {
    using Params = RankParams;
    static constexpr int kGroups = 1;
    static constexpr bool kAuxMode = false;
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void pre_tile(const Params&, const CUtensorMap&, EpiStore&, int, int, int,
                                                    const TileSched&, int, int) {}
    template <int BLOCK_N>
    __device__ __forceinline__ static void prefetch(const Params&, int, int, const TileSched&) {}
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void tile(const Params& ep, const CUtensorMap&, const CUtensorMap&, uint32_t taddr,
                                                EpiStore& st, int m0, int q, int n0, int, const TileSched& s, int, int) {
        const int row0 = m0 + q * 32;
        if (row0 >= s.M) return;
        const int row = row0 + st.lane;
        const int rr = row < s.M ? row : s.M - 1;
        const float t = __ldg(ep.tgt_score + rr);
        const int tg = __ldg(ep.tgt + rr);
        int c_end = (s.N - n0 + 31) / 32;
        if (c_end > BLOCK_N / 32) c_end = BLOCK_N / 32;
        int cnt = 0;
        float seen = 0.f;
        bool have_seen = false;
        // one 32-column chunk: v[j] = s[row, col0 + j]
        auto chunk = [&](const uint32_t (&v)[32], int col0) {
            uint32_t live = ~ep.hist_bits[(size_t)rr * ep.Wc + (col0 >> 5)];
            if (col0 + 32 > s.N) live &= (1u << (s.N - col0)) - 1u;          // (s.N - col0 in [1, 31] here)
            const int tj = tg - col0;                                       // target's position in this chunk, if any
            if (tj >= 0 && tj < 32) { live &= ~(1u << tj); have_seen = true; }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = __uint_as_float(v[j]);
                seen = (j == tj) ? x : seen;
                const bool before = x > t || (x == t && col0 + j < tg);
                cnt += (before && ((live >> j) & 1u)) ? 1 : 0;
            }
        };
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < c_end; c += 2) {
            tc_wait_ld();
            if (c + 1 < c_end) tmem_ld32(taddr + (c + 1) * 32, vb);
            chunk(va, n0 + c * 32);
            if (c + 1 < c_end) {
                tc_wait_ld();
                if (c + 2 < c_end) tmem_ld32(taddr + (c + 2) * 32, va);
                chunk(vb, n0 + (c + 1) * 32);
            }
        }
        if (have_seen && ep.tgt_seen && row < s.M) ep.tgt_seen[row] = seen;
        if (row < s.M && cnt) atomicAdd(ep.count + row, cnt);
    }
};

}  // namespace morec

using namespace morec;

extern "C" int morec_eval_hist_bits(const int32_t* hist_ptr, const int64_t* hist_items, int U, int n_hist, int n_cols,
                                    uint32_t* bits, void* stream) {
    MOREC_CHECK_ARG(hist_ptr && bits && (hist_items || n_hist == 0), "eval_hist_bits: null pointer");
    MOREC_CHECK_ARG(U > 0 && n_cols > 0 && n_hist >= 0, "eval_hist_bits: empty problem");
    const int Wc = (n_cols + 31) / 32;
    MOREC_CUDA(cudaMemsetAsync(bits, 0, (size_t)U * Wc * sizeof(uint32_t), (cudaStream_t)stream));
    const int total = n_hist + U;
    int blocks = (total + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    eval_hist_bits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(hist_ptr, hist_items, U, n_cols, Wc, bits);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_eval_rank(const void* P, const void* E, const uint32_t* hist_bits, const float* tgt_score,
                               const int32_t* tgt, int U, int n_cols, int D, int dtype, int32_t* count, float* tgt_seen,
                               void* stream) {
    MOREC_CHECK_ARG(P && E && hist_bits && tgt_score && tgt && count, "eval_rank: null pointer");
    MOREC_CHECK_ARG(U > 0 && n_cols > 0 && D > 0, "eval_rank: empty problem");
    MOREC_CUDA(cudaMemsetAsync(count, 0, (size_t)U * sizeof(int32_t), (cudaStream_t)stream));
    GemmArgs g{};
    g.A = P; g.B = E; g.C = nullptr; g.C2 = nullptr;
    g.M = U; g.N = n_cols; g.K = D; g.lda = D; g.ldb = D; g.ldc = n_cols;
    g.dtype = dtype;
    RankParams ep{};
    ep.hist_bits = hist_bits; ep.Wc = (n_cols + 31) / 32; ep.tgt_score = tgt_score; ep.tgt = tgt; ep.count = count;
    ep.tgt_seen = tgt_seen;
    return gemm_dispatch_auto<RankEpi>(g, ep, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// BCE head of the bce_* packages (bce_text/main-end2end/model/model.py:44-51): the same encoder + SASRec kernels with
// a loss that needs no [R, C] matrix --
//   pos[r] = <P[r], Epos[r]>, neg[r] = <P[r], Eneg[r]>,
//   loss = mean_valid softplus(-pos) + mean_valid softplus(neg)            (two nn.BCEWithLogitsLoss means)
// One warp per row; fp32 math; inputs fp32 / bf16 / fp16.
// ------------------------------------------------------------------------------------------------
namespace morec {

template <typename T>
__global__ void __launch_bounds__(256) bce_fwd_kernel(const T* __restrict__ P, const T* __restrict__ Epos,
                                                      const T* __restrict__ Eneg, const float* __restrict__ log_mask,
                                                      int R, int D, float* __restrict__ pos, float* __restrict__ neg,
                                                      float* __restrict__ sum_cnt /* [2], zero-initialised */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float lsum = 0.f, lcnt = 0.f;
    for (int r = blockIdx.x * 8 + wid; r < R; r += gridDim.x * 8) {
        float dp = 0.f, dn = 0.f;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 p = ld4<T>(P + (size_t)r * D + c);
            const float4 a = ld4<T>(Epos + (size_t)r * D + c);
            const float4 b = ld4<T>(Eneg + (size_t)r * D + c);
            dp += p.x * a.x + p.y * a.y + p.z * a.z + p.w * a.w;
            dn += p.x * b.x + p.y * b.y + p.z * b.z + p.w * b.w;
        }
        dp = warp_sum(dp);
        dn = warp_sum(dn);
        if (lane == 0) {
            pos[r] = dp;
            neg[r] = dn;
            if (log_mask[r] != 0.f) {
                // softplus(x) = max(x, 0) + log1p(exp(-|x|))  (the stable form torch uses)
                lsum += fmaxf(-dp, 0.f) + log1pf(__expf(-fabsf(dp))) + fmaxf(dn, 0.f) + log1pf(__expf(-fabsf(dn)));
                lcnt += 1.f;
            }
        }
    }
    __shared__ float ss[8], sc[8];
    if (lane == 0) { ss[wid] = lsum; sc[wid] = lcnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, c = 0.f;
        for (int i = 0; i < 8; ++i) { s += ss[i]; c += sc[i]; }
        if (c > 0.f) { atomicAdd(sum_cnt, s); atomicAdd(sum_cnt + 1, c); }
    }
}

// dP = gp*Epos + gn*Eneg ; dEpos = gp*P ; dEneg = gn*P with gp = (sigmoid(pos)-1)*g/n, gn = sigmoid(neg)*g/n on valid rows
template <typename T>
__global__ void __launch_bounds__(256) bce_bwd_kernel(const T* __restrict__ P, const T* __restrict__ Epos,
                                                      const T* __restrict__ Eneg, const float* __restrict__ log_mask,
                                                      const float* __restrict__ pos, const float* __restrict__ neg,
                                                      const float* __restrict__ grad_out, const float* __restrict__ sum_cnt,
                                                      int R, int D, T* __restrict__ dP, T* __restrict__ dEpos,
                                                      T* __restrict__ dEneg) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float w = __ldg(grad_out) / fmaxf(__ldg(sum_cnt + 1), 1.f);
    for (int r = blockIdx.x * 8 + wid; r < R; r += gridDim.x * 8) {
        const bool valid = log_mask[r] != 0.f;
        const float gp = valid ? (1.f / (1.f + __expf(-pos[r])) - 1.f) * w : 0.f;
        const float gn = valid ? (1.f / (1.f + __expf(-neg[r]))) * w : 0.f;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 p = ld4<T>(P + (size_t)r * D + c);
            const float4 a = ld4<T>(Epos + (size_t)r * D + c);
            const float4 b = ld4<T>(Eneg + (size_t)r * D + c);
            st4<T>(dP + (size_t)r * D + c, make_float4(gp * a.x + gn * b.x, gp * a.y + gn * b.y, gp * a.z + gn * b.z, gp * a.w + gn * b.w));
            st4<T>(dEpos + (size_t)r * D + c, make_float4(gp * p.x, gp * p.y, gp * p.z, gp * p.w));
            st4<T>(dEneg + (size_t)r * D + c, make_float4(gn * p.x, gn * p.y, gn * p.z, gn * p.w));
        }
    }
}

}  // namespace morec

extern "C" int morec_bce_fwd(const void* P, const void* Epos, const void* Eneg, const float* log_mask, int R, int D,
                             int dtype, float* pos, float* neg, float* sum_cnt, void* stream) {
    MOREC_CHECK_ARG(P && Epos && Eneg && log_mask && pos && neg && sum_cnt, "bce_fwd: null pointer");
    MOREC_CHECK_ARG(R > 0 && D > 0 && D % 4 == 0, "bce_fwd: D must be a positive multiple of 4");
    MOREC_CUDA(cudaMemsetAsync(sum_cnt, 0, 2 * sizeof(float), (cudaStream_t)stream));
    int blocks = (R + 7) / 8;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    MOREC_DISPATCH_T(dtype, (bce_fwd_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)P, (const T*)Epos, (const T*)Eneg, log_mask, R, D, pos, neg, sum_cnt)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_bce_bwd(const void* P, const void* Epos, const void* Eneg, const float* log_mask, const float* pos,
                             const float* neg, const float* grad_out, const float* sum_cnt, int R, int D, int dtype,
                             void* dP, void* dEpos, void* dEneg, void* stream) {
    MOREC_CHECK_ARG(P && Epos && Eneg && log_mask && pos && neg && grad_out && sum_cnt && dP && dEpos && dEneg,
                    "bce_bwd: null pointer");
    MOREC_CHECK_ARG(R > 0 && D > 0 && D % 4 == 0, "bce_bwd: D must be a positive multiple of 4");
    int blocks = (R + 7) / 8;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    MOREC_DISPATCH_T(dtype, (bce_bwd_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)P, (const T*)Epos, (const T*)Eneg, log_mask, pos, neg, grad_out, sum_cnt, R, D, (T*)dP, (T*)dEpos, (T*)dEneg)));
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}
