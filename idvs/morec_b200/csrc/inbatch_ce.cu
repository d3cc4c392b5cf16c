// In-batch popularity-debiased softmax cross-entropy, fused with the scoring GEMM (tcgen05).
//
// Replaces model/model.py:45-67 of the reference: label build (:45-48), logits = P.E^T - log p (:49-50), pad-column
// mask (:51-52), the per-user Python double loop building the reject mask (:53-63) and nn.CrossEntropyLoss over the
// valid rows (:65-67).  Closed form (pinned bit-exact by tests/test_oracle.py):
//     masked(r,c) = (id_c == 0) or (id_c in ids(user(r)) and c != target(r));   masked logits := -1e4
//     loss = mean_{valid r} [ logsumexp_c S[r,c] - S[r,target(r)] ]
//
// Kernels
//   1. inbatch_mask_kernel  (integer, bit-exact): per user a bit row member[b][c] = (id_c in ids(b)), plus pad bits.
//   2. scoring GEMM with the CE-partials epilogue: the [R,C] logits never leave TMEM/registers; every (row, column
//      tile) writes (max, sum-exp) and the tile that holds the row's target column writes the target logit.
//   3. inbatch_ce_combine_kernel: merges the partials -> row lse, row loss, mean loss over valid rows.
//   4. backward: the same GEMM with the dlogits epilogue writes dS = (softmax - onehot) * valid * g / n_valid, then
//      dP = dS.E and dE = dS^T.P run on the standard GEMM (MN-major operands, no transposes).
#include "../../../include/morec_b200.h"
#include "gemm2_tcgen05.cuh"

namespace morec {

constexpr float kNegMask = -1e4f;

// ------------------------------------------------------------------------------------------------
// 1. membership bit matrix
// ------------------------------------------------------------------------------------------------
// row_ids [B, L+1] : ids of the LOCAL users (rows);  col_ids [C] : ids of all score columns (local or all-gathered)
// member [B, Wc] (Wc = ceil(C/32)) ; pad [Wc]
__global__ void inbatch_mask_kernel(const int64_t* __restrict__ row_ids, const int64_t* __restrict__ col_ids,
                                    uint32_t* __restrict__ member, uint32_t* __restrict__ pad, int B, int Lp1, int C,
                                    int Wc) {
    extern __shared__ int64_t own[];   // [Lp1]
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < Lp1; t += blockDim.x) own[t] = row_ids[(size_t)b * Lp1 + t];
    __syncthreads();
    for (int c = threadIdx.x; c < Wc * 32; c += blockDim.x) {
        bool m = false, pd = false;
        if (c < C) {
            const int64_t id = col_ids[c];
            pd = id == 0;
            for (int t = 0; t < Lp1; ++t) m |= (own[t] == id);
        }
        const uint32_t mb = __ballot_sync(0xffffffffu, m);
        const uint32_t pb = __ballot_sync(0xffffffffu, pd);
        if ((threadIdx.x & 31) == 0) {
            member[(size_t)b * Wc + (c >> 5)] = mb;
            if (b == 0) pad[c >> 5] = pb;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2/4. GEMM epilogues
// ------------------------------------------------------------------------------------------------
struct CeParams {
    const uint32_t* member;   // [B, Wc]
    const uint32_t* pad;      // [Wc]
    const float* log_pop;     // [C] log p(id_c)
    int L, Wc, col_offset;    // target(r) = col_offset + (r/L)*(L+1) + r%L + 1
    // forward
    float* part_m; float* part_l; float* tgt_logit; int NT;
    // backward
    const float* row_lse; const float* log_mask; const float* grad_out; const float* n_valid;
};

struct CeFwdEpi {
    using Params = CeParams;
    static constexpr int kGroups = 1;      // row statistics span the whole tile: one warp per row quarter
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
        const int b = rr / ep.L;
        const int tgt = ep.col_offset + b * (ep.L + 1) + (rr - b * ep.L) + 1;
        int c_end = (s.N - n0 + 31) / 32;
        if (c_end > BLOCK_N / 32) c_end = BLOCK_N / 32;
        float m = -INFINITY, l = 0.f;
#pragma unroll 1
        for (int c = 0; c < c_end; ++c) {
            const int col0 = n0 + c * 32;
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tc_wait_ld();
            const uint32_t mw = ep.member[(size_t)b * ep.Wc + (col0 >> 5)];
            const uint32_t pw = ep.pad[col0 >> 5];
            float x[32];
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                float xv = __uint_as_float(v[j]) - (col < s.N ? __ldg(ep.log_pop + col) : 0.f);
                const bool masked = ((pw >> j) & 1u) || (((mw >> j) & 1u) && col != tgt);
                xv = masked ? kNegMask : xv;
                if (col == tgt && row < s.M) ep.tgt_logit[row] = xv;
                xv = col < s.N ? xv : -INFINITY;
                x[j] = xv;
                cm = fmaxf(cm, xv);
            }
            const float nm = fmaxf(m, cm);
            float add = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) add += __expf(x[j] - nm);
            l = l * __expf(m - nm) + add;
            m = nm;
        }
        if (row < s.M) {
            const int nt = n0 / BLOCK_N;
            ep.part_m[(size_t)row * ep.NT + nt] = m;
            ep.part_l[(size_t)row * ep.NT + nt] = l;
        }
    }
};

struct CeBwdEpi {
    using Params = CeParams;
    static constexpr int kGroups = 1;
    static constexpr bool kAuxMode = false;
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void pre_tile(const Params&, const CUtensorMap&, EpiStore&, int, int, int,
                                                    const TileSched&, int, int) {}
    template <int BLOCK_N>
    __device__ __forceinline__ static void prefetch(const Params&, int, int, const TileSched&) {}
    template <int KIND, int BLOCK_N>
    __device__ __forceinline__ static void tile(const Params& ep, const CUtensorMap& tmC, const CUtensorMap&,
                                                uint32_t taddr, EpiStore& st, int m0, int q, int n0, int,
                                                const TileSched& s, int, int) {
        const int row0 = m0 + q * 32;
        if (row0 >= s.M) return;
        const int row = row0 + st.lane;
        const int rr = row < s.M ? row : s.M - 1;
        const int b = rr / ep.L;
        const int tgt = ep.col_offset + b * (ep.L + 1) + (rr - b * ep.L) + 1;
        const float lse = ep.row_lse[rr];
        const float wrow = (row < s.M && ep.log_mask[rr] != 0.f) ? __ldg(ep.grad_out) / __ldg(ep.n_valid) : 0.f;
        int c_end = (s.N - n0 + 31) / 32;
        if (c_end > BLOCK_N / 32) c_end = BLOCK_N / 32;
        if (s.out_bf16) c_end = (c_end + 1) & ~1;
        st.c_end = c_end;
#pragma unroll 1
        for (int c = 0; c < c_end; ++c) {
            const int col0 = n0 + c * 32;
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tc_wait_ld();
            const bool inb = col0 < s.N;
            const uint32_t mw = inb ? ep.member[(size_t)b * ep.Wc + (col0 >> 5)] : 0u;
            const uint32_t pw = inb ? ep.pad[col0 >> 5] : 0u;
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                const float xv = __uint_as_float(v[j]) - (col < s.N ? __ldg(ep.log_pop + col) : 0.f);
                const bool masked = ((pw >> j) & 1u) || (((mw >> j) & 1u) && col != tgt);
                float pr = masked ? 0.f : __expf(xv - lse);      // exp(-1e4 - lse) underflows to exactly 0 in fp32
                if (col == tgt) pr -= 1.f;
                x[j] = col < s.N ? pr * wrow : 0.f;
            }
            st.emit(&tmC, x, c, n0, row0, s.out_bf16 != 0, false);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// 3. combine
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) inbatch_ce_combine_kernel(const float* __restrict__ part_m,
                                                                  const float* __restrict__ part_l,
                                                                  const float* __restrict__ tgt_logit,
                                                                  const float* __restrict__ log_mask,
                                                                  float* __restrict__ row_lse,
                                                                  float* __restrict__ row_loss,
                                                                  float* __restrict__ out /* [2]: sum, n_valid */,
                                                                  float* __restrict__ loss, int R, int NT) {
    __shared__ float ssum[32], scnt[32];
    float sum = 0.f, cnt = 0.f;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float M = -INFINITY;
        for (int t = 0; t < NT; ++t) M = fmaxf(M, part_m[(size_t)r * NT + t]);
        float acc = 0.f;
        for (int t = 0; t < NT; ++t) acc += part_l[(size_t)r * NT + t] * __expf(part_m[(size_t)r * NT + t] - M);
        const float lse = M + logf(acc);
        const float lr = lse - tgt_logit[r];
        row_lse[r] = lse;
        const bool valid = log_mask[r] != 0.f;
        row_loss[r] = valid ? lr : 0.f;
        if (valid) { sum += lr; cnt += 1.f; }
    }
    sum = warp_sum(sum);
    cnt = warp_sum(cnt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        sum = lane < (blockDim.x >> 5) ? ssum[lane] : 0.f;
        cnt = lane < (blockDim.x >> 5) ? scnt[lane] : 0.f;
        sum = warp_sum(sum);
        cnt = warp_sum(cnt);
        if (lane == 0) {
            out[0] = sum;
            out[1] = cnt;
            if (loss) *loss = sum / cnt;
        }
    }
}

}  // namespace morec

using namespace morec;

extern "C" int morec_inbatch_mask(const int64_t* row_ids, const int64_t* col_ids, uint32_t* member, uint32_t* pad,
                                  int B, int L, int C, void* stream) {
    MOREC_CHECK_ARG(row_ids && col_ids && member && pad, "inbatch_mask: null pointer");
    MOREC_CHECK_ARG(B > 0 && L > 0 && C > 0, "inbatch_mask: empty batch");
    const int Wc = (C + 31) / 32;
    inbatch_mask_kernel<<<B, 256, (L + 1) * sizeof(int64_t), (cudaStream_t)stream>>>(row_ids, col_ids, member, pad, B,
                                                                                     L + 1, C, Wc);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

static int ce_tiles(int C, int dtype) {
    const int bn = gemm_block_n(C, dtype);
    return (C + bn - 1) / bn;
}

extern "C" int morec_inbatch_ce_num_tiles(int C, int dtype) { return ce_tiles(C, dtype); }

extern "C" int morec_inbatch_ce_fwd(const void* P, const void* E, const uint32_t* member, const uint32_t* pad,
                                    const float* log_pop, const float* log_mask, int B, int L, int D, int C,
                                    int col_offset, int dtype, float* part_m, float* part_l, float* tgt_logit,
                                    float* row_lse, float* row_loss, float* sum_cnt, float* loss, void* stream) {
    MOREC_CHECK_ARG(P && E && member && pad && log_pop && log_mask && part_m && part_l && tgt_logit && row_lse &&
                        row_loss && sum_cnt,
                    "inbatch_ce_fwd: null pointer");
    const int R = B * L;
    GemmArgs g{};
    g.A = P; g.B = E; g.C = nullptr; g.C2 = nullptr;
    g.M = R; g.N = C; g.K = D; g.lda = D; g.ldb = D; g.ldc = C;
    g.dtype = dtype;
    CeParams ep{};
    ep.member = member; ep.pad = pad; ep.log_pop = log_pop; ep.L = L; ep.Wc = (C + 31) / 32; ep.col_offset = col_offset;
    ep.part_m = part_m; ep.part_l = part_l; ep.tgt_logit = tgt_logit; ep.NT = ce_tiles(C, dtype);
    if (int rc = gemm_dispatch_auto<CeFwdEpi>(g, ep, (cudaStream_t)stream)) return rc;
    inbatch_ce_combine_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(part_m, part_l, tgt_logit, log_mask, row_lse,
                                                                   row_loss, sum_cnt, loss, R, ep.NT);
    MOREC_LAUNCH_CHECK();
    return MOREC_OK;
}

extern "C" int morec_inbatch_ce_dlogits(const void* P, const void* E, const uint32_t* member, const uint32_t* pad,
                                        const float* log_pop, const float* log_mask, const float* row_lse,
                                        const float* grad_out, const float* n_valid, int B, int L, int D, int C,
                                        int col_offset, int dtype, void* dS, int ldds, void* stream) {
    MOREC_CHECK_ARG(P && E && member && pad && log_pop && log_mask && row_lse && grad_out && n_valid && dS,
                    "inbatch_ce_dlogits: null pointer");
    GemmArgs g{};
    g.A = P; g.B = E; g.C = dS; g.C2 = nullptr;
    g.M = B * L; g.N = C; g.K = D; g.lda = D; g.ldb = D; g.ldc = ldds;
    g.dtype = dtype; g.out_bf16 = MOREC_DT_IS16(dtype); g.out_f16 = dtype == 3;
    CeParams ep{};
    ep.member = member; ep.pad = pad; ep.log_pop = log_pop; ep.L = L; ep.Wc = (C + 31) / 32; ep.col_offset = col_offset;
    ep.row_lse = row_lse; ep.log_mask = log_mask; ep.grad_out = grad_out; ep.n_valid = n_valid;
    return gemm_dispatch_auto<CeBwdEpi>(g, ep, (cudaStream_t)stream);
}
