// CTA-pair (cta_group::2) variant of the tcgen05 GEMM mainloop: two CTAs of a 2-CTA cluster (one TPC) cooperate on a
// 256 x 256 output tile.  Each CTA loads its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows) per
// K-block, so the per-SM operand traffic drops from 48 KB to 32 KB per K-block (L2 -> SMEM and SMEM -> tensor core)
// and the operand ring gets 6 stages instead of 4.  The leader CTA (cluster rank 0) issues `tcgen05.mma.cta_group::2`
// (M = 256, N = 256) for the pair; accumulators live in both CTAs' TMEM (128 lanes x 256 columns each, double
// buffered) and each CTA runs the epilogue for its own 128 rows.
//
// Synchronisation (all mbarriers; SMEM offsets are identical in both CTAs):
//   full[s]   (leader's is used)   count 2: both producers arrive; the leader's arrive carries expect_tx of BOTH CTAs'
//                                   bytes, and both CTAs' TMA loads complete_tx on the leader's barrier (peer bit masked)
//   empty[s]  (per CTA)            count 1: tcgen05.commit multicast from the leader frees the slot in both CTAs
//   tfull[a]  (per CTA)            count 1: tcgen05.commit multicast when the accumulator stage is complete
//   tempty[a] (leader's is used)   count 2*4*groups: every epilogue warp of BOTH CTAs arrives (remote arrive from the peer)
#pragma once
#include <stdlib.h>

#include "gemm_tcgen05.cuh"

namespace morec {

// tools/gemm_trace.cu compiles this header with MOREC_GEMM_TRACE: CTA 0 records clock64() at the hand-over points of
// its producer / MMA / epilogue warps (no effect on the library build)
#ifdef MOREC_GEMM_TRACE
__device__ long long g_gemm_trace[6][4096];
#define MOREC_TRACE(row, idx)                                                                 \
    do {                                                                                      \
        if (blockIdx.x == 0 && (idx) < 4096) g_gemm_trace[row][idx] = clock64();              \
    } while (0)
#else
#define MOREC_TRACE(row, idx) do { } while (0)
#endif

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same SMEM offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(local_bar),
        "r"(rank)
        : "memory");
}
// ---- dynamic tile scheduling: cluster-scope release / acquire hand-over of the next work item
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t local_bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(local_bar),
        "r"(rank)
        : "memory");
}
// consumer -> scheduler "slot read": RELAXED.  A release arrive by the MMA-issuing thread waited for its outstanding
// tensor-core work (measured: ~2,000 idle cycles at every tile boundary); nothing has to be published here -- the
// consumer only read the slot, and the arrive's address depends on the value read, so it cannot overtake the load.
__device__ __forceinline__ void mbar_arrive_relaxed_cluster(uint32_t local_bar, uint32_t rank, uint32_t dep) {
    asm volatile(
        "{\n\t.reg .b32 ra, rz;\n\t"
        "and.b32 rz, %2, 0;\n\t"
        "add.u32 rz, rz, %0;\n\t"
        "mapa.shared::cluster.u32 ra, rz, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(local_bar),
        "r"(rank), "r"(dep)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAITC_DONE;\n\t"
        "bra.uni WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t local_addr, uint32_t rank, uint32_t v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}\n" ::"r"(local_addr),
        "r"(rank), "r"(v)
        : "memory");
}
// Work items are dealt DYNAMICALLY (CL == 2): the first item of a pair is its index, every further one comes from a
// global counter.  A scheduler thread (leader CTA, warp 3) draws item i+1 as soon as the producer has STARTED item i
// and publishes it through a 4-deep ring in the shared memory of BOTH CTAs:
//   pstart    (leader's, count 1)           the leader's producer arrives when it starts an item (look-ahead of one item:
//                                           a pair never hoards work it has not begun while other pairs run dry)
//   sfull[s]  (per CTA, count 1)            the scheduler arrives (release.cluster) after writing slot s in both CTAs
//   sempty[s] (leader's, count 3 + 8*groups) every consumer -- both producers, the MMA warp, all epilogue warps of both
//                                           CTAs -- arrives after reading slot s
// The hand-over must not run on the producer thread: its two cluster-scope release arrives cost ~2,000 cycles, more
// than the ~1,700 cycles by which the operand ring (6 stages, ~2,900 cycles TMA latency under load) is ahead of the
// tensor core -- measured 51.4 -> 56.9 us at 12037 x 3072 x 768 with the producer as scheduler.
// A pair whose CTAs became resident late (the SMs were held by an NCCL kernel or by a GEMM of the other stream) simply
// takes fewer items; with the static `w += n_pairs` deal such a pair runs its whole share after everybody else has
// finished.  OPT-IN (gemm_sched_slot() in gemm_tcgen05.cu has the measurements: free in isolation, no gain next to
// NCCL's kernels on 2 GPUs except when the collective is squeezed onto 2 CTAs, where it recovers 0.46 ms per step).
constexpr int kSchedRing = 4;
struct SchedRing {
    uint32_t sfull, sempty, tiles;      // shared-memory addresses (this CTA)
    uint32_t lead_rank;
    // consumer side: work item number `it` (>= 1) of this pair.  WARP: called by all 32 lanes of a converged warp (one
    // arrival per warp, after every lane has read the slot); otherwise by a single thread
    template <bool WARP>
    __device__ __forceinline__ int next(int it, int lane) const {
        const int slot = it & (kSchedRing - 1);
        const uint32_t ph = ((uint32_t)(it - 1) / kSchedRing) & 1u;
        mbar_wait_acquire_cluster(sfull + slot * 8, ph);
        int w;
        asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(w) : "r"(tiles + slot * 4) : "memory");
        if (WARP) __syncwarp();
        if (!WARP || lane == 0) mbar_arrive_relaxed_cluster(sempty + slot * 8, lead_rank, (uint32_t)w);
        return w;
    }
    // scheduler thread: hand item `w` (number `it`) to every consumer of the pair
    __device__ __forceinline__ void publish(int it, int w) const {
        const int slot = it & (kSchedRing - 1);
        const uint32_t ph = ((uint32_t)(it - 1) / kSchedRing) & 1u;
        mbar_wait_acquire_cluster(sempty + slot * 8, ph ^ 1u);
        st_shared_cluster_u32(tiles + slot * 4, lead_rank, (uint32_t)w);
        st_shared_cluster_u32(tiles + slot * 4, lead_rank + 1, (uint32_t)w);
        mbar_arrive_release_cluster(sfull + slot * 8, lead_rank);
        mbar_arrive_release_cluster(sfull + slot * 8, lead_rank + 1);
    }
};

__device__ __forceinline__ void tma_load_2d_2sm(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1) {
    // executed by both CTAs; the peer bit of the barrier address is cleared so the bytes are credited to CTA 0
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// ... and additionally written to the same SMEM offset of every CTA in `mask` (cluster ranks); the bytes are credited
// to the leader of each destination CTA's pair
__device__ __forceinline__ void tma_load_2d_2sm_mc(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}

template <int KIND>
struct Cfg2 {
    static constexpr int ELEM = KIND == 1 ? 2 : 4;
    static constexpr int BLOCK_N = 256;                 // per pair
    static constexpr int HALF_N = 128;                  // B rows loaded by each CTA
    static constexpr int PAIR_M = 256;
    static constexpr int BLOCK_K = 128 / ELEM;
    static constexpr int UMMA_K = 32 / ELEM;
    static constexpr int CHUNK = 128 / ELEM;
    static constexpr int A_BYTES = BLOCK_M * 128;
    static constexpr int B_BYTES = HALF_N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // per CTA
    static constexpr int EPI_SUBS = 16;                 // 4 KB staging sub-buffers per CTA, shared out over the epilogue warps
    static constexpr int EPI_BYTES = EPI_SUBS * kEpiBufBytes;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int STAGES_RAW = (227 * 1024 - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
};

// Epilogue warp groups: Epi::kGroups == 2 adds warps 8-11 as a second epilogue group (same TMEM lane quarters, the
// other half of the tile's columns).  One warp per scheduler could not keep up with math-heavy epilogues: the GELU
// forward / GELU-gradient epilogues (~25 instructions per element, dependent chains) paced the kernel at 143 us against
// 50 us for the same GEMM with a plain epilogue.
template <class Epi>
constexpr int gemm2_threads() { return kThreads + (Epi::kGroups == 2 ? 128 : 0); }

// CL = cluster size: 2 (one CTA pair) or 4 (two CTA pairs stacked along M that share every B tile: each CTA fetches a
// QUARTER of the B tile and multicasts it to its counterpart in the other pair, so the cluster reads 96 KB per K-block
// from L2 instead of 128 KB).  See gemm2_cluster4_enabled() for the measurement that keeps CL = 4 opt-in.
template <int KIND, class Epi, int CL>
__global__ void __launch_bounds__(gemm2_threads<Epi>(), 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const TileSched p,
             const typename Epi::Params ep) {
    using C = Cfg2<KIND>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + C::EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tfull_bar = bars + 2 * C::STAGES;
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    uint64_t* aux_bars = bars + 32;               // [8] one per epilogue warp: side-input TMA loads
    uint64_t* sfull_bar = bars + 40;              // [kSchedRing] dynamic scheduling (see SchedRing)
    uint64_t* sempty_bar = bars + 44;             // [kSchedRing]
    uint32_t* sched_tiles = reinterpret_cast<uint32_t*>(bars + 48);   // [kSchedRing]
    uint64_t* pstart_bar = bars + 52;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int PAIRS = CL / 2;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;             // rank within the CTA pair
    const uint32_t psub = crank >> 1;             // pair within the cluster
    const uint32_t lead_rank = crank & ~1u;       // cluster rank of this pair's leader
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmC2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 2);
            mbar_init(smem_u32(&empty_bar[s]), PAIRS);        // every pair that reads the slot commits to it
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), 2 * kEpiWarps * Epi::kGroups);
        }
        for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&aux_bars[s]), 1);
        for (int s = 0; s < kSchedRing; ++s) {
            mbar_init(smem_u32(&sfull_bar[s]), 1);
            mbar_init(smem_u32(&sempty_bar[s]), 3 + 2 * kEpiWarps * Epi::kGroups);
        }
        mbar_init(smem_u32(pstart_bar), 1);
        mbar_fence_init();
    }
    cluster_sync_all();                       // barrier inits of both CTAs visible before any remote arrive / TMA
    if (warp == 2) {
        tmem_alloc_2sm(smem_u32(tmem_slot), C::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                               // (barriers, TMEM and descriptor prefetch above overlap the predecessor's tail)
    pdl_trigger();

    const int pair = blockIdx.x / CL;             // work is dealt to clusters; the pairs of a cluster share (n-tile, split)
    const int n_pairs = gridDim.x / CL;
    const int total_work = p.m_tiles * p.n_tiles * p.splits;      // m_tiles counts (PAIRS x 256)-row cluster tiles here
    const bool dyn = CL == 2 && p.sched_ctr != nullptr;
    SchedRing ring;
    ring.sfull = smem_u32(sfull_bar); ring.sempty = smem_u32(sempty_bar); ring.tiles = smem_u32(sched_tiles);
    ring.lead_rank = lead_rank;

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] int tr = 0;
            int it = 0;
            for (int w = pair; w < total_work;) {
                if (dyn && leader) mbar_arrive(smem_u32(pstart_bar));         // the scheduler may draw the next item
                const int split = w % p.splits;
                const int tile = w / p.splits;
                const int m0 = ((tile / p.n_tiles) * PAIRS + (int)psub) * C::PAIR_M + (int)rank * BLOCK_M;
                const int n0 = (tile % p.n_tiles) * C::BLOCK_N + (int)rank * C::HALF_N;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    MOREC_TRACE(0, tr); ++tr;
                    const uint32_t fb = smem_u32(&full_bar[stage]);
#ifdef MOREC_TRACE_NOLOAD          /* diagnostic: MMAs on whatever the ring holds, no operand traffic at all */
                    if (leader) mbar_arrive(fb);
                    else mbar_arrive_cluster(fb, lead_rank);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    continue;
#endif
                    if (leader) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);
                    else mbar_arrive_cluster(fb, lead_rank);
                    const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint32_t sb = sa + C::A_BYTES;
                    const int k0 = kb * C::BLOCK_K;
                    if (!p.a_mn) {
                        tma_load_2d_2sm(&tmA, fb, sa, k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_M / C::CHUNK; ++c)
                            tma_load_2d_2sm(&tmA, fb, sa + c * (C::BLOCK_K * 128), m0 + c * C::CHUNK, k0);
                    }
                    if constexpr (CL == 2) {
                        if (!p.b_mn) {
                            tma_load_2d_2sm(&tmB, fb, sb, k0, n0);
                        } else {
#pragma unroll
                            for (int c = 0; c < C::HALF_N / C::CHUNK; ++c)
                                tma_load_2d_2sm(&tmB, fb, sb + c * (C::BLOCK_K * 128), n0 + c * C::CHUNK, k0);
                        }
                    } else {
                        // this CTA's quarter of the B tile -> itself and the CTA of equal pair rank in the other pair
                        const uint16_t mask = (uint16_t)(0x5u << rank);
                        constexpr int QROWS = C::HALF_N / PAIRS;                 // 64 N-rows per CTA
                        if (!p.b_mn) {
                            tma_load_2d_2sm_mc(&tmB, fb, sb + psub * (QROWS * 128), k0, n0 + (int)psub * QROWS, mask);
                        } else {
                            constexpr int NCH = C::HALF_N / C::CHUNK / PAIRS;    // MN-major chunks per CTA
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const int cc = (int)psub * NCH + c;
                                tma_load_2d_2sm_mc(&tmB, fb, sb + cc * (C::BLOCK_K * 128), n0 + cc * C::CHUNK, k0, mask);
                            }
                        }
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ++it;
                w = dyn ? ring.template next<false>(it, 0) : w + n_pairs;
            }
        }
    } else if (warp == 3 && leader && dyn) {
        // ================================ work-item scheduler (leader CTA only) ================================
        if (lane == 0) {
            int w = pair;
            for (int it = 0; w < total_work;) {
                mbar_wait(smem_u32(pstart_bar), (uint32_t)it & 1u);           // the producer has started item `it`
                w = n_pairs + atomicAdd(p.sched_ctr, 1);
                ++it;
                ring.publish(it, w);
            }
            // this pair has drawn its last item; the last pair of the grid to get here re-arms the counter slot
            if (atomicAdd(p.sched_ctr + 1, 1) == n_pairs - 1) {
                p.sched_ctr[0] = 0;
                p.sched_ctr[1] = 0;
                __threadfence();
            }
        }
    } else if (warp == 1 && leader) {
        // ================================ MMA issuer (leader CTA only) ================================
        const uint32_t fmt = KIND == 1 ? (p.in_f16 ? 0u : 1u) : 2u;      // f16 / bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                               ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(C::BLOCK_N >> 3) << 17) |
                               ((uint32_t)(C::PAIR_M >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        [[maybe_unused]] int tr = 0, trt = 0;
        int it = 0;
        for (int w = pair; w < total_work;) {
            const int split = w % p.splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
            tc_fence_after();
            if (lane == 0) MOREC_TRACE(3, trt);
            ++trt;
            const uint32_t d_tmem = tmem_base + acc * C::BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                tc_fence_after();
                if (lane == 0) MOREC_TRACE(1, tr);
                if (lane == 0) {
                    const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
                        constexpr uint32_t mn_sbo = KIND == 1 ? 1024 : 512;
                        constexpr uint32_t mn_lt = KIND == 1 ? 2 : 1;
                        const uint64_t ad = p.a_mn ? make_smem_desc(sa + k * (C::UMMA_K * 128), C::BLOCK_K * 128, mn_sbo, mn_lt)
                                                   : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = p.b_mn ? make_smem_desc(sb + k * (C::UMMA_K * 128), C::BLOCK_K * 128, mn_sbo, mn_lt)
                                                   : make_smem_desc(sb + k * 32, 16, 1024);
                        tc_mma_2sm<KIND == 1 ? 1 : 0>(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit_2sm_mc(smem_u32(&empty_bar[stage]), (uint16_t)((1u << CL) - 1));   // slot consumed by this pair: tell every CTA of the cluster
                    if (kb == kb1 - 1) tc_commit_2sm_mc(smem_u32(&tfull_bar[acc]), (uint16_t)(3u << (2 * psub)));   // accumulators ready in both CTAs of the pair
                    MOREC_TRACE(2, tr);
                }
                ++tr;
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            ++it;
            w = dyn ? ring.template next<true>(it, lane) : w + n_pairs;
        }
    } else if (warp >= 4 && warp < 4 + 4 * Epi::kGroups) {
        // ================================ epilogue (both CTAs, own 128 rows) ================================
        const int q = warp & 3;                    // TMEM lane quarter == warp % 4
        const int cg = (warp - 4) >> 2;            // column group of this warp
        constexpr int NSUB = C::EPI_SUBS / (4 * Epi::kGroups);
        EpiStore st;
        st.bufs = epi_base + (cg * 4 + q) * (NSUB * kEpiBufBytes);
        st.nsub = NSUB;
        st.grp = 0;
        st.aux_bar = smem_u32(&aux_bars[warp - 4]); st.aux_phase = 0; st.aux_groups = 0;
        st.c_end = 0;
        st.lane = lane;
        st.f16 = p.out_f16 != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        [[maybe_unused]] int trt = 0;
        if (pair < total_work) {                            // epilogue side inputs of the first tile -> L2
            const int tile = pair / p.splits;
            Epi::template prefetch<C::BLOCK_N>(ep, ((tile / p.n_tiles) * PAIRS + (int)psub) * C::PAIR_M + (int)rank * BLOCK_M + q * 32 + lane,
                                               (tile % p.n_tiles) * C::BLOCK_N, p);
        }
        int it = 0;
        for (int w = pair; w < total_work;) {
            const int split = w % p.splits;
            const int tile = w / p.splits;
            const int m0 = ((tile / p.n_tiles) * PAIRS + (int)psub) * C::PAIR_M + (int)rank * BLOCK_M;
            const int n0 = (tile % p.n_tiles) * C::BLOCK_N;
            if (!dyn && w + n_pairs < total_work) {         // ... and of the next tile, one tile time ahead (static deal only)
                const int nt = (w + n_pairs) / p.splits;
                Epi::template prefetch<C::BLOCK_N>(ep, ((nt / p.n_tiles) * PAIRS + (int)psub) * C::PAIR_M + (int)rank * BLOCK_M + q * 32 + lane,
                                                   (nt % p.n_tiles) * C::BLOCK_N, p);
            }
            Epi::template pre_tile<KIND, C::BLOCK_N>(ep, tmC2, st, m0, q, n0, p, cg, Epi::kGroups);
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            tc_fence_after();
            if (warp == 4 && lane == 0) MOREC_TRACE(4, trt);
            Epi::template tile<KIND, C::BLOCK_N>(ep, tmC, tmC2, tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::BLOCK_N, st, m0, q,
                                                 n0, split, p, cg, Epi::kGroups);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), lead_rank);    // the pair leader's barrier
            if (warp == 4 && lane == 0) MOREC_TRACE(5, trt);
            ++trt;
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            ++it;
            w = dyn ? ring.template next<true>(it, lane) : w + n_pairs;
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();                       // the peer's SMEM / TMEM stay alive until the leader's MMAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// Cluster size for a problem.  The four-CTA cluster (B multicast) is OPT-IN (MOREC_GEMM_CL4=1): measured on B200 it is
// 8-30 % SLOWER than the plain CTA pair at every BERT-base shape (12037x3072x768: 59.6 us vs 54.6 us) although its
// per-CTA cycle trace is identical -- the mainloop is not limited by L2 delivery: with the operand loads removed
// altogether (tools/gemm_trace.cu, -DMOREC_TRACE_NOLOAD) a K-block still takes 650-700 cycles against 766 with loads,
// i.e. the SS-mode MMA itself paces the kernel.  Kept for the record and for other shapes.
inline bool gemm2_cluster4_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOREC_GEMM_CL4");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int KIND, class Epi, int CL>
int gemm2_launch_cl(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream);

template <int KIND, class Epi>
int gemm2_launch(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream) {
    const int pair_tiles = (g.M + Cfg2<KIND>::PAIR_M - 1) / Cfg2<KIND>::PAIR_M;
    if (gemm2_cluster4_enabled() && pair_tiles >= 2 && (pair_tiles % 2 == 0 || pair_tiles >= 24))
        return gemm2_launch_cl<KIND, Epi, 4>(g, ep, stream);
    return gemm2_launch_cl<KIND, Epi, 2>(g, ep, stream);
}

template <int KIND, class Epi, int CL>
int gemm2_launch_cl(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream) {
    using C = Cfg2<KIND>;
    constexpr int PAIRS = CL / 2;
    constexpr int ELEM = C::ELEM;
    const bool bf = KIND == 1;
    constexpr bool SW32 = KIND != 1;
    CUtensorMap tmA, tmB, tmC, tmC2;
    int rc;
    if (!g.a_mn) rc = make_tmap_2d(&tmA, g.A, bf, g.K, g.M, (uint64_t)g.lda * ELEM, C::BLOCK_K, BLOCK_M);
    else         rc = make_tmap_2d(&tmA, g.A, bf, g.M, g.K, (uint64_t)g.lda * ELEM, C::CHUNK, C::BLOCK_K, SW32);
    if (rc) return rc;
    if (!g.b_mn) rc = make_tmap_2d(&tmB, g.B, bf, g.K, g.N, (uint64_t)g.ldb * ELEM, C::BLOCK_K, C::HALF_N / PAIRS);
    else         rc = make_tmap_2d(&tmB, g.B, bf, g.N, g.K, (uint64_t)g.ldb * ELEM, C::CHUNK, C::BLOCK_K, SW32);
    if (rc) return rc;
    const int out_elem = g.out_bf16 ? 2 : 4;
    const int out_cols = 128 / out_elem;
    if (g.C) {
        rc = make_tmap_2d(&tmC, g.C, g.out_bf16 != 0, g.N, g.M, (uint64_t)g.ldc * out_elem, out_cols, 32);
        if (rc) return rc;
    } else {
        tmC = tmA;
    }
    bool aux_tma = false;
    if (g.C2) {
        rc = make_tmap_2d(&tmC2, g.C2, g.out_bf16 != 0, g.N, g.M, (uint64_t)g.ldc * out_elem, out_cols, 32);
        if (rc) return rc;
    } else if (Epi::kAuxMode && g.aux && bf && g.out_bf16 && (g.ldaux % 8) == 0 &&
               (reinterpret_cast<uintptr_t>(g.aux) & 15) == 0) {
        // side input of the activation-gradient epilogues, staged by TMA in the boxes of the output (32 rows x 128 B)
        rc = make_tmap_2d(&tmC2, g.aux, true, g.N, g.M, (uint64_t)g.ldaux * 2, out_cols, 32);
        if (rc) return rc;
        aux_tma = true;
    } else {
        tmC2 = tmC;
    }
    TileSched p;
    p.sched_ctr = CL == 2 ? gemm_sched_slot() : nullptr;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.num_kb = (g.K + C::BLOCK_K - 1) / C::BLOCK_K;
    p.m_tiles = (g.M + C::PAIR_M * PAIRS - 1) / (C::PAIR_M * PAIRS);
    p.n_tiles = (g.N + C::BLOCK_N - 1) / C::BLOCK_N;
    p.a_mn = g.a_mn; p.b_mn = g.b_mn;
    p.accumulate = g.accumulate;
    p.out_bf16 = g.out_bf16;
    p.aux_tma = aux_tma ? 1 : 0;
    p.in_f16 = g.dtype == 3;
    p.out_f16 = g.out_f16;
    auto kern = gemm2_kernel<KIND, Epi, CL>;
    static int clusters_max = 0;              // per instantiation: co-resident clusters of this kernel on the device
    if (clusters_max == 0) {
        MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(num_sms() / CL * CL, 1, 1);
        qc.blockDim = dim3(gemm2_threads<Epi>(), 1, 1);
        qc.dynamicSmemBytes = C::SMEM_BYTES;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int n = 0;
        MOREC_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &qc));
        if (n > num_sms() / CL) n = num_sms() / CL;
        clusters_max = n > 0 ? n : 1;
    }
    const int pairs_max = clusters_max;
    int splits = 1;
    if (g.accumulate && g.allow_split_k) {
        const int tiles = p.m_tiles * p.n_tiles;
        splits = pairs_max / tiles;
        if (splits < 1) splits = 1;
        const int max_splits = p.num_kb / 8 > 0 ? p.num_kb / 8 : 1;
        if (splits > max_splits) splits = max_splits;
    }
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    p.splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int pairs = total < pairs_max ? total : pairs_max;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * pairs, 1, 1);
    cfg.blockDim = dim3(gemm2_threads<Epi>(), 1, 1);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute la[2];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = CL; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    la[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    la[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = la; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MOREC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmC2, p, ep));
    return MOREC_OK;
}

// Picks the CTA-pair kernel for large, wide problems (fp32/tf32 one-pass and bf16); everything else -- 3xTF32, narrow
// N, small M -- runs on the single-CTA kernel.  MOREC_GEMM2=0 in the environment disables the pair kernel.
inline bool gemm2_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOREC_GEMM2");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

template <class Epi>
int gemm_dispatch_auto(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream) {
    if (gemm2_enabled() && g.dtype != 2 && gemm_is_wide(g.N) && g.M >= 512 && g.M > 0 && g.N > 0 && g.K > 0) {
        if (g.dtype == 0) return gemm2_launch<0, Epi>(g, ep, stream);
        if (g.dtype == 1 || g.dtype == 3) return gemm2_launch<1, Epi>(g, ep, stream);
    }
    return gemm_dispatch<Epi>(g, ep, stream);
}

}  // namespace morec
