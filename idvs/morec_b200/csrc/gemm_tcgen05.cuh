// Persistent, warp-specialised tcgen05 GEMM mainloop for sm_100a (header: kernel template + launcher template).
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )
//
//   * operands arrive by TMA (SWIZZLE_128B) into a multi-stage shared-memory ring,
//   * one elected thread issues tcgen05.mma (kind::tf32 on fp32 storage, or kind::f16 on bf16 storage); fp32
//     accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * four epilogue warps read TMEM (tcgen05.ld 32x32b) and run the Epilogue functor (fused bias / GELU / ReLU /
//     activation-gradient / in-batch-CE partials ...); outputs are staged as 32-row x 128-byte sub-tiles in
//     swizzled shared memory and written by TMA store (or TMA reduce-add for split-K accumulation),
//   * either operand may be K-major ([rows, K] row-major) or MN-major ([K, rows] row-major), so forward, dgrad and
//     wgrad of a linear layer all run on this one kernel without materialising a transpose.
//
// Reference op sites this replaces (eager cuBLAS + separate elementwise kernels in the reference):
//   inbatch_sasrec_e2e_text/model/encoders.py:68-70 (HF BertModel linears + fc), model/modules.py:14-17,52-63
//   (SASRec linears), model/model.py:49 (scoring matmul; the CE epilogues live in inbatch_ce.cu).
#pragma once
#include "common.cuh"

namespace morec {

// swizzle32: false -> SWIZZLE_128B (16B atoms), true -> SWIZZLE_128B_ATOM_32B (the only layout tcgen05 accepts for
// MN-major 32-bit (tf32) operands: UMMA layout type SWIZZLE_128B_BASE32B, 4-row x 128B swizzle atoms)
int make_tmap_2d(CUtensorMap* map, const void* base, bool is_bf16, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle32 = false);

struct GemmArgs {
    const void* A;
    const void* B;
    void* C;       // primary output (may be null for epilogues that store nothing)
    void* C2;      // secondary output (e.g. pre-activation), same shape/ld as C
    int M, N, K;
    int lda, ldb, ldc;     // leading dimensions in elements
    int a_mn, b_mn;        // 0: operand is [rows, K] row-major (K-major); 1: [K, rows] row-major (MN-major)
    int dtype;             // 0: fp32 storage / kind::tf32, 1: bf16 storage / kind::f16, 2: fp32 / 3xTF32, 3: fp16 / kind::f16
    int out_bf16;          // output element type of C/C2: 0 fp32, non-zero 16-bit (bf16, or fp16 when out_f16)
    int out_f16 = 0;       // 16-bit outputs are IEEE half instead of bfloat16
    int accumulate;        // 1: C += result via TMA reduce-add (C must be fp32); enables split-K
    int allow_split_k;
    const void* aux = nullptr;   // epilogue side input [M, ldaux] (activation-gradient epilogues), element type of the operands
    int ldaux = 0;
};

constexpr int BLOCK_M = 128;
constexpr int kThreads = 256;        // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 spare, warps4-7 epilogue
constexpr int kEpiWarps = 4;
constexpr int kEpiBufBytes = 4096;   // one 32-row x 128B sub-tile (32 fp32 or 64 bf16 columns), SWIZZLE_128B
constexpr int kEpiBufsPerWarp = 2;      // 2 x 4 KB staging per epilogue warp: leaves room for a 4-stage operand ring at
                                        // BLOCK_N = 256 (TMA latency ~1 us x 96 B/clk/SM needs ~190 KB in flight)

struct TileSched {
    int M, N, K;
    int num_kb, kb_per_split, splits;
    int m_tiles, n_tiles;
    int a_mn, b_mn;
    int accumulate;
    int out_bf16;
    int aux_tma;           // 1: tmC2 describes the epilogue side input (16-bit) and the CTA-pair kernel stages it by TMA
    int in_f16;            // KIND 1 operands are IEEE half (instruction-descriptor format 0) instead of bfloat16 (1)
    int out_f16;           // 16-bit outputs are IEEE half
    int* sched_ctr;        // CTA-pair kernel, dynamic tile scheduling: {next work item, finished pairs} (null: static)
};

// One {work counter, finished-pairs counter} slot per launch of the CTA-pair kernel (round-robin over a small pool;
// the last pair to finish zeroes its slot).  Returns a device pointer to two ints, or null when dynamic scheduling
// is disabled (MOREC_GEMM_DYN=0).
int* gemm_sched_slot();

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2 /* SWIZZLE_128B */) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}

// KIND: 0 = fp32 storage, one kind::tf32 pass          ("tf32": ~1e-3 relative, operands truncated to 10 mantissa bits)
//       1 = bf16 storage, kind::f16
//       2 = fp32 storage, error-compensated 3xTF32     ("fp32" parity mode): each operand tile is split IN SHARED
//           MEMORY into hi = tf32-truncated(x) and lo = x - hi by four converter warps, and the tensor core
//           accumulates hi*hi + hi*lo + lo*hi into the same TMEM accumulator (~2^-21 relative, fp32 grade).
template <int KIND, int BLOCK_N>
struct Cfg {
    static constexpr bool X3 = KIND == 2;
    static constexpr int ELEM = KIND == 1 ? 2 : 4;
    static constexpr int BLOCK_K = 128 / ELEM;          // K elements per stage (one 128B swizzle row)
    static constexpr int UMMA_K = 32 / ELEM;            // 8 (tf32) / 16 (bf16)
    static constexpr int CHUNK = 128 / ELEM;            // MN elements per 128B row of an MN-major tile
    static constexpr int A_BYTES = BLOCK_M * 128;
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int LOAD_BYTES = A_BYTES + B_BYTES;                 // bytes landed by TMA per stage
    static constexpr int STAGE_BYTES = LOAD_BYTES * (X3 ? 2 : 1);        // x3: [A_hi | B_hi | A_lo | B_lo]
    static constexpr int THREADS = kThreads + (X3 ? 128 : 0);            // x3: warps 8-11 split the operands
    static constexpr int EPI_BUFS = kEpiBufsPerWarp;
    static constexpr int EPI_BYTES = kEpiWarps * EPI_BUFS * kEpiBufBytes;
    static constexpr int BAR_BYTES = 1024;
    static constexpr int SMEM_LIMIT = 227 * 1024;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;       // double-buffered fp32 accumulator
    static_assert(TMEM_COLS <= 512, "TMEM overflow");
    static_assert(STAGES >= 2, "not enough smem stages");
};

// Staging + TMA store of the epilogue output.  Each epilogue warp owns `nsub` sub-buffers of 32 rows x 128 bytes
// (SWIZZLE_128B; 32 fp32 or 64 bf16 columns each).  Chunks of 32 accumulator columns (thread == row, 32 registers) are
// written into the sub-buffers of the current GROUP; when the group is full (or the tile ends) ONE fence + ONE
// elected thread issues all its TMA stores (or TMA reduce-adds) and commits them as one bulk group.  With the default
// 4 sub-buffers a group is 128 fp32 / 256 bf16 columns of one output stream (half of that with two streams, e.g.
// GELU writing pre-activation + activation), so the per-chunk fence / sync / wait overhead of a naive epilogue is
// paid once or twice per tile instead of eight times.
struct EpiStore {
    uint8_t* bufs;            // this warp's staging sub-buffers
    int lane;
    int nsub;                 // sub-buffers owned by this warp (2 in the single-CTA kernels, 4 in the CTA-pair kernel)
    int c_end;                // number of 32-column chunks of the current tile (set by the epilogue)
    int grp;                  // flushed-group counter: selects the staging half
    uint32_t aux_bar;         // mbarrier of this warp's side-input TMA loads (0: not used by this kernel)
    uint32_t aux_phase;
    int aux_groups;           // side-input groups staged for the current tile (0: read the side input directly)
    bool f16;                 // 16-bit outputs are IEEE half (else bfloat16)
    const CUtensorMap* tm[2]; // output tensor map per stream (0: C, 1: C2)

    // The staging memory is used as TWO halves when it has at least two sub-buffers per output stream: a group is
    // flushed (one bulk group of TMA stores) while the next group fills the other half, and only the group before the
    // previous one must have been read out (wait_group.read 1).  With a single half every new group waited for the
    // store it had just issued: measured 173 us for the two-stream GELU forward against 54 us for the same GEMM with a
    // plain epilogue.
    __device__ __forceinline__ void put(const CUtensorMap* tmap, const float (&x)[32], int c, bool out_bf16, int stream,
                                        int nstreams) {
        const int cps = out_bf16 ? 2 : 1;                 // chunks per sub-buffer
        const int halves = nsub >= 2 * nstreams ? 2 : 1;
        const int S = nsub / (halves * nstreams);         // sub-buffers per stream per group
        const int G = S * cps;                            // chunks per group
        const int g = c % G;
        if (g == 0 && stream == nstreams - 1) {           // first write of a new group: the stores that last used
            if (lane == 0) {                              // this half must have finished reading the staging memory
                if (halves == 2) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
            }
            __syncwarp();
        }
        tm[stream] = tmap;
        const int half = halves == 2 ? (grp & 1) : 0;
        uint8_t* rowp = bufs + ((half * nstreams + stream) * S + g / cps) * kEpiBufBytes + lane * 128;
        if (!out_bf16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pj = j ^ (lane & 7);
                *reinterpret_cast<float4*>(rowp + pj * 16) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
        } else {
            const int half = g % cps;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int pj = (half * 4 + j) ^ (lane & 7);
                uint4 u;
                u.x = pack2_16(x[8 * j], x[8 * j + 1], f16); u.y = pack2_16(x[8 * j + 2], x[8 * j + 3], f16);
                u.z = pack2_16(x[8 * j + 4], x[8 * j + 5], f16); u.w = pack2_16(x[8 * j + 6], x[8 * j + 7], f16);
                *reinterpret_cast<uint4*>(rowp + pj * 16) = u;
            }
        }
    }
    // call once per chunk after all streams were put: flushes the group when it is full or the tile ends
    __device__ __forceinline__ void end_chunk(int c, int n0, int row0, bool out_bf16, bool reduce_add, int nstreams) {
        const int cps = out_bf16 ? 2 : 1;
        const int halves = nsub >= 2 * nstreams ? 2 : 1;
        const int S = nsub / (halves * nstreams);
        const int G = S * cps;
        const int g = c % G;
        if (g != G - 1 && c != c_end - 1) return;
        fence_proxy_async_smem();
        __syncwarp();
        const int half = halves == 2 ? (grp & 1) : 0;
        if (lane == 0) {
            const int nfilled = g / cps + 1;              // sub-buffers filled per stream
            const int col_base = n0 + (c - g) * 32;
            const int cols_per_sub = out_bf16 ? 64 : 32;
            for (int st = 0; st < nstreams; ++st) {
                for (int sb = 0; sb < nfilled; ++sb) {
                    const uint32_t src = smem_u32(bufs + ((half * nstreams + st) * S + sb) * kEpiBufBytes);
                    if (reduce_add) tma_reduce_add_2d(tm[st], src, col_base + sb * cols_per_sub, row0);
                    else tma_store_2d(tm[st], src, col_base + sb * cols_per_sub, row0);
                }
            }
            tma_store_commit();
        }
        ++grp;
    }
    // Compile-time flavour of put()/end_chunk() for the standard epilogues (two sub-buffers per warp in both kernels):
    // NS output streams, O16 = 16-bit outputs (F16: IEEE half).  One sub-buffer per stream and group, a group is one
    // chunk (fp32) or two (16-bit); NS == 1 ping-pongs between the two sub-buffers.  All index arithmetic folds to
    // shifts and masks: the runtime form above spent ~15 % of the GELU epilogue in integer division and control code.
    template <int NS, bool O16, bool F16>
    __device__ __forceinline__ void put_c(const CUtensorMap* tmap, const float (&x)[32], int c, int stream) {
        constexpr int CPS = O16 ? 2 : 1;
        constexpr bool PINGPONG = NS == 1;
        const int g = c & (CPS - 1);
        if (g == 0 && stream == NS - 1) {
            if (lane == 0) {
                if (PINGPONG) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
            }
            __syncwarp();
        }
        tm[stream] = tmap;
        const int half = PINGPONG ? (grp & 1) : 0;
        uint8_t* rowp = bufs + (half * NS + stream) * kEpiBufBytes + lane * 128;
        if constexpr (!O16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pj = j ^ (lane & 7);
                *reinterpret_cast<float4*>(rowp + pj * 16) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int pj = (g * 4 + j) ^ (lane & 7);
                uint4 u;
                u.x = pack2_16(x[8 * j], x[8 * j + 1], F16); u.y = pack2_16(x[8 * j + 2], x[8 * j + 3], F16);
                u.z = pack2_16(x[8 * j + 4], x[8 * j + 5], F16); u.w = pack2_16(x[8 * j + 6], x[8 * j + 7], F16);
                *reinterpret_cast<uint4*>(rowp + pj * 16) = u;
            }
        }
    }
    template <int NS, bool O16>
    __device__ __forceinline__ void end_chunk_c(int c, int n0, int row0, bool reduce_add) {
        constexpr int CPS = O16 ? 2 : 1;
        constexpr bool PINGPONG = NS == 1;
        const int g = c & (CPS - 1);
        if (g != CPS - 1 && c != c_end - 1) return;
        fence_proxy_async_smem();
        __syncwarp();
        const int half = PINGPONG ? (grp & 1) : 0;
        if (lane == 0) {
            const int col_base = n0 + (c - g) * 32;
#pragma unroll
            for (int st = 0; st < NS; ++st) {
                const uint32_t src = smem_u32(bufs + (half * NS + st) * kEpiBufBytes);
                if (reduce_add) tma_reduce_add_2d(tm[st], src, col_base, row0);
                else tma_store_2d(tm[st], src, col_base, row0);
            }
            tma_store_commit();
        }
        ++grp;
    }
    // compatibility helper used by single-stream epilogues
    __device__ __forceinline__ void emit(const CUtensorMap* tmap, const float (&x)[32], int c, int n0, int row0,
                                         bool out_bf16, bool reduce_add, int slot = 0, int nslots = 1) {
        put(tmap, x, c, out_bf16, slot, nslots);
        if (slot == 0) end_chunk(c, n0, row0, out_bf16, reduce_add, nslots);
    }
};

// Epilogue functor contract:
//   struct Epi { struct Params {...};
//     template <int KIND, int BLOCK_N> static __device__ void tile(const Params&, const CUtensorMap& tmC, const CUtensorMap& tmC2, uint32_t taddr,
//                                 EpiStore& st, int m0, int q, int n0, int split, const TileSched& s,
//                                 int cg, int ncg);      // this warp handles column group cg of ncg
//     static constexpr int kGroups;                      // epilogue warp groups the CTA-pair kernel should run (1 or 2) }
//     template <int BLOCK_N> static __device__ void prefetch(const Params&, int row, int n0, const TileSched&);
//   taddr already includes the warp's lane quarter and the accumulator stage; thread `lane` owns row m0+q*32+lane.
//     template <int KIND, int BLOCK_N> static __device__ void pre_tile(const Params&, const CUtensorMap& tmC2, EpiStore&,
//                                 int m0, int q, int n0, const TileSched&, int cg, int ncg);
//   pre_tile() runs right BEFORE the wait for the tile's accumulator (CTA-pair kernel only): TMA loads issued here
//   land while the tensor core still works on the tile.
//   prefetch() is called one tile ahead: it may pull the epilogue's side inputs of (row, [n0, n0+BLOCK_N)) into L2.

template <int KIND, int BLOCK_N, class Epi>
__global__ void __launch_bounds__(Cfg<KIND, BLOCK_N>::THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const TileSched p,
            const typename Epi::Params ep) {
    using C = Cfg<KIND, BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // 1024B alignment for SWIZZLE_128B
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + C::EPI_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES]
    uint64_t* empty_bar = bars + C::STAGES;           // [STAGES]
    uint64_t* tfull_bar = bars + 2 * C::STAGES;       // [2]
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
    uint64_t* conv_bar = bars + 2 * C::STAGES + 4;    // [STAGES] (x3 only): operand split done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmC2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
            mbar_init(smem_u32(&conv_bar[s]), 128);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_trigger();

    const int total_work = p.m_tiles * p.n_tiles * p.splits;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int split = w % p.splits;
                const int tile = w / p.splits;
                const int m0 = (tile / p.n_tiles) * BLOCK_M;
                const int n0 = (tile % p.n_tiles) * BLOCK_N;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, C::LOAD_BYTES);
                    const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint32_t sb = sa + C::A_BYTES;
                    const int k0 = kb * C::BLOCK_K;
                    if (!p.a_mn) {
                        tma_load_2d(&tmA, fb, sa, k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_M / C::CHUNK; ++c)
                            tma_load_2d(&tmA, fb, sa + c * (C::BLOCK_K * 128), m0 + c * C::CHUNK, k0);
                    }
                    if (!p.b_mn) {
                        tma_load_2d(&tmB, fb, sb, k0, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_N / C::CHUNK; ++c)
                            tma_load_2d(&tmB, fb, sb + c * (C::BLOCK_K * 128), n0 + c * C::CHUNK, k0);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // instruction descriptor: D=f32 [4,6), A fmt [7,10), B fmt [10,13), A major [15], B major [16],
        // N>>3 [17,23), M>>4 [24,29)
        const uint32_t fmt = KIND == 1 ? (p.in_f16 ? 0u : 1u) : 2u;   // f16 / bf16 : tf32
        constexpr int MK = KIND == 1 ? 1 : 0;           // tc_mma kind
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                               ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(C::X3 ? &conv_bar[stage] : &full_bar[stage]), phase);
                tc_fence_after();
                if (lane == 0) {   // one fixed thread issues every MMA and commit of this CTA
                    const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
                    for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
                        // K-major : 32 bytes per UMMA_K inside the 128B swizzle row; SBO = 8 rows * 128B.
                        // MN-major: UMMA_K k-rows of 128B each; LBO = next 128B chunk along MN, SBO = next 8 k-rows.
                        // (tf32 MN-major uses the 32B-atom swizzle: 4 k-rows per atom -> SBO = 512B, layout type 1.)
                        constexpr uint32_t mn_sbo = KIND == 1 ? 1024 : 512;
                        constexpr uint32_t mn_lt = KIND == 1 ? 2 : 1;
                        const uint64_t ad = p.a_mn ? make_smem_desc(sa + k * (C::UMMA_K * 128), C::BLOCK_K * 128, mn_sbo, mn_lt)
                                                   : make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = p.b_mn ? make_smem_desc(sb + k * (C::UMMA_K * 128), C::BLOCK_K * 128, mn_sbo, mn_lt)
                                                   : make_smem_desc(sb + k * 32, 16, 1024);
                        tc_mma<MK>(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        if constexpr (C::X3) {
                            // + hi*lo + lo*hi : the lo tiles sit LOAD_BYTES behind the hi tiles (start address
                            // field counts 16-byte units)
                            constexpr uint64_t lo_off = (uint64_t)(C::LOAD_BYTES >> 4);
                            tc_mma<MK>(d_tmem, ad, bd + lo_off, idesc, 1u);
                            tc_mma<MK>(d_tmem, ad + lo_off, bd, idesc, 1u);
                        }
                    }
                    tc_commit(smem_u32(&empty_bar[stage]));      // frees the smem slot when these MMAs retire
                    if (kb == kb1 - 1) tc_commit(smem_u32(&tfull_bar[acc]));   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (C::X3 && warp >= 8) {
        // ================================ operand split (3xTF32) ================================
        const int ct = threadIdx.x - 256;   // 0..127
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                float4* hi = reinterpret_cast<float4*>(stage_base + stage * C::STAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(stage_base + stage * C::STAGE_BYTES + C::LOAD_BYTES);
#pragma unroll 4
                for (int i = ct; i < C::LOAD_BYTES / 16; i += 128) {
                    const float4 x = hi[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(smem_u32(&conv_bar[stage]));
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ epilogue ================================
        const int q = warp - 4;  // TMEM lane quarter == warp % 4
        EpiStore st;
        st.bufs = epi_base + q * (C::EPI_BUFS * kEpiBufBytes);
        st.nsub = C::EPI_BUFS;
        st.grp = 0;
        st.aux_bar = 0; st.aux_phase = 0; st.aux_groups = 0;
        st.c_end = 0;
        st.lane = lane;
        st.f16 = p.out_f16 != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        if ((int)blockIdx.x < total_work) {                 // epilogue side inputs of the first tile -> L2
            const int tile = (int)blockIdx.x / p.splits;
            Epi::template prefetch<BLOCK_N>(ep, (tile / p.n_tiles) * BLOCK_M + q * 32 + lane, (tile % p.n_tiles) * BLOCK_N, p);
        }
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int tile = w / p.splits;
            const int m0 = (tile / p.n_tiles) * BLOCK_M;
            const int n0 = (tile % p.n_tiles) * BLOCK_N;
            if (w + (int)gridDim.x < total_work) {          // ... and of the next tile, one tile time ahead
                const int nt = (w + (int)gridDim.x) / p.splits;
                Epi::template prefetch<BLOCK_N>(ep, (nt / p.n_tiles) * BLOCK_M + q * 32 + lane, (nt % p.n_tiles) * BLOCK_N, p);
            }
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            tc_fence_after();
            Epi::template tile<KIND, BLOCK_N>(ep, tmC, tmC2, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N, st, m0, q,
                                              n0, split, p, 0, 1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host launcher (template; instantiated per epilogue family)
// ------------------------------------------------------------------------------------------------
template <int KIND, int BLOCK_N, class Epi>
int gemm_launch(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream) {
    using C = Cfg<KIND, BLOCK_N>;
    constexpr int ELEM = C::ELEM;
    const bool bf = KIND == 1;
    constexpr bool SW32 = KIND != 1;   // MN-major 32-bit operands need the 32B-atom swizzle
    CUtensorMap tmA, tmB, tmC, tmC2;
    int rc;
    if (!g.a_mn) rc = make_tmap_2d(&tmA, g.A, bf, g.K, g.M, (uint64_t)g.lda * ELEM, C::BLOCK_K, BLOCK_M);
    else         rc = make_tmap_2d(&tmA, g.A, bf, g.M, g.K, (uint64_t)g.lda * ELEM, C::CHUNK, C::BLOCK_K, SW32);
    if (rc) return rc;
    if (!g.b_mn) rc = make_tmap_2d(&tmB, g.B, bf, g.K, g.N, (uint64_t)g.ldb * ELEM, C::BLOCK_K, BLOCK_N);
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
    if (g.C2) {
        rc = make_tmap_2d(&tmC2, g.C2, g.out_bf16 != 0, g.N, g.M, (uint64_t)g.ldc * out_elem, out_cols, 32);
        if (rc) return rc;
    } else {
        tmC2 = tmC;
    }
    if (g.accumulate && g.out_bf16) {
        set_last_error("gemm: accumulate (TMA reduce-add) requires fp32 output");
        return MOREC_ERR_ARG;
    }

    TileSched p;
    p.sched_ctr = nullptr;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.num_kb = (g.K + C::BLOCK_K - 1) / C::BLOCK_K;
    p.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = (g.N + BLOCK_N - 1) / BLOCK_N;
    p.a_mn = g.a_mn; p.b_mn = g.b_mn;
    p.accumulate = g.accumulate;
    p.out_bf16 = g.out_bf16;
    p.aux_tma = 0;
    p.in_f16 = g.dtype == 3;
    p.out_f16 = g.out_f16;
    const int sms = num_sms();
    int splits = 1;
    if (g.accumulate && g.allow_split_k) {
        const int tiles = p.m_tiles * p.n_tiles;
        splits = sms / tiles;
        if (splits < 1) splits = 1;
        const int max_splits = p.num_kb / 8 > 0 ? p.num_kb / 8 : 1;   // keep >= 8 k-blocks per split
        if (splits > max_splits) splits = max_splits;
    }
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    p.splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int grid = total < sms ? total : sms;

    auto kern = gemm_kernel<KIND, BLOCK_N, Epi>;
    static bool attr_set = false;
    if (!attr_set) {
        MOREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    MOREC_CUDA(launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, stream, tmA, tmB, tmC, tmC2, p, ep));
    return MOREC_OK;
}

inline bool gemm_is_wide(int N) { return (N % 256 == 0) || N > 384; }
inline int gemm_block_n(int N, int dtype) { return dtype == 2 ? 128 : (gemm_is_wide(N) ? 256 : 128); }

template <class Epi>
int gemm_dispatch(const GemmArgs& g, const typename Epi::Params& ep, cudaStream_t stream) {
    MOREC_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
    MOREC_CHECK_ARG(g.dtype >= 0 && g.dtype <= 3, "gemm: dtype must be 0 (fp32/tf32), 1 (bf16), 2 (fp32/3xtf32) or 3 (fp16)");
    const bool wide = gemm_is_wide(g.N);
    if (g.dtype == 2) return gemm_launch<2, 128, Epi>(g, ep, stream);   // doubled stages: 128-wide tiles only
    if (g.dtype == 0) return wide ? gemm_launch<0, 256, Epi>(g, ep, stream) : gemm_launch<0, 128, Epi>(g, ep, stream);
    return wide ? gemm_launch<1, 256, Epi>(g, ep, stream) : gemm_launch<1, 128, Epi>(g, ep, stream);   // bf16 and fp16
}

}  // namespace morec
