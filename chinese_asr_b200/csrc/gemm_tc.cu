// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  with fp32-faithful products.
//
// Why split precision: parity with the fp32 reference must be token-exact (BASELINE.json north_star),
// and bf16 / single-pass tf32 tensor-core products (8 / 11 significant bits) flip beam decisions
// (hard part 1).  Every fp32 operand x is split once into
//     x_hi = fp16(x)             (round to nearest even: 11 significant bits like tf32, but 2 bytes and
//                                 the full-rate kind::f16 MMA; saturating, see hi_part())
//     x_lo = x - x_hi            (the residual, exact in fp32, |x_lo| <= 2^-11 |x|)
// and the product is accumulated in fp32 in TMEM as
//     (a_lo*w + a*w_lo)          "cross" pass: one tcgen05.mma.kind::f16 on bf16 operands (K = 16) per 8
//                                values of k; the cross operand holds, per 8 values of k,
//                                [8 x bf16(a_lo) | 8 x bf16(a)] on the A side and [8 x bf16(w) | 8 x bf16(w_lo)]
//                                on the W side
//   + a_hi*w_hi                  "hi" pass: one tcgen05.mma.kind::f16 on fp16 operands per 16 values of k.
// The cross terms are 2^-11 of the product, so bf16's 2^-9 relative accuracy leaves ~2^-20.
//
// Order of accumulation.  The tensor core adds every MMA into the fp32 accumulator with TRUNCATION, so each
// MMA costs up to one ulp of the accumulator's magnitude, biased towards zero.  The first engine
// interleaved cross and hi MMAs per K slab: 3*K/16 truncations at full magnitude (192 for K = 1024, measured
// 2-5e-6 of max |C| against fp64).  Here a tile runs ALL of its cross MMAs first, while the accumulator only
// holds the 2^-11-sized cross sum (their truncations are 2^-11 smaller too), then the K/16 hi MMAs: a third of
// the full-magnitude truncations for the same operand bytes and MMA count.  Both operand kinds are 128-byte
// K-major rows with SWIZZLE_128B (cross: 32 values of k per row, hi: 64), so a stage of the ring is the same
// [A 16 KB | W half-tile] in both passes.
//
// Kernel anatomy (persistent; CTA PAIRS with tcgen05.mma.cta_group::2: one MMA of M = 256 spans both SMs'
// tensor cores, each CTA holds its 128 rows of A and HALF of the W tile, a third less operand traffic per SM
// than one CTA per tile):
//   warp 0   : TMA producer - cp.async.bulk.tensor.2d of the A / W K-slabs into a 6..8-stage ring; both CTAs'
//              bytes complete on the LEADER's `full` mbarrier (expect_tx); out-of-bounds rows / K are
//              zero-filled by TMA, so M, N and K tails need no special code.
//   warp 1   : allocates TMEM (2 accumulators of BN fp32 columns); in the leader CTA one elected lane issues
//              the MMAs, tcgen05.commit (multicast) releases the stage in both CTAs and signals `tfull[acc]`.
//   warps 2-5: epilogue of tile i while the main loop of tile i+1 runs - tcgen05.ld 32x32b.x32 of the
//              accumulator rows, then one of
//                * bias (/ temperature) -> smem transpose -> coalesced stores,
//                * the fused LSTM cell (gates, E'[token] lookup, h/c and the split of h for the next GEMM),
//                * the vocabulary epilogue (KP > 0): per (row, tile) the log-sum-exp partial (max, sum) and
//                  the top-KP logits with their token ids - the logits never reach HBM (decoder.py:133,
//                  model.py:835-836, 863-867 up to the per-utterance merge in decoder.cu).
// SASS evidence: UTCHMMA.2CTA, UTMALDG, LDTM, UTCBAR in `cuobjdump -sass` (profiles/r02_sass.txt).
#include <cuda.h>
#include <math_constants.h>

#include <algorithm>
#include <atomic>

#include "asr_internal.cuh"

namespace asr {

namespace tc {

constexpr int BM = 128;
constexpr unsigned kSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > kSpinLimit) __trap();      // turn a protocol bug into an error, not a hang
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// one lane of a converged warp (the same one every time): the guard of the single-thread TMA / MMA instructions.
// The loops around them run on all 32 lanes with warp-uniform values so that descriptors, coordinates and barrier
// addresses live in uniform registers; under `if (lane == 0)` around the whole loop ptxas could not prove that
// and wrapped every UTCHMMA / UTMALDG in an ELECT + 7 x R2UR + BRA.U.ANY loop (ncu r02c: the MMA issue loop, not
// the epilogue, paced the vocabulary GEMM).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory, completion bytes on the mbarrier at cluster address `bar_cluster`
// (the leader CTA's `full` barrier)
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
// kind::f16 covers both operand formats of the engine: fp16 (hi) and bf16 (cross), chosen by the descriptor
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (128-byte rows, 8-row / 1024-byte atoms):
// start address >> 4 | SBO 1024 >> 4 at [32,46) | version 1 at [46,48) | layout SWIZZLE_128B (2) at [61,64).
// LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t make_kmajor_desc(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 with bf16 operands: D=f32 (1<<4), A=B=BF16 (1<<7, 1<<10)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with fp16 operands: D=f32 (1<<4), A=B=F16 (0<<7, 0<<10)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// two floats -> packed bf16x2 (first argument in the low half-word), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_half, float hi_half) {
    uint32_t u;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi_half), "f"(lo_half));
    return u;
}

// The epilogue runs on 4 warps per SM: libm expf / tanhf / IEEE division are long dependent chains with
// slow-path branches that one warp per scheduler cannot overlap (measured: ~10k cycles per 32-column
// chunk).  ex2.approx / rcp.approx forms (2 ulp) are branch-free; the encoder recurrence uses the same.
// One LSTM cell (nn.LSTMCell: c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c')) - the form the encoder recurrence uses
// (encoder_tc3.cu): 5 ex2.approx.ftz + 3 rcp.approx with shared reciprocals (1/A and 1/G from one rcp(A G),
// s(o) tanh(c') = (C - 2) / (O C)), arguments clamped so that no denominator product overflows.
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void lstm_cell_fast(float gi, float gf, float gg, float go, float c_prev, float& c, float& h) {
    constexpr float kL2e = 1.4426950408889634f;
    const float A = 1.f + ex2_fast(-kL2e * fmaxf(gi, -40.f));
    const float Bf = 1.f + ex2_fast(-kL2e * fmaxf(gf, -40.f));
    const float G = ex2_fast((2.f * kL2e) * fminf(gg, 20.f)) + 1.f;
    const float O = 1.f + ex2_fast(-kL2e * fmaxf(go, -40.f));
    const float r = rcp_fast(A * G);
    const float si = r * G;                       // 1 / A
    const float tg = 1.f - 2.f * (r * A);         // 1 - 2 / G
    c = rcp_fast(Bf) * c_prev + si * tg;
    const float C = ex2_fast((2.f * kL2e) * fminf(fmaxf(c, -40.f), 20.f)) + 1.f;
    h = (C - 2.f) * rcp_fast(O * C);
}

// Sorting networks on registers (all indices compile-time): the vocabulary epilogue keeps the KP largest
// logits of a row by sorting every group of KP new values and merging it into the running list - about
// (log2 KP)^2 / 2 + log2 KP + 1 min/max pairs per value, against 2 KP for a sorted insertion, and wide enough
// (KP / 2 independent compare-exchanges per stage) to hide the min/max latency.
template <int N>
__device__ __forceinline__ void bitonic_sort_desc(float (&a)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float hi = fmaxf(a[i], a[l]), lo = fminf(a[i], a[l]);
                    const bool desc = (i & k) == 0;
                    a[i] = desc ? hi : lo;
                    a[l] = desc ? lo : hi;
                }
            }
        }
    }
}
// t (sorted descending) <- the N largest of t and b (sorted descending): max(t[i], b[N-1-i]) is a bitonic
// sequence holding exactly those, one bitonic merge sorts it
template <int N>
__device__ __forceinline__ void merge_top_desc(float (&t)[N], const float (&b)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = fmaxf(t[i], b[N - 1 - i]);
#pragma unroll
    for (int j = N >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int l = i ^ j;
            if (l > i) {
                const float hi = fmaxf(t[i], t[l]), lo = fminf(t[i], t[l]);
                t[i] = hi;
                t[l] = lo;
            }
        }
    }
}

template <int BN, int STAGES, int KP = 0, bool LSTM = false>
struct SmemLayout {
    static constexpr int kATile = BM * 128;             // 128 rows x 128 bytes (32 cross words or 64 fp16 values of k)
    static constexpr int kWTile = (BN / 2) * 128;       // this CTA's half of the W tile
    static constexpr int kStage = kATile + kWTile;
    static_assert(kWTile % 1024 == 0, "swizzle atoms");
    // KP = 0: per epilogue warp 32 rows x (32 + 4) floats (transpose for coalesced stores);
    // KP > 0: 8 epilogue warps, each its half of the bias tile (128 floats) + 32 lanes x (KP + 1) floats to hand its
    // top-KP list to the warp that shares its rows
    // the vocabulary and LSTM-cell epilogues run on 8 warps (two per TMEM lane quarter, each half of the columns):
    // their per-tile work is long enough that, with one warp per scheduler, the epilogue of a CTA's LAST tile - which
    // nothing overlaps - was as long as the tile's main loop (ncu r02c: cell GEMM 56 us for 31 us of MMAs)
    static constexpr int kEpiWarps = (KP > 0 || LSTM) ? 8 : 4;
    static constexpr int kScratch = KP > 0 ? 8 * (128 + 32 * (KP + 1)) * 4 : (LSTM ? 0 : 4 * 32 * 36 * 4);
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kBytes = STAGES * kStage + kScratch + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert((2 * STAGES + 4) * 8 + 8 <= 256, "barrier block");
    static_assert(kBytes <= 232448, "shared memory per CTA");
};

// KP = 0: bias / LSTM-cell epilogues (epi.kind); KP > 0: the vocabulary epilogue keeping the top-KP logits of
// every (row, tile).  A separate instantiation keeps its 2 x KP + 32 live registers away from the others.
template <int BN, int STAGES, int KP, bool LSTM>
__global__ void __launch_bounds__((SmemLayout<BN, STAGES, KP, LSTM>::kThreads), 1)
gemm_split_pair_kernel(const __grid_constant__ CUtensorMap map_a_x, const __grid_constant__ CUtensorMap map_a_h,
                       const __grid_constant__ CUtensorMap map_w_x, const __grid_constant__ CUtensorMap map_w_h,
                       int M, int N, int K, GemmEpilogue epi) {
    griddep_launch_dependents();
    using L = SmemLayout<BN, STAGES, KP, LSTM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* scratch = reinterpret_cast<float*>(smem + STAGES * L::kStage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage + L::kScratch);
    uint64_t* full = bars;                     // [STAGES]  (used in the leader CTA)
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2] accumulator ready for the epilogue
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2] accumulator drained by the 8 epilogue warps of the pair
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    constexpr int kTmemCols = 2 * BN > 256 ? 512 : 2 * BN;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN;
    const int mgroups = ((M + BM - 1) / BM + 1) / 2;          // pairs of vertically adjacent 128-row tiles
    const int nwork = mgroups * tiles_n;
    const int nkx = (K + 31) / 32;                            // cross slabs: 32 values of k per 128-byte row
    const int nkh = (K + 63) / 64;                            // hi slabs: 64 values of k per 128-byte row
    const int crank = (int)(blockIdx.x & 1);                  // == %cluster_ctarank: 1-D grid of 2-CTA clusters
    const int work0 = blockIdx.x >> 1, work_step = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * L::kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_h) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_h) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                        // the peer's barriers are initialised too
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // everything above touched no global memory: it overlaps the tail of the previous kernel in the stream
    griddep_wait();
    const bool stopped = epi.stop_flag && *epi.stop_flag >= 0;      // early stop of the decode loop (model.py:578, 897)

    if (stopped) {
        // nothing to do: fall through to the common exit (cluster barrier, TMEM release)
    } else if (warp == 0) {
        // all 32 lanes run the loop (uniform values), one elected lane issues the copies
        int it = 0;
        for (int work = work0; work < nwork; work += work_step) {
            const int m0 = ((work / tiles_n) * 2 + crank) * BM;
            const int nrow = (work % tiles_n) * BN + crank * (BN / 2);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const CUtensorMap* ma = pass == 0 ? &map_a_x : &map_a_h;
                const CUtensorMap* mw = pass == 0 ? &map_w_x : &map_w_h;
                const int nk = pass == 0 ? nkx : nkh, kstep = pass == 0 ? 32 : 64;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * L::kStage;
                    const uint32_t fb = mapa_rank(smem_u32(&full[s]), 0);
                    if (elect_one()) {
                        // both CTAs' bytes complete on the LEADER's barrier (its MMA warp consumes both halves)
                        if (crank == 0) mbar_expect_tx(&full[s], 2 * L::kStage);
                        tma_load_2d_2sm(ma, fb, st, kb * kstep, m0);
                        tma_load_2d_2sm(mw, fb, st + L::kATile, kb * kstep, nrow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        if (crank == 0) {                                      // only the leader CTA issues
            constexpr uint32_t idesc_h = make_idesc_f16(2 * BM, BN);
            constexpr uint32_t idesc_x = make_idesc_bf16(2 * BM, BN);
            int it = 0, lt = 0;
            for (int work = work0; work < nwork; work += work_step, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
                uint32_t first = 0u;                            // the tile's first MMA overwrites the accumulator
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int nk = pass == 0 ? nkx : nkh;
                    const uint32_t idesc = pass == 0 ? idesc_x : idesc_h;
                    for (int kb = 0; kb < nk; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full[s], (it / STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        uint8_t* st = smem + s * L::kStage;
                        const uint64_t d_a = make_kmajor_desc(st);
                        const uint64_t d_w = make_kmajor_desc(st + L::kATile);
                        if (elect_one()) {
                            // every MMA consumes 32 bytes of K per row: 8 values of k of the cross operand, 16 of hi
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                                umma2_f16(tacc, d_a + adv, d_w + adv, idesc, kk == 0 ? first : 1u);
                            }
                            umma2_commit_mc(&empty[s], 3);
                        }
                        __syncwarp();
                        first = 1u;
                    }
                }
                if (elect_one()) umma2_commit_mc(&tfull[acc], 3);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        float* scr = KP > 0 ? scratch + (warp - 2) * (128 + 32 * (KP + 1)) : scratch + q * (32 * 36);
        (void)scr;
        int lt = 0;
        for (int work = work0; work < nwork; work += work_step, ++lt) {
            const int m0 = ((work / tiles_n) * 2 + crank) * BM, n0 = (work % tiles_n) * BN;
            const int acc = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            const uint32_t tacc = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
            const int rbase = m0 + q * 32;
            if constexpr (KP > 0) {
                // ---- vocabulary epilogue: one accumulator row per thread, TWO warps per 32 rows ------------
                // Warps w and w + 4 read the same TMEM lanes (rows) and split the tile's columns (chunks of 32:
                // [0, 128) and [128, BN)).  pass 1: the KP largest logits of the warp's columns (values only,
                // sorting networks on registers); the two lists are exchanged through shared memory and merged,
                // which gives both warps the tile's KP-th largest value thr and maximum mx; pass 2 (the
                // accumulator is read again from TMEM): sum of exp against mx, and every logit above thr - plus
                // as many equal to it as still fit, in column order - goes out with its token id (the lower
                // column half fills its slots first, so ties resolve towards lower token ids).
                static_assert(BN % 32 == 0 && BN > 128 && BN <= 256 && 32 % KP == 0, "two column halves of <= 128");
                const int half = (warp - 2) >> 2;               // 0: columns [0, 128), 1: [128, BN)
                const int cbeg = half * 128, cend = half ? BN : 128;
                const int row = rbase + lane;
                const bool row_ok = row < M;
                const bool sc = epi.kind == Epi::kBiasScale;
                const int tn = work % tiles_n;
                float* s_bias = scr;                            // [128] this warp's part of the bias tile
                float* s_list = scr + 128;                      // [32][KP + 1] this warp's top-KP lists
                const float* p_list = s_list + (half ? -4 : 4) * (128 + 32 * (KP + 1));      // the partner warp's
                // columns past the vocabulary (zero-filled W rows) get -inf through the bias
                for (int c = lane; c < cend - cbeg; c += 32)
                    s_bias[c] = n0 + cbeg + c < N ? __ldg(epi.bias + n0 + cbeg + c) : -CUDART_INF_F;
                __syncwarp();
                mbar_wait(&tfull[acc], aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float t[KP];
#pragma unroll
                for (int s = 0; s < KP; ++s) t[s] = -CUDART_INF_F;
                // the 32 logits of a chunk: accumulator + bias, / temperature (model.py:834; IEEE division like
                // torch, kept out of the straight-line min/max code below)
                auto load_chunk = [&](int c0, float (&v)[32]) {
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + (c0 - cbeg));
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bb = b4[j >> 2];
                        v[j] = __uint_as_float(r[j]) + bb.x;
                        v[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
                        v[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
                        v[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
                    }
                    if (sc) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] / epi.scale;
                    }
                };
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 32) {
                    float v[32];
                    load_chunk(c0, v);
#pragma unroll
                    for (int g = 0; g < 32; g += KP) {
                        float b[KP];
#pragma unroll
                        for (int s = 0; s < KP; ++s) b[s] = v[g + s];
                        bitonic_sort_desc<KP>(b);
                        merge_top_desc<KP>(t, b);
                    }
                }
                // exchange with the warp sharing these rows (named barrier per lane quarter, 64 threads)
#pragma unroll
                for (int s = 0; s < KP; ++s) s_list[lane * (KP + 1) + s] = t[s];
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                float o[KP], own[KP];
#pragma unroll
                for (int s = 0; s < KP; ++s) { o[s] = p_list[lane * (KP + 1) + s]; own[s] = t[s]; }
                merge_top_desc<KP>(t, o);                       // t: the tile's top KP of this row
                const float thr = t[KP - 1], mx = t[0];
                int ngt = 0, gt_own = 0, eq_own = 0, gt_oth = 0, eq_oth = 0;
#pragma unroll
                for (int s = 0; s < KP; ++s) {
                    ngt += t[s] > thr ? 1 : 0;
                    gt_own += own[s] > thr ? 1 : 0;
                    eq_own += own[s] == thr ? 1 : 0;
                    gt_oth += o[s] > thr ? 1 : 0;
                    eq_oth += o[s] == thr ? 1 : 0;
                }
                // slots [0, ngt): above thr, lower column half first; [ngt, KP): equal to thr, lower half first
                int cg = half ? gt_oth : 0;
                int ce = ngt + (half ? min(eq_oth, KP - ngt) : 0);
                (void)gt_own; (void)eq_own;
                float ssum = 0.f;
                uint2* out = epi.topk_part + ((size_t)tn * M + (row_ok ? row : 0)) * KP;
                const int nvalid = N - n0;                                 // columns of this tile inside the vocabulary
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 32) {
                    float v[32];
                    load_chunk(c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) ssum += __expf(v[j] - mx);     // columns past the vocabulary: exp(-inf) = 0
                    // candidates: branch-free slot arithmetic, one predicated 8-byte store per value
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const bool in = row_ok && c0 + j < nvalid;
                        const bool gt = in && v[j] > thr;
                        const bool eq = in && v[j] == thr && ce < KP;
                        const int pos = gt ? cg : ce;
                        if (gt || eq) out[pos] = make_uint2(__float_as_uint(v[j]), (uint32_t)(n0 + c0 + j));
                        cg += gt ? 1 : 0;
                        ce += eq ? 1 : 0;
                    }
                }
                if (row_ok) epi.topk_ms[((size_t)tn * 2 + half) * M + row] = make_float2(mx, ssum);
                // the partner must have read this warp's list before the next tile overwrites it
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            } else if constexpr (LSTM) {
                // One accumulator row per thread, 32 columns (8 hidden units x i,f,g,o) per chunk; warps w and w + 4
                // share the rows of a TMEM lane quarter and take half of the tile's columns each.  The per-row
                // operands of chunk c+1 (previous cell state of the source beam, the E'[token] row segment) are
                // fetched while chunk c is computed, and those of the first chunk before the accumulator is complete.
                const int half = (warp - 2) >> 2;
                const int cbeg = half * (BN / 2);
                const int row = rbase + lane;
                const bool row_ok = row < M;
                int crow = row;
                if (row_ok && epi.c_rowidx) crow = epi.c_rowidx[row];
                const float* cbase = epi.c_prev + (size_t)crow * epi.H;
                const float* abase = (epi.addrow && row_ok) ? epi.addrow + (size_t)epi.addrow_idx[row] * epi.addrow_ld : nullptr;
                // two chunks of per-row operands in flight (ncu r02d: with one, a quarter of the epilogue's stall
                // samples sat on the E' rows - every chunk waited out an L2 / DRAM round trip)
                constexpr int NCH = BN / 64;                     // chunks of this warp's column half
                float4 cpn[2][2], adn[2][8];
                auto fetch = [&](int slot, int n) {
                    if (row_ok && n < N) {
                        cpn[slot][0] = *reinterpret_cast<const float4*>(cbase + (n >> 2));
                        cpn[slot][1] = *reinterpret_cast<const float4*>(cbase + (n >> 2) + 4);
                        if (abase) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) adn[slot][j] = __ldg(reinterpret_cast<const float4*>(abase + n + 4 * j));
                        }
                    }
                };
                fetch(0, n0 + cbeg);
                if (NCH > 1) fetch(1, n0 + cbeg + 32);
                mbar_wait(&tfull[acc], aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const int c0 = cbeg + 32 * ch;
                    const int n = n0 + c0;
                    const float4* cpv = cpn[ch & 1];
                    const float4* adv4 = adn[ch & 1];
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    if (!row_ok || n >= N) continue;
                    const float cp[8] = {cpv[0].x, cpv[0].y, cpv[0].z, cpv[0].w, cpv[1].x, cpv[1].y, cpv[1].z, cpv[1].w};
                    float hv[8], cv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 4 * j));
                        if (abase) { b4.x += adv4[j].x; b4.y += adv4[j].y; b4.z += adv4[j].z; b4.w += adv4[j].w; }
                        const float gi = __uint_as_float(r[4 * j]) + b4.x;
                        const float gf = __uint_as_float(r[4 * j + 1]) + b4.y;
                        const float gg = __uint_as_float(r[4 * j + 2]) + b4.z;
                        const float go = __uint_as_float(r[4 * j + 3]) + b4.w;
                        lstm_cell_fast(gi, gf, gg, go, cp[j], cv[j], hv[j]);
                    }
                    if (ch + 2 < NCH) fetch(ch & 1, n + 64);     // this slot's operands are consumed: refill it two chunks ahead
                    float4* ho = reinterpret_cast<float4*>(epi.h_out + (size_t)row * epi.H + (n >> 2));
                    float4* co = reinterpret_cast<float4*>(epi.c_out + (size_t)row * epi.H + (n >> 2));
                    ho[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    ho[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    co[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    co[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
                    if (epi.split_hi) {
                        // the 8 hidden units of this thread are one 8-float block of the split operand (|h| < 1)
                        float hh[8], hl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { hh[j] = __half2float(__float2half_rn(hv[j])); hl[j] = hv[j] - hh[j]; }
                        uint4* sh = reinterpret_cast<uint4*>(epi.split_hi + (size_t)row * epi.split_ld + (n >> 2));
                        uint4* sx = reinterpret_cast<uint4*>(epi.split_lo + (size_t)row * epi.split_ld + (n >> 2));
                        sh[0] = make_uint4(pack_hi2(hh[0], hh[1]), pack_hi2(hh[2], hh[3]), pack_hi2(hh[4], hh[5]),
                                           pack_hi2(hh[6], hh[7]));
                        sx[0] = make_uint4(pack_bf16x2(hl[0], hl[1]), pack_bf16x2(hl[2], hl[3]),
                                           pack_bf16x2(hl[4], hl[5]), pack_bf16x2(hl[6], hl[7]));
                        sx[1] = make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]),
                                           pack_bf16x2(hv[4], hv[5]), pack_bf16x2(hv[6], hv[7]));
                    }
                }
            } else {
                const bool sc = epi.kind == Epi::kBiasScale;
                mbar_wait(&tfull[acc], aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    float* pr = scr + lane * 36;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(pr + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                         __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    __syncwarp();
                    // 8 lanes cover one 128-byte row segment, a warp stores 4 rows per instruction;
                    // all shared-memory loads are issued before the stores (independent, unrolled)
                    const int cq = lane & 7, rq = lane >> 3;
                    const int n = n0 + c0 + 4 * cq;
                    if (n + 3 < N) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
                        float4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(scr + (4 * i + rq) * 36 + 4 * cq);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = rbase + 4 * i + rq;
                            if (row < M) {
                                float4 o = make_float4(v[i].x + b4.x, v[i].y + b4.y, v[i].z + b4.z, v[i].w + b4.w);
                                if (sc) { o.x /= epi.scale; o.y /= epi.scale; o.z /= epi.scale; o.w /= epi.scale; }
                                *reinterpret_cast<float4*>(epi.C + (size_t)row * epi.ldc + n) = o;
                            }
                        }
                    } else if (n < N) {
                        for (int i = 0; i < 8; ++i) {
                            const int row = rbase + 4 * i + rq;
                            for (int jj = 0; jj < 4 && row < M; ++jj)
                                if (n + jj < N) {
                                    float o = scr[(4 * i + rq) * 36 + 4 * cq + jj] + __ldg(epi.bias + n + jj);
                                    if (sc) o /= epi.scale;
                                    epi.C[(size_t)row * epi.ldc + n + jj] = o;
                                }
                        }
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                // the leader's MMA overwrites both CTAs' accumulators, so both epilogues report to it
                if (crank != 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0));
                else mbar_arrive(&tempty[acc]);
            }
        }
    }
    cluster_sync_all();                                        // no CTA leaves while its peer may still signal it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand split with the AOperand gather fused: hi = fp16(x) and the "cross" operand, per 8-float block 16
// bf16 values - activations [bf16(x - hi) x8 | bf16(x) x8], weights [bf16(w) x8 | bf16(w - hi) x8] - so one
// kind::f16 MMA (K = 16) over a block adds  x_lo*w + x*w_lo  for 8 values of k.
__global__ void split_operand_kernel(AOperand A, int M, int K, hi_t* __restrict__ hi_out, float* __restrict__ lo,
                                     const int* stop_flag, int fmt) {
    if (stop_flag && *stop_flag >= 0) return;
    const int k8n = K >> 3;
    const long long total = (long long)M * k8n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / k8n);
        const int k = (int)(i - (long long)row * k8n) * 8;
        int g = 0;
        if (A.nseg > 1 && k >= A.seg[0].kend) g = 1;
        if (A.nseg > 2 && k >= A.seg[1].kend) g = 2;
        const ASeg& s = A.seg[g];
        const int kstart = g == 0 ? 0 : A.seg[g - 1].kend;
        const int r = s.rowidx ? s.rowidx[row] : row;
        const float4* src = reinterpret_cast<const float4*>(s.base + (size_t)r * s.ld + (k - kstart));
        const float4 v0 = src[0], v1 = src[1];
        const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { h[j] = hi_part(x[j]); l[j] = x[j] - h[j]; }
        *reinterpret_cast<uint4*>(hi_out + (size_t)row * K + k) =
            make_uint4(pack_hi2(h[0], h[1]), pack_hi2(h[2], h[3]), pack_hi2(h[4], h[5]), pack_hi2(h[6], h[7]));
        const uint4 pl = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                    pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        const uint4 px = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                    pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
        uint4* pc = reinterpret_cast<uint4*>(lo + (size_t)row * K + k);
        pc[0] = fmt == kSplitAct ? pl : px;
        pc[1] = fmt == kSplitAct ? px : pl;
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static std::atomic<EncodeTiledFn> fn{nullptr};
    EncodeTiledFn f = fn.load(std::memory_order_acquire);
    if (!f) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            f = reinterpret_cast<EncodeTiledFn>(p);
        fn.store(f, std::memory_order_release);
    }
    return f;
}

// Both operand kinds are 2-D row-major tensors read as boxes of 128-byte rows with 128-byte swizzle:
// cross [rows, K] 4-byte words (two bf16 each) -> box [box_rows, 32]; hi [rows, K] fp16 -> box [box_rows, 64].
static int make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows, int ld, bool hi) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ASR_ERR_CUDA; }
    const size_t esz = hi ? sizeof(hi_t) : sizeof(float);
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : K) * esz};
    cuuint32_t box[2] = {hi ? 64u : 32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, hi ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d hi=%d", (int)r, rows, K, (int)hi); return ASR_ERR_CUDA; }
    return ASR_OK;
}

}  // namespace tc

static int check_split_args(const AOperand& A, int K) {
    if (K % 8) { set_error("split: K %% 8"); return ASR_ERR_ARG; }
    for (int g = 0; g + 1 < A.nseg; ++g)
        if (A.seg[g].kend % 8) { set_error("split: segment boundary %% 8"); return ASR_ERR_ARG; }
    return ASR_OK;
}

int split_operand(const AOperand& A, int M, int K, hi_t* hi, float* lo, const int* stop_flag, cudaStream_t st,
                  int64_t* launches, int fmt) {
    if (M <= 0) return ASR_OK;
    ASR_TRY(check_split_args(A, K));
    long long total = (long long)M * (K / 8);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    tc::split_operand_kernel<<<grid, 256, 0, st>>>(A, M, K, hi, lo, stop_flag, fmt);
    ASR_CHECK_LAUNCH();
    if (launches) ++*launches;
    return ASR_OK;
}

template <int BN, int STAGES, int KP, bool LSTM = false>
static int launch_pair(const hi_t* a_hi, const float* a_lo, const hi_t* w_hi, const float* w_lo, int M, int N, int K,
                       const GemmEpilogue& epi, cudaStream_t st) {
    auto kern = tc::gemm_split_pair_kernel<BN, STAGES, KP, LSTM>;
    constexpr int smem = tc::SmemLayout<BN, STAGES, KP, LSTM>::kBytes;
    CUtensorMap ma_x, ma_h, mw_x, mw_h;          // each CTA of the pair fetches its 128 rows of A, half of the W tile rows
    ASR_TRY(tc::make_map(&ma_x, a_lo, M, K, tc::BM, epi.lda, false));
    ASR_TRY(tc::make_map(&ma_h, a_hi, M, K, tc::BM, epi.lda, true));
    ASR_TRY(tc::make_map(&mw_x, w_lo, N, K, BN / 2, epi.ldw, false));
    ASR_TRY(tc::make_map(&mw_h, w_hi, N, K, BN / 2, epi.ldw, true));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // decoder-step launches: see griddep_wait()
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(tc::SmemLayout<BN, STAGES, KP, LSTM>::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    // function attributes and cluster occupancy are per device: one slot per device id, set on first use there
    static std::atomic<int> max_clusters[64];
    int dev = 0;
    ASR_CUDA(cudaGetDevice(&dev));
    int mc = max_clusters[dev & 63].load(std::memory_order_acquire);
    if (!mc) {
        ASR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cfg.gridDim = dim3(2 * kNumSMs);
        ASR_CUDA(cudaOccupancyMaxActiveClusters(&mc, kern, &cfg));
        if (mc < 1) { set_error("gemm_tc: no co-resident CTA pair (BN = %d)", BN); return ASR_ERR_CUDA; }
        max_clusters[dev & 63].store(mc, std::memory_order_release);
    }
    const int tiles_m = (M + tc::BM - 1) / tc::BM, tiles_n = (N + BN - 1) / BN;
    const int nwork = ((tiles_m + 1) / 2) * tiles_n;
    cfg.gridDim = dim3(2 * std::min(nwork, mc));
    cfg.numAttrs = epi.pdl ? 2 : 1;
    ASR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma_x, ma_h, mw_x, mw_h, M, N, K, epi));
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}

int vocab_topk_slots(int k) { return k <= 1 ? 2 : (k <= 4 ? 8 : (k <= 8 ? 16 : 32)); }

// C = A * W^T with pre-split operands (a_hi/a_lo [M,K], w_hi/w_lo [N,K], row strides epi.lda / epi.ldw or K)
int launch_gemm_tc(const hi_t* a_hi, const float* a_lo, const hi_t* w_hi, const float* w_lo, int M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t st, int64_t* launches) {
    if (M <= 0) return ASR_OK;
    if (K % 8) { set_error("gemm_tc: K must be a multiple of 8"); return ASR_ERR_ARG; }
    if (epi.kind == Epi::kLstmCell && (N % 32)) { set_error("gemm_tc: LSTM epilogue needs N %% 32 == 0"); return ASR_ERR_ARG; }
    if (epi.topk_slots) {
        // vocabulary projection with the fused log-sum-exp / top-k partials: 224-wide tiles (kVocabTiles = 23)
        if (N != kVocab || !epi.topk_part || !epi.topk_ms || epi.kind == Epi::kLstmCell) { set_error("gemm_tc: bad top-k epilogue"); return ASR_ERR_ARG; }
        switch (epi.topk_slots) {
            case 2: ASR_TRY((launch_pair<kVocabTileN, 6, 2>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st))); break;
            case 8: ASR_TRY((launch_pair<kVocabTileN, 6, 8>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st))); break;
            case 16: ASR_TRY((launch_pair<kVocabTileN, 6, 16>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st))); break;
            case 32: ASR_TRY((launch_pair<kVocabTileN, 6, 32>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st))); break;
            default: set_error("gemm_tc: top-k slots %d", epi.topk_slots); return ASR_ERR_ARG;
        }
        if (launches) ++*launches;
        return ASR_OK;
    }
    // tile width: 128 for narrow outputs; otherwise 256 or 224 columns, whichever leaves the smaller
    // last wave over the 74 CTA pairs
    int bn = N >= 1024 ? 256 : 128;
    if (bn == 256 && epi.kind != Epi::kLstmCell) {
        const int mg = ((M + tc::BM - 1) / tc::BM + 1) / 2, ncl = kNumSMs / 2;
        auto cost = [&](int w) { return (long long)((mg * ((N + w - 1) / w) + ncl - 1) / ncl) * w; };
        if (cost(224) < cost(256)) bn = 224;
    }
    if (epi.kind == Epi::kLstmCell) {
        ASR_TRY((launch_pair<256, 6, 0, true>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else if (bn == 256) {
        ASR_TRY((launch_pair<256, 6, 0>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else if (bn == 224) {
        ASR_TRY((launch_pair<224, 6, 0>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else if (N <= 128 && ((M + 2 * tc::BM - 1) / (2 * tc::BM)) * 4 <= kNumSMs / 2) {
        // a narrow output over few rows (the decoder's query projection: 4096 x 128): 32-column tiles put 4x as many
        // CTA pairs to work - the GEMM is a latency chain (prologue, pipeline fill, epilogue), not a throughput problem
        ASR_TRY((launch_pair<32, 8, 0>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else {
        ASR_TRY((launch_pair<128, 8, 0>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    }
    if (launches) ++*launches;
    return ASR_OK;
}

}  // namespace asr
