// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  with fp32-faithful products.
//
// Why split precision: parity with the fp32 reference must be token-exact (BASELINE.json north_star),
// and bf16 / single-pass tf32 tensor-core products (8 / 11 significant bits) flip beam decisions
// (hard part 1).  Every fp32 operand x is split once into
//     x_hi = fp16(x)             (round to nearest even: 11 significant bits like tf32, but 2 bytes and
//                                 the full-rate kind::f16 MMA; saturating, see hi_part())
//     x_lo = x - x_hi            (the residual, exact in fp32, |x_lo| <= 2^-11 |x|)
// and the product is accumulated in fp32 in TMEM as
//     a_hi*w_hi                  one tcgen05.mma.kind::f16 on fp16 operands per 16 values of k
//   + (a_lo*w + a*w_lo)          one tcgen05.mma.kind::f16 on bf16 operands (K = 16) per 8 values of k: the
//                                "cross" operand holds, per 8 values of k, [8 x bf16(a_lo) | 8 x bf16(a)] on
//                                the A side and [8 x bf16(w) | 8 x bf16(w_lo)] on the W side.
// The cross terms are 2^-11 of the product, so bf16's 2^-9 relative accuracy leaves ~2^-20; measured
// 2-5e-6 relative error against fp64, the same as three tf32 MMAs (the tensor core's truncating
// accumulation dominates).  Three MMA issue slots per 16 values of k (the tf32 hi form needed four: a
// kind::tf32 MMA covers only 8 values of k in the time a 16-bit MMA covers 16).
//
// Kernel anatomy (persistent, one CTA per SM, cta_group::1, optionally a cluster of 2 CTAs):
//   warp 0   : TMA producer - cp.async.bulk.tensor.2d of the A_cross / W_cross (128-byte rows, SWIZZLE_128B)
//              and A_hi / W_hi (64-byte rows, SWIZZLE_64B) K-slabs of 32 values of k into a 4-stage ring,
//              completion on `full` mbarriers (expect_tx); out-of-bounds rows / K are zero-filled by
//              TMA, so M, N and K tails need no special code.  In a CTA pair each CTA fetches half of
//              the W tile and multicasts it into both CTAs.
//   warp 1   : allocates TMEM (2 accumulators of BN fp32 columns), issues the MMAs (M=128, N=BN) from
//              one elected lane, tcgen05.commit releases the stage (`empty` mbarrier, of both CTAs of
//              a pair) and signals `tfull[acc]`.
//   warps 2-5: epilogue of tile i while the main loop of tile i+1 runs - tcgen05.ld 32x32b.x32 of the
//              accumulator rows, smem transpose, coalesced stores with fused bias / temperature, or
//              the fused LSTM cell (gates, E'[token] lookup, h/c and the split of h for the next GEMM).
// SASS evidence: UTCHMMA-class (UTC*MMA), UTMALDG, LDTM, UTCBAR in `cuobjdump -sass`.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// kind::f16 covers both operand formats of the engine: fp16 (hi) and bf16 (cross), chosen by the descriptor
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (one 128-byte swizzle atom along K):
// start address >> 4 | SBO (8 rows * 128 B = 1024 B) >> 4 at [32,46) | version 1 at [46,48) |
// layout SWIZZLE_128B (2) at [61,64).  LBO is unused for swizzled K-major operands.
template <int BKF>
__device__ __forceinline__ uint64_t make_kmajor_desc(const void* smem) {
    // BKF = 32: 128-byte rows, SWIZZLE_128B (2), 1024-byte atoms; BKF = 16: 64-byte rows, SWIZZLE_64B (4),
    // 512-byte atoms
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)((BKF == 32 ? 1024 : 512) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BKF == 32 ? 2 : 4) << 61;
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
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_e(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }
// ---------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over output tiles; the TMA producer and the MMA issuer run
// ahead across tile boundaries and the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so
// the epilogue of tile i (TMEM -> smem transpose -> coalesced stores) overlaps the main loop of tile
// i+1.  Measured on the one-tile-per-CTA kernel: main loop 27.4k cycles at 90 % tensor-pipe
// efficiency, but 2.6k cycles of prologue and 7.7k cycles of epilogue per tile were exposed.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA over M = 256 spans both CTAs' tensor cores; each CTA holds its
// 128 rows of A and HALF of the W tile, which the hardware shares between the two SMs.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
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
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// TMA load into this CTA's shared memory, completion bytes on the mbarrier at cluster address `bar_cluster`
// (the leader CTA's `full` barrier)
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int BN, int BKF, int STAGES, bool TWO = false>
struct SmemLayoutP {
    static_assert(BKF == 32, "K slabs of 32 values: 128-byte cross rows, 64-byte fp16 hi rows");
    static constexpr int kATile = BM * BKF * 4;                       // cross operand, 4 bytes per value of k
    static constexpr int kBTile = (TWO ? BN / 2 : BN) * BKF * 4;      // 2-SM form: each CTA keeps half of the W tile
    static constexpr int kAHi = kATile / 2;                           // fp16 hi operand, 2 bytes per value of k
    static constexpr int kBHi = kBTile / 2;
    // stage = [A_cross | W_cross | A_hi | W_hi]; every tile is a multiple of 1024 bytes (swizzle atoms)
    static constexpr int kOffAX = 0, kOffWX = kATile, kOffAH = kATile + kBTile, kOffWH = kATile + kBTile + kAHi;
    static constexpr int kStage = kATile + kBTile + kAHi + kBHi;
    static_assert(kBTile % 1024 == 0 && kBHi % 512 == 0, "tile alignment");
    static constexpr int kScratch = 4 * 32 * 36 * 4;      // per epilogue warp: 32 rows x (32 + 4) floats
    static constexpr int kBytes = STAGES * kStage + kScratch + 1024 /*align slack*/ + 256 /*barriers*/;
};

// CL = 2: the CTA pair of a cluster works on two vertically adjacent 128-row tiles of the same N tile;
// each CTA fetches HALF of the W tile and multicasts it into both CTAs' shared memory (the main loop is
// bound by L2 -> SM traffic, this removes a third of it).  A stage is reused only after the MMA
// warps of BOTH CTAs have committed it (multicast tcgen05.commit onto both `empty` barriers).
template <int BN, int BKF, int STAGES, int CL, bool TWO = false>
__global__ void __launch_bounds__(192, 1)
gemm_split_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                              const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                              int M, int N, int K, GemmEpilogue epi, int dbg) {
    if (epi.stop_flag && *epi.stop_flag >= 0) return;
    static_assert(!TWO || CL == 2, "the 2-SM form is a CTA pair");
    using L = SmemLayoutP<BN, BKF, STAGES, TWO>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* scratch = reinterpret_cast<float*>(smem + STAGES * L::kStage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage + L::kScratch);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;       // [2] accumulator ready for the epilogue
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2] accumulator drained by the 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN;
    const int mgroups = ((M + BM - 1) / BM + CL - 1) / CL;        // groups of CL vertically adjacent tiles
    const int nwork = mgroups * tiles_n;
    const int nkb = (K + BKF - 1) / BKF;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int work0 = blockIdx.x / CL, work_step = gridDim.x / CL;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TWO ? 1 : CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], TWO ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_lo) : "memory");
    }
    if (warp == 1) {
        if (TWO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(2 * BN > 256 ? 512 : 2 * BN)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(2 * BN > 256 ? 512 : 2 * BN)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CL > 1) cluster_sync_all(); else __syncthreads();       // peer barriers are initialised too
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int work = work0; work < nwork; work += work_step) {
                const int m0 = ((work / tiles_n) * CL + crank) * BM, n0 = (work % tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * L::kStage;
                    if (TWO) {
                        // both CTAs' bytes complete on the LEADER's barrier (its MMA warp consumes both halves)
                        if (crank == 0) mbar_expect_tx(&full[s], 2 * L::kStage);
                        const uint32_t fb = mapa_rank(smem_u32(&full[s]), 0);
                        const int nrow = n0 + crank * (BN / 2);
                        tma_load_2d_2sm(&map_a_lo, fb, st + L::kOffAX, kb * BKF, m0);
                        tma_load_2d_2sm(&map_w_lo, fb, st + L::kOffWX, kb * BKF, nrow);
                        tma_load_2d_2sm(&map_a_hi, fb, st + L::kOffAH, kb * BKF, m0);
                        tma_load_2d_2sm(&map_w_hi, fb, st + L::kOffWH, kb * BKF, nrow);
                        continue;
                    }
                    mbar_expect_tx(&full[s], L::kStage);
                    tma_load_2d(&map_a_lo, &full[s], st + L::kOffAX, kb * BKF, m0);
                    tma_load_2d(&map_a_hi, &full[s], st + L::kOffAH, kb * BKF, m0);
                    if (CL == 1) {
                        tma_load_2d(&map_w_lo, &full[s], st + L::kOffWX, kb * BKF, n0);
                        tma_load_2d(&map_w_hi, &full[s], st + L::kOffWH, kb * BKF, n0);
                    } else {
                        const int nrow = n0 + crank * (BN / CL);
                        tma_load_2d_mc(&map_w_lo, &full[s], st + L::kOffWX + crank * (L::kBTile / CL), kb * BKF, nrow, kMask);
                        tma_load_2d_mc(&map_w_hi, &full[s], st + L::kOffWH + crank * (L::kBHi / CL), kb * BKF, nrow, kMask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && !(TWO && crank != 0)) {               // 2-SM form: only the leader CTA issues
            constexpr uint32_t idesc_h = make_idesc_f16(TWO ? 2 * BM : BM, BN);
            constexpr uint32_t idesc_x = make_idesc_bf16(TWO ? 2 * BM : BM, BN);
            int it = 0, lt = 0;
            for (int work = work0; work < nwork; work += work_step, ++lt) {
                const int acc = lt & 1;
                const uint32_t aph = (lt >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);          // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint8_t* st = smem + s * L::kStage;
                    const uint64_t d_ax = make_kmajor_desc<32>(st + L::kOffAX);     // 128-byte rows, SWIZZLE_128B
                    const uint64_t d_wx = make_kmajor_desc<32>(st + L::kOffWX);
                    const uint64_t d_ah = make_kmajor_desc<16>(st + L::kOffAH);     // 64-byte rows, SWIZZLE_64B
                    const uint64_t d_wh = make_kmajor_desc<16>(st + L::kOffWH);
                    // every MMA consumes 32 bytes of K per row: 8 values of k of the cross operand, 16 of the hi operand
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                        if (TWO) umma2_bf16(tacc, d_ax + adv, d_wx + adv, idesc_x, (kb | kk) ? 1u : 0u);
                        else umma_bf16(tacc, d_ax + adv, d_wx + adv, idesc_x, (kb | kk) ? 1u : 0u);
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                        if (TWO) umma2_bf16(tacc, d_ah + adv, d_wh + adv, idesc_h, 1u);
                        else umma_bf16(tacc, d_ah + adv, d_wh + adv, idesc_h, 1u);
                    }
                    if (TWO) umma2_commit_mc(&empty[s], kMask);
                    else if (CL == 1) umma_commit(&empty[s]);
                    else umma_commit_mc(&empty[s], kMask);
                }
                if (TWO) umma2_commit_mc(&tfull[acc], kMask); else umma_commit(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        float* scr = scratch + q * (32 * 36);
        int lt = 0;
        for (int work = work0; work < nwork; work += work_step, ++lt) {
            const int m0 = ((work / tiles_n) * CL + crank) * BM, n0 = (work % tiles_n) * BN;
            const int acc = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            const uint32_t tacc = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
            const int rbase = m0 + q * 32;
            if (epi.kind == Epi::kLstmCell) {
                // One accumulator row per thread, 32 columns (8 hidden units x i,f,g,o) per chunk.  The
                // per-row operands of chunk c+1 (previous cell state of the source beam, the E'[token]
                // row segment) are fetched while chunk c is computed, and those of chunk 0 before the
                // accumulator is even complete: their DRAM latency used to be exposed 8 times per tile.
                const int row = rbase + lane;
                const bool row_ok = row < M;
                int crow = row;
                if (row_ok && epi.c_rowidx) crow = epi.c_rowidx[row];
                const float* cbase = epi.c_prev + (size_t)crow * epi.H;
                const float* abase = (epi.addrow && row_ok) ? epi.addrow + (size_t)epi.addrow_idx[row] * epi.addrow_ld : nullptr;
                float4 cpn[2], adn[8];
                auto fetch = [&](int n) {
                    if (row_ok && n < N) {
                        cpn[0] = *reinterpret_cast<const float4*>(cbase + (n >> 2));
                        cpn[1] = *reinterpret_cast<const float4*>(cbase + (n >> 2) + 4);
                        if (abase) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) adn[j] = __ldg(reinterpret_cast<const float4*>(abase + n + 4 * j));
                        }
                    }
                };
                fetch(n0);
                mbar_wait(&tfull[acc], aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 2
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    const int n = n0 + c0;
                    float4 cpv[2], adv4[8];
                    cpv[0] = cpn[0]; cpv[1] = cpn[1];
#pragma unroll
                    for (int j = 0; j < 8; ++j) adv4[j] = adn[j];
                    if (c0 + 32 < BN) fetch(n + 32);
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
                        cv[j] = sigm(gf) * cp[j] + sigm(gi) * tanh_e(gg);
                        hv[j] = sigm(go) * tanh_e(cv[j]);
                    }
                    float4* ho = reinterpret_cast<float4*>(epi.h_out + (size_t)row * epi.H + (n >> 2));
                    float4* co = reinterpret_cast<float4*>(epi.c_out + (size_t)row * epi.H + (n >> 2));
                    ho[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    ho[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    co[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    co[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
                    if (epi.split_hi) {
                        // the 8 hidden units of this thread are one 8-float block of the split operand
                        float hh[8], hl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { hh[j] = hi_part(hv[j]); hl[j] = hv[j] - hh[j]; }
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
                    if (n + 3 < N && !(dbg & 8)) {
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
                    } else if (n < N && !(dbg & 8)) {
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
                // 2-SM form: the leader's MMA overwrites both CTAs' accumulators, so both epilogues report to it
                if (TWO && crank != 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0));
                else mbar_arrive(&tempty[acc]);
            }
        }
    }
    if (CL > 1) cluster_sync_all(); else __syncthreads();       // no CTA leaves while its peer may still signal it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (TWO)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN > 256 ? 512 : 2 * BN) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN > 256 ? 512 : 2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand split with the AOperand gather fused (rn_tf32 only serves the legacy recurrence format)
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// LEGACY (kSplitLegacy): hi = rn_tf32(x), lo = rn_tf32(x - hi), both fp32 (the encoder recurrence packs its own
//      operands from them); otherwise hi = fp16(x) and
//      kSplitAct / kSplitWeight  "cross" operand: per 8-float block 16 bf16 values,
//          activations [bf16(x - hi) x8 | bf16(x) x8],  weights [bf16(w) x8 | bf16(w - hi) x8],
//      so one kind::f16 MMA (K = 16) over a block adds  x_lo*w + x*w_lo  for 8 values of k.
template <bool LEGACY>
__global__ void split_operand_kernel(AOperand A, int M, int K, void* __restrict__ hi_out, float* __restrict__ lo,
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
        if (LEGACY) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = rn_tf32(x[j]); l[j] = x[j] - h[j]; }
            float4* ph = reinterpret_cast<float4*>(static_cast<float*>(hi_out) + (size_t)row * K + k);
            ph[0] = make_float4(h[0], h[1], h[2], h[3]);
            ph[1] = make_float4(h[4], h[5], h[6], h[7]);
            float4* pl = reinterpret_cast<float4*>(lo + (size_t)row * K + k);
            pl[0] = make_float4(rn_tf32(l[0]), rn_tf32(l[1]), rn_tf32(l[2]), rn_tf32(l[3]));
            pl[1] = make_float4(rn_tf32(l[4]), rn_tf32(l[5]), rn_tf32(l[6]), rn_tf32(l[7]));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hi_part(x[j]); l[j] = x[j] - h[j]; }
            *reinterpret_cast<uint4*>(static_cast<hi_t*>(hi_out) + (size_t)row * K + k) =
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
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// cross operand: 2-D tensor [rows, K] of 4-byte words (two bf16 each), row-major -> box [box_rows, 32] = 128-byte
// rows with 128-byte swizzle.  hi operand: [rows, K] fp16 -> box [box_rows, 32] = 64-byte rows, 64-byte swizzle.
static int make_map_t(CUtensorMap* map, const void* base, int rows, int K, int box_rows, int ld, bool hi) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ASR_ERR_CUDA; }
    const size_t esz = hi ? sizeof(hi_t) : sizeof(float);
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : K) * esz};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, hi ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     hi ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d hi=%d", (int)r, rows, K, (int)hi); return ASR_ERR_CUDA; }
    return ASR_OK;
}
static int make_map(CUtensorMap* map, const float* base, int rows, int K, int box_rows, int, int ld = 0) {
    return make_map_t(map, base, rows, K, box_rows, ld, false);
}
static int make_map(CUtensorMap* map, const hi_t* base, int rows, int K, int box_rows, int, int ld = 0) {
    return make_map_t(map, base, rows, K, box_rows, ld, true);
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
    if (fmt == kSplitLegacy) { set_error("split: use split_operand_legacy"); return ASR_ERR_ARG; }
    long long total = (long long)M * (K / 8);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    tc::split_operand_kernel<false><<<grid, 256, 0, st>>>(A, M, K, hi, lo, stop_flag, fmt);
    ASR_CHECK_LAUNCH();
    if (launches) ++*launches;
    return ASR_OK;
}

int split_operand_legacy(const AOperand& A, int M, int K, float* hi, float* lo, cudaStream_t st) {
    if (M <= 0) return ASR_OK;
    ASR_TRY(check_split_args(A, K));
    long long total = (long long)M * (K / 8);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    tc::split_operand_kernel<true><<<grid, 256, 0, st>>>(A, M, K, hi, lo, nullptr, kSplitLegacy);
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}

// CL CTAs of a cluster work on CL vertically adjacent tiles and share the W tile by multicast
template <int BN, int BKF, int STAGES, int CL, bool TWO = false>
static int launch_tc_cluster(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const hi_t* w_hi, const float* w_lo,
                             int M, int N, int K, const GemmEpilogue& epi, cudaStream_t st, int dbg_p) {
    auto kern = tc::gemm_split_persistent_kernel<BN, BKF, STAGES, CL, TWO>;
    const int smem_p = tc::SmemLayoutP<BN, BKF, STAGES, TWO>::kBytes;
    const int tiles_m = (M + tc::BM - 1) / tc::BM, tiles_n = (N + BN - 1) / BN;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem_p;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        ASR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
        cfg.gridDim = dim3(CL * kNumSMs);
        ASR_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
        if (max_clusters < 1) { set_error("gemm_tc: no co-resident cluster of %d CTAs", CL); return ASR_ERR_CUDA; }
        if (dbg_p & 16) fprintf(stderr, "[gemm_tc] BN=%d cluster %d: %d co-resident clusters\n", BN, CL, max_clusters);
    }
    CUtensorMap mw_hi, mw_lo;       // each CTA of the cluster fetches 1/CL of the W tile rows
    ASR_TRY(tc::make_map(&mw_hi, w_hi, N, K, BN / CL, BKF, epi.ldw));
    ASR_TRY(tc::make_map(&mw_lo, w_lo, N, K, BN / CL, BKF, epi.ldw));
    const int nwork = ((tiles_m + CL - 1) / CL) * tiles_n;
    cfg.gridDim = dim3(CL * std::min(nwork, max_clusters));
    ASR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mw_hi, mw_lo, M, N, K, epi, dbg_p));
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}

// C = A * W^T with pre-split operands (a_hi/a_lo [M,K], w_hi/w_lo [N,K], all dense row-major)
template <int BN, int BKF, int STAGES>
static int launch_tc_cfg(const hi_t* a_hi, const float* a_lo, const hi_t* w_hi, const float* w_lo, int M, int N,
                         int K, const GemmEpilogue& epi, cudaStream_t st) {
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    ASR_TRY(tc::make_map(&ma_hi, a_hi, M, K, tc::BM, BKF, epi.lda));
    ASR_TRY(tc::make_map(&ma_lo, a_lo, M, K, tc::BM, BKF, epi.lda));
    ASR_TRY(tc::make_map(&mw_hi, w_hi, N, K, BN, BKF, epi.ldw));
    ASR_TRY(tc::make_map(&mw_lo, w_lo, N, K, BN, BKF, epi.ldw));
    static const int cl = getenv("ASR_B200_GEMM_CLUSTER") ? atoi(getenv("ASR_B200_GEMM_CLUSTER")) : 22;   // 22 = CTA pair with cta_group::2 MMAs
    static const int dbg_p = getenv("ASR_B200_GEMM_DBG") ? atoi(getenv("ASR_B200_GEMM_DBG")) : 0;
    const int smem_p = tc::SmemLayoutP<BN, BKF, STAGES>::kBytes;
    const int tiles_m = (M + tc::BM - 1) / tc::BM, tiles_n = (N + BN - 1) / BN;
    if (cl == 2) return launch_tc_cluster<BN, BKF, STAGES, 2>(ma_hi, ma_lo, w_hi, w_lo, M, N, K, epi, st, dbg_p);
    if (cl == 4) return launch_tc_cluster<BN, BKF, STAGES, 4>(ma_hi, ma_lo, w_hi, w_lo, M, N, K, epi, st, dbg_p);
    if (cl == 22) return launch_tc_cluster<BN, BKF, STAGES + 2, 2, true>(ma_hi, ma_lo, w_hi, w_lo, M, N, K, epi, st, dbg_p);
    {
        static bool attr_p = false;
        static int num_sms = 0;
        if (!attr_p) {
            int dev = 0;
            ASR_CUDA(cudaGetDevice(&dev));
            ASR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
            ASR_CUDA(cudaFuncSetAttribute(tc::gemm_split_persistent_kernel<BN, BKF, STAGES, 1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
            attr_p = true;
        }
        const int ntiles = tiles_n * tiles_m;
        const int grid_p = ntiles < num_sms ? ntiles : num_sms;
        tc::gemm_split_persistent_kernel<BN, BKF, STAGES, 1><<<grid_p, 192, smem_p, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, M, N, K, epi, dbg_p);
        ASR_CHECK_LAUNCH();
        return ASR_OK;
    }
}

int launch_gemm_tc(const hi_t* a_hi, const float* a_lo, const hi_t* w_hi, const float* w_lo, int M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t st, int64_t* launches) {
    if (M <= 0) return ASR_OK;
    if (K % 8) { set_error("gemm_tc: K must be a multiple of 8"); return ASR_ERR_ARG; }
    if (epi.kind == Epi::kLstmCell && (N % 32)) { set_error("gemm_tc: LSTM epilogue needs N %% 32 == 0"); return ASR_ERR_ARG; }
    static const char* env = getenv("ASR_B200_GEMM_TILE");
    // tile width: 128 for narrow outputs; otherwise 256 or 224 columns, whichever leaves the smaller
    // last wave over the 74 CTA pairs (vocabulary GEMM: 5004 = 23 x 224 -> 368 pair tiles = 4.97 waves of
    // 224 columns instead of 320 = 4.32 -> 5 waves of 256)
    int bn = N >= 1024 ? 256 : 128;
    if (bn == 256 && epi.kind != Epi::kLstmCell) {
        const int mg = ((M + tc::BM - 1) / tc::BM + 1) / 2, ncl = kNumSMs / 2;
        auto cost = [&](int w) { return (long long)((mg * ((N + w - 1) / w) + ncl - 1) / ncl) * w; };
        if (cost(224) < cost(256)) bn = 224;
    }
    static const char* env_cell = getenv("ASR_B200_CELL_TILE");
    if (env_cell && epi.kind == Epi::kLstmCell) bn = atoi(env_cell);
    if (env) bn = atoi(env);
    if (bn == 256) {
        ASR_TRY((launch_tc_cfg<256, 32, 2>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else if (bn == 224) {
        ASR_TRY((launch_tc_cfg<224, 32, 2>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else {
        ASR_TRY((launch_tc_cfg<128, 32, 3>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    }
    if (launches) ++*launches;
    return ASR_OK;
}

}  // namespace asr
