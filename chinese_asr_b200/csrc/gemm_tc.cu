// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  in "3xTF32".
//
// Why 3xTF32: parity with the fp32 reference must be token-exact (BASELINE.json north_star), and
// bf16 / single-pass tf32 tensor-core products (8 / 11 significant bits) flip beam decisions
// (hard part 1).  Every fp32 operand x is split once into
//     x_hi = rn_tf32(x)          (cvt.rna.tf32.f32, exact in tf32)
//     x_lo = rn_tf32(x - x_hi)   (the residual, |x_lo| <= 2^-11 |x|)
// and the product is accumulated in fp32 in TMEM as  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// (the dropped a_lo*w_lo term and the rounding of the residuals are ~2^-22 relative), i.e.
// fp32-faithful products at 1/3 of the tf32 tensor rate instead of the CUDA-core fp32 rate.
//
// Kernel anatomy (one 128 x BN output tile per CTA, cta_group::1):
//   warp 0   : TMA producer - cp.async.bulk.tensor.2d of the A_hi/A_lo/W_hi/W_lo K-slabs
//              (32 fp32 = 128 B rows, SWIZZLE_128B) into a 3-stage shared-memory ring,
//              completion on `full` mbarriers (expect_tx); out-of-bounds rows / K are zero-filled
//              by TMA, so M, N and K tails need no special code.
//   warp 1   : allocates TMEM (BN fp32 columns), issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8)
//              from one elected lane: 4 K-steps x 3 products per stage, tcgen05.commit releases
//              the stage (`empty` mbarrier) and finally signals `tmem_full`.
//   warps 2-5: epilogue - tcgen05.ld 32x32b.x32 of the accumulator rows (one row per thread),
//              fused bias / temperature / LSTM-cell non-linearities, direct global stores.
// SASS evidence: UTCHMMA-class (UTC*MMA), UTMALDG, LDTM, UTCBAR in `cuobjdump -sass`.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "asr_internal.cuh"

namespace asr {

namespace tc {

constexpr int BM = 128;
constexpr int UMMA_K = 8;
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
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

// instruction descriptor, kind::tf32: D=f32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float rn_tf32_e(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

template <int BN, int BKF, int STAGES>
struct SmemLayout {
    static constexpr int kATile = BM * BKF * 4;      // 16 KB
    static constexpr int kBTile = BN * BKF * 4;
    static constexpr int kStage = 2 * kATile + 2 * kBTile;
    static constexpr int kBytes = STAGES * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ long long g_gemm_trace[1024];

template <int BN, int BKF, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                   int M, int N, int K, GemmEpilogue epi, int dbg) {
    if (epi.stop_flag && *epi.stop_flag >= 0) return;
    const bool trace = (dbg & 4) && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2;
    if (trace && threadIdx.x == 0) g_gemm_trace[0] = clock64();
    using L = SmemLayout<BN, BKF, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStage);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = (K + BKF - 1) / BKF;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (trace && kb < 200) g_gemm_trace[16 + kb * 4] = clock64();
                uint8_t* st = smem + s * L::kStage;
                if (dbg & 2) {     // experiment: load only the hi tiles
                    mbar_expect_tx(&full[s], L::kATile + L::kBTile);
                    tma_load_2d(&map_a_hi, &full[s], st, kb * BKF, m0);
                    tma_load_2d(&map_w_hi, &full[s], st + 2 * L::kATile, kb * BKF, n0);
                    continue;
                }
                mbar_expect_tx(&full[s], L::kStage);
                tma_load_2d(&map_a_hi, &full[s], st, kb * BKF, m0);
                tma_load_2d(&map_a_lo, &full[s], st + L::kATile, kb * BKF, m0);
                tma_load_2d(&map_w_hi, &full[s], st + 2 * L::kATile, kb * BKF, n0);
                tma_load_2d(&map_w_lo, &full[s], st + 2 * L::kATile + L::kBTile, kb * BKF, n0);
                if (trace && kb < 200) g_gemm_trace[16 + kb * 4 + 1] = clock64();
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (trace && kb < 200) g_gemm_trace[16 + kb * 4 + 2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint8_t* st = smem + s * L::kStage;
                const uint64_t d_ah = make_kmajor_desc<BKF>(st);
                const uint64_t d_al = make_kmajor_desc<BKF>(st + L::kATile);
                const uint64_t d_wh = make_kmajor_desc<BKF>(st + 2 * L::kATile);
                const uint64_t d_wl = make_kmajor_desc<BKF>(st + 2 * L::kATile + L::kBTile);
#pragma unroll
                for (int kk = 0; kk < BKF / UMMA_K; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);      // +32 B per K-step
                    if (dbg & 1) {   // experiment: one product per K-step
                        umma_tf32(tmem_base, d_ah + adv, d_wh + adv, idesc, (kb | kk) ? 1u : 0u);
                        continue;
                    }
                    umma_tf32(tmem_base, d_al + adv, d_wh + adv, idesc, (kb | kk) ? 1u : 0u);
                    umma_tf32(tmem_base, d_ah + adv, d_wl + adv, idesc, 1u);
                    umma_tf32(tmem_base, d_ah + adv, d_wh + adv, idesc, 1u);
                }
                umma_commit(&empty[s]);            // frees the stage when these MMAs have read it
                if (trace && kb < 200) g_gemm_trace[16 + kb * 4 + 3] = clock64();
            }
            umma_commit(tmem_full);                // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ------------------------
        // TMEM gives every thread one accumulator ROW; writing rows straight to global memory touches
        // 32 different 128-byte lines per store instruction (measured: the epilogue took as long as
        // the main loop).  Instead each warp transposes its 32 x BN block through the (now idle)
        // pipeline shared memory and stores full rows, 512 contiguous bytes per instruction.
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace && warp == 2 && lane == 0) g_gemm_trace[1] = clock64();
        constexpr int LD = BN + 4;                       // floats; keeps 16-byte alignment, conflict-free
        float* scr = reinterpret_cast<float*>(smem) + (size_t)q * 32 * LD;
        const int rbase = m0 + q * 32;
        if (epi.kind == Epi::kLstmCell) {
            constexpr int LH = BN / 4 + 4;
            float* scr_h = scr;
            float* scr_c = scr + 32 * LH;
            const int row = rbase + lane;
            const bool row_ok = row < M;
            int crow = row;
            if (row_ok && epi.c_rowidx) crow = epi.c_rowidx[row];
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                const int n = n0 + c0;
                float hv[8], cv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { hv[j] = 0.f; cv[j] = 0.f; }
                if (row_ok && n < N) {
                    // 32 columns = 8 hidden units x (i, f, g, o)
                    const float4 cp0 = *reinterpret_cast<const float4*>(epi.c_prev + (size_t)crow * epi.H + (n >> 2));
                    const float4 cp1 = *reinterpret_cast<const float4*>(epi.c_prev + (size_t)crow * epi.H + (n >> 2) + 4);
                    const float cp[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 4 * j));
                        const float gi = __uint_as_float(r[4 * j]) + b4.x;
                        const float gf = __uint_as_float(r[4 * j + 1]) + b4.y;
                        const float gg = __uint_as_float(r[4 * j + 2]) + b4.z;
                        const float go = __uint_as_float(r[4 * j + 3]) + b4.w;
                        cv[j] = sigm(gf) * cp[j] + sigm(gi) * tanhf(gg);
                        hv[j] = sigm(go) * tanhf(cv[j]);
                    }
                }
                float* ph = scr_h + lane * LH + (c0 >> 2);
                float* pc = scr_c + lane * LH + (c0 >> 2);
                *reinterpret_cast<float4*>(ph) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(ph + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                *reinterpret_cast<float4*>(pc) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                *reinterpret_cast<float4*>(pc + 4) = make_float4(cv[4], cv[5], cv[6], cv[7]);
            }
            __syncwarp();
            const int u0 = (n0 >> 2) + 2 * lane;         // BN/4 = 64 (or 32) units per tile
            if (2 * lane < BN / 4 && u0 + 1 < epi.H) {
                for (int rr = 0; rr < 32 && rbase + rr < M; ++rr) {
                    const float2 hh = *reinterpret_cast<const float2*>(scr_h + rr * LH + 2 * lane);
                    const float2 cc = *reinterpret_cast<const float2*>(scr_c + rr * LH + 2 * lane);
                    *reinterpret_cast<float2*>(epi.h_out + (size_t)(rbase + rr) * epi.H + u0) = hh;
                    *reinterpret_cast<float2*>(epi.c_out + (size_t)(rbase + rr) * epi.H + u0) = cc;
                }
            }
        } else {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                float* pr = scr + lane * LD + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(pr + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
            __syncwarp();
            const bool sc = epi.kind == Epi::kBiasScale;
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) {
                const int col = 128 * i + 4 * lane;
                const int n = n0 + col;
                if (n >= N) continue;
                float b4[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) b4[jj] = n + jj < N ? __ldg(epi.bias + n + jj) : 0.f;
                const bool full4 = n + 3 < N;
                for (int rr = 0; rr < 32 && rbase + rr < M; ++rr) {
                    float4 v = *reinterpret_cast<const float4*>(scr + rr * LD + col);
                    v.x += b4[0]; v.y += b4[1]; v.z += b4[2]; v.w += b4[3];
                    if (sc) { v.x /= epi.scale; v.y /= epi.scale; v.z /= epi.scale; v.w /= epi.scale; }
                    float* dst = epi.C + (size_t)(rbase + rr) * epi.ldc + n;
                    if (full4) {
                        *reinterpret_cast<float4*>(dst) = v;
                    } else {
                        const float vv[4] = {v.x, v.y, v.z, v.w};
                        for (int jj = 0; jj < 4; ++jj) if (n + jj < N) dst[jj] = vv[jj];
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (trace && warp == 2 && lane == 0) g_gemm_trace[2] = clock64();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over output tiles; the TMA producer and the MMA issuer run
// ahead across tile boundaries and the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so
// the epilogue of tile i (TMEM -> smem transpose -> coalesced stores) overlaps the main loop of tile
// i+1.  Measured on the one-tile-per-CTA kernel: main loop 27.4k cycles at 90 % tensor-pipe
// efficiency, but 2.6k cycles of prologue and 7.7k cycles of epilogue per tile were exposed.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int BKF, int STAGES>
struct SmemLayoutP {
    static constexpr int kATile = BM * BKF * 4;
    static constexpr int kBTile = BN * BKF * 4;
    static constexpr int kStage = 2 * kATile + 2 * kBTile;
    static constexpr int kScratch = 4 * 32 * 36 * 4;      // per epilogue warp: 32 rows x (32 + 4) floats
    static constexpr int kBytes = STAGES * kStage + kScratch + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int BKF, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tf32x3_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                              const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                              int M, int N, int K, GemmEpilogue epi, int dbg) {
    if (epi.stop_flag && *epi.stop_flag >= 0) return;
    using L = SmemLayoutP<BN, BKF, STAGES>;
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
    const int ntiles = ((M + BM - 1) / BM) * tiles_n;
    const int nkb = (K + BKF - 1) / BKF;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * L::kStage;
                    mbar_expect_tx(&full[s], L::kStage);
                    tma_load_2d(&map_a_hi, &full[s], st, kb * BKF, m0);
                    tma_load_2d(&map_a_lo, &full[s], st + L::kATile, kb * BKF, m0);
                    tma_load_2d(&map_w_hi, &full[s], st + 2 * L::kATile, kb * BKF, n0);
                    tma_load_2d(&map_w_lo, &full[s], st + 2 * L::kATile + L::kBTile, kb * BKF, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            int it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
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
                    const uint64_t d_ah = make_kmajor_desc<BKF>(st);
                    const uint64_t d_al = make_kmajor_desc<BKF>(st + L::kATile);
                    const uint64_t d_wh = make_kmajor_desc<BKF>(st + 2 * L::kATile);
                    const uint64_t d_wl = make_kmajor_desc<BKF>(st + 2 * L::kATile + L::kBTile);
#pragma unroll
                    for (int kk = 0; kk < BKF / UMMA_K; ++kk) {
                        const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);
                        umma_tf32(tacc, d_al + adv, d_wh + adv, idesc, (kb | kk) ? 1u : 0u);
                        umma_tf32(tacc, d_ah + adv, d_wl + adv, idesc, 1u);
                        umma_tf32(tacc, d_ah + adv, d_wh + adv, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        float* scr = scratch + q * (32 * 36);
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            const int acc = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
            const int rbase = m0 + q * 32;
            if (epi.kind == Epi::kLstmCell) {
                const int row = rbase + lane;
                const bool row_ok = row < M;
                int crow = row;
                if (row_ok && epi.c_rowidx) crow = epi.c_rowidx[row];
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tacc + (uint32_t)c0, r);
                    const int n = n0 + c0;
                    if (!row_ok || n >= N) continue;
                    const float4 cp0 = *reinterpret_cast<const float4*>(epi.c_prev + (size_t)crow * epi.H + (n >> 2));
                    const float4 cp1 = *reinterpret_cast<const float4*>(epi.c_prev + (size_t)crow * epi.H + (n >> 2) + 4);
                    const float cp[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
                    const float* arow = epi.addrow ? epi.addrow + (size_t)epi.addrow_idx[row] * epi.addrow_ld + n : nullptr;
                    float hv[8], cv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n + 4 * j));
                        if (arow) {
                            const float4 e4 = __ldg(reinterpret_cast<const float4*>(arow + 4 * j));
                            b4.x += e4.x; b4.y += e4.y; b4.z += e4.z; b4.w += e4.w;
                        }
                        const float gi = __uint_as_float(r[4 * j]) + b4.x;
                        const float gf = __uint_as_float(r[4 * j + 1]) + b4.y;
                        const float gg = __uint_as_float(r[4 * j + 2]) + b4.z;
                        const float go = __uint_as_float(r[4 * j + 3]) + b4.w;
                        cv[j] = sigm(gf) * cp[j] + sigm(gi) * tanhf(gg);
                        hv[j] = sigm(go) * tanhf(cv[j]);
                    }
                    float4* ho = reinterpret_cast<float4*>(epi.h_out + (size_t)row * epi.H + (n >> 2));
                    float4* co = reinterpret_cast<float4*>(epi.c_out + (size_t)row * epi.H + (n >> 2));
                    ho[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    ho[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    co[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    co[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
                    if (epi.split_hi) {
                        float hh[8], hl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { hh[j] = rn_tf32_e(hv[j]); hl[j] = rn_tf32_e(hv[j] - hh[j]); }
                        float4* sh = reinterpret_cast<float4*>(epi.split_hi + (size_t)row * epi.split_ld + (n >> 2));
                        float4* sl = reinterpret_cast<float4*>(epi.split_lo + (size_t)row * epi.split_ld + (n >> 2));
                        sh[0] = make_float4(hh[0], hh[1], hh[2], hh[3]);
                        sh[1] = make_float4(hh[4], hh[5], hh[6], hh[7]);
                        sl[0] = make_float4(hl[0], hl[1], hl[2], hl[3]);
                        sl[1] = make_float4(hl[4], hl[5], hl[6], hl[7]);
                    }
                }
            } else {
                const bool sc = epi.kind == Epi::kBiasScale;
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
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand split: x -> (rn_tf32(x), rn_tf32(x - rn_tf32(x))), with the AOperand gather fused
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__global__ void split_operand_kernel(AOperand A, int M, int K, float* __restrict__ hi, float* __restrict__ lo,
                                     const int* stop_flag) {
    if (stop_flag && *stop_flag >= 0) return;
    const int k4n = K >> 2;
    const long long total = (long long)M * k4n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / k4n);
        const int k = (int)(i - (long long)row * k4n) * 4;
        int g = 0;
        if (A.nseg > 1 && k >= A.seg[0].kend) g = 1;
        if (A.nseg > 2 && k >= A.seg[1].kend) g = 2;
        const ASeg& s = A.seg[g];
        const int kstart = g == 0 ? 0 : A.seg[g - 1].kend;
        const int r = s.rowidx ? s.rowidx[row] : row;
        const float4 v = *reinterpret_cast<const float4*>(s.base + (size_t)r * s.ld + (k - kstart));
        float4 h, l;
        h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
        h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
        h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
        h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
        *reinterpret_cast<float4*>(hi + (size_t)row * K + k) = h;
        *reinterpret_cast<float4*>(lo + (size_t)row * K + k) = l;
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

// 2-D fp32 tensor [rows, K] row-major -> box [box_rows, 32] with 128-byte swizzle
static int make_map(CUtensorMap* map, const float* base, int rows, int K, int box_rows, int BKF, int ld = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ASR_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : K) * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, BKF == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K); return ASR_ERR_CUDA; }
    return ASR_OK;
}

}  // namespace tc

int split_operand(const AOperand& A, int M, int K, float* hi, float* lo, const int* stop_flag, cudaStream_t st,
                  int64_t* launches) {
    if (M <= 0) return ASR_OK;
    if (K % 4) { set_error("split: K %% 4"); return ASR_ERR_ARG; }
    long long total = (long long)M * (K / 4);
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
    tc::split_operand_kernel<<<grid, 256, 0, st>>>(A, M, K, hi, lo, stop_flag);
    ASR_CHECK_LAUNCH();
    if (launches) ++*launches;
    return ASR_OK;
}

// C = A * W^T with pre-split operands (a_hi/a_lo [M,K], w_hi/w_lo [N,K], all dense row-major)
template <int BN, int BKF, int STAGES>
static int launch_tc_cfg(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N,
                         int K, const GemmEpilogue& epi, cudaStream_t st) {
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    ASR_TRY(tc::make_map(&ma_hi, a_hi, M, K, tc::BM, BKF, epi.lda));
    ASR_TRY(tc::make_map(&ma_lo, a_lo, M, K, tc::BM, BKF, epi.lda));
    ASR_TRY(tc::make_map(&mw_hi, w_hi, N, K, BN, BKF, epi.ldw));
    ASR_TRY(tc::make_map(&mw_lo, w_lo, N, K, BN, BKF, epi.ldw));
    static const bool persist = !(getenv("ASR_B200_GEMM_PERSIST") && atoi(getenv("ASR_B200_GEMM_PERSIST")) == 0);
    if (persist) {
        static bool attr_p = false;
        static int num_sms = 0;
        const int smem_p = tc::SmemLayoutP<BN, BKF, STAGES>::kBytes;
        if (!attr_p) {
            int dev = 0;
            ASR_CUDA(cudaGetDevice(&dev));
            ASR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
            ASR_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_persistent_kernel<BN, BKF, STAGES>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
            attr_p = true;
        }
        static const int dbg_p = getenv("ASR_B200_GEMM_DBG") ? atoi(getenv("ASR_B200_GEMM_DBG")) : 0;
        const int ntiles = ((N + BN - 1) / BN) * ((M + tc::BM - 1) / tc::BM);
        const int grid_p = ntiles < num_sms ? ntiles : num_sms;
        tc::gemm_tf32x3_persistent_kernel<BN, BKF, STAGES><<<grid_p, 192, smem_p, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, M, N, K, epi, dbg_p);
        ASR_CHECK_LAUNCH();
        return ASR_OK;
    }
    static bool attr = false;
    const int smem = tc::SmemLayout<BN, BKF, STAGES>::kBytes;
    if (!attr) {
        ASR_CUDA(cudaFuncSetAttribute(tc::gemm_tf32x3_kernel<BN, BKF, STAGES>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    static const int dbg = getenv("ASR_B200_GEMM_DBG") ? atoi(getenv("ASR_B200_GEMM_DBG")) : 0;
    dim3 grid((N + BN - 1) / BN, (M + tc::BM - 1) / tc::BM);
    tc::gemm_tf32x3_kernel<BN, BKF, STAGES><<<grid, 192, smem, st>>>(ma_hi, ma_lo, mw_hi, mw_lo, M, N, K, epi, dbg);
    ASR_CHECK_LAUNCH();
    if ((dbg & 4) && M > 100000 && K == 512) {
        static int printed = 0;
        if (printed++ < 1) {
            long long t[1024];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(t, tc::g_gemm_trace, sizeof(t));
            const int nkb = (K + BKF - 1) / BKF;
            fprintf(stderr, "[gemm trace] BN=%d BKF=%d STAGES=%d nkb=%d: start->first full %lld, epilogue %lld, total %lld cycles\n", BN, BKF,
                    STAGES, nkb, t[16 + 2] - t[0], t[2] - t[1], t[2] - t[0]);
            for (int kb = 0; kb < nkb; kb += (kb < 8 ? 1 : 4))
                fprintf(stderr, "  kb %2d: prod wait-done %6lld issued %6lld | mma full-done %6lld issued %6lld\n", kb,
                        t[16 + kb * 4] - t[0], t[16 + kb * 4 + 1] - t[0], t[16 + kb * 4 + 2] - t[0], t[16 + kb * 4 + 3] - t[0]);
        }
    }
    return ASR_OK;
}

int launch_gemm_tc(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo, int M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t st, int64_t* launches) {
    if (M <= 0) return ASR_OK;
    if ((K * 4) % 16) { set_error("gemm_tc: K*4 must be a multiple of 16"); return ASR_ERR_ARG; }
    if (epi.kind == Epi::kLstmCell && (N % 32)) { set_error("gemm_tc: LSTM epilogue needs N %% 32 == 0"); return ASR_ERR_ARG; }
    static const char* env = getenv("ASR_B200_GEMM_TILE");
    const bool wide = env ? (atoi(env) == 256) : (N >= 1024);
    if (wide) {
        // 128 x 256 tile, 64-byte K slabs, 4 stages: twice the MMA work per byte of A in flight
        ASR_TRY((launch_tc_cfg<256, 16, 4>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    } else {
        ASR_TRY((launch_tc_cfg<128, 32, 3>(a_hi, a_lo, w_hi, w_lo, M, N, K, epi, st)));
    }
    if (launches) ++*launches;
    return ASR_OK;
}

}  // namespace asr
