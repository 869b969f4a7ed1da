// Tensor-core LSTM recurrence, weights fully resident in TENSOR MEMORY.
//
// The whole recurrent weight slice of a CTA (128 gate rows x 256 k) lives in TMEM, split on the fly at kernel
// start (gemm_tc.cu's fp16 hi part; the residual is carried as a second fp16, scaled by 2^11 into the range of
// the value itself):
//     columns [  0,128) : W_hi  = fp16(w), 2 per column                       128 lanes x 256 fp16
//     columns [128,256) : W_lo' = fp16((w - W_hi) * 2^11)
//     columns [256,256 + 2 NB) : fp32 accumulators D_hi[128 gate rows, NB sequences] | D_lo[128, NB]
// Shared memory holds the h operand tile: per 32-wide K range a slab of 2 NB rows x 64 bytes (SWIZZLE_64B,
// K-major), rows [0, NB) = fp16(h_hi), rows [NB, 2 NB) = fp16(h_lo * 2^11).  Per step 32 kind::f16 MMAs (two per
// 16 values of k):
//     [D_hi | D_lo] += W_hi  (TMEM) * [h_hi ; h_lo']^T      N = 2 NB  : w_hi h_hi  and  w_hi h_lo'
//          D_lo     += W_lo' (TMEM) * h_hi^T                N = NB    : w_lo' h_hi
// and the gate warps read  D_hi + 2^-11 D_lo.  An MMA with A in TMEM is paced by the fetch of its 128 lanes x
// 32 bytes of A (~74 cycles whatever N is), so sharing one fetch of W_hi between the h_hi and h_lo' products
// costs 32 A fetches per step where the [hi | bf16 cross] operand form (round 1 / early round 2: 16 fp16 + 32
// bf16 MMAs over K = 8) paid 48; the exchanged image of h shrinks from 6 to 4 bytes per value, and the residual
// products are rounded to 2^-11 relative (fp16) instead of 2^-9 (bf16): 2^-23 of the product.
// Exchange of h_{t+1}: each CTA writes the image of its 32-unit slab (already in the swizzled UMMA layout) to
// global staging and multicasts it with one cp.async.bulk into the tiles of all 8 CTAs; mbarriers only, no
// cluster barrier per step.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "asr_internal.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace asr {

using namespace tcx;

namespace rec3 {

__device__ __forceinline__ float rcp_f(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// One LSTM cell (nn.LSTM: c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c')) on ex2.approx / rcp.approx.  The gate
// phase of a recurrence step is paced by the MUFU pipe (16 lanes per clock), so reciprocals are shared:
// 1/A and 1/G from one rcp(A G), and s(o) tanh(c') = (C - 2) / (O C) with C = e^(2c') + 1 from one rcp(O C) -
// 5 ex2 + 3 rcp per cell instead of 5 + 5.  Arguments are clamped to +-40 (s, tanh are saturated to fp32 there) so
// that no product of two denominators overflows.  Relative error ~3 ulp per factor, like the unshared form.
__device__ __forceinline__ float ex2_f(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void lstm_cell(float gi, float gf, float gg, float go, float c_prev, float& c, float& h) {
    // e^x = ex2(x log2 e) without __expf's denormal-range rescaling (a compare and two predicated multiplies per
    // call): every argument is clamped, so no result is denormal or infinite
    constexpr float kL2e = 1.4426950408889634f;
    const float A = 1.f + ex2_f(-kL2e * fmaxf(gi, -40.f));
    const float Bf = 1.f + ex2_f(-kL2e * fmaxf(gf, -40.f));
    const float G = ex2_f((2.f * kL2e) * fminf(gg, 20.f)) + 1.f;
    const float O = 1.f + ex2_f(-kL2e * fmaxf(go, -40.f));
    const float r = rcp_f(A * G);
    const float si = r * G;                       // 1 / A
    const float tg = 1.f - 2.f * (r * A);         // 1 - 2 / G
    c = rcp_f(Bf) * c_prev + si * tg;
    const float C = ex2_f((2.f * kL2e) * fminf(fmaxf(c, -40.f), 20.f)) + 1.f;
    h = (C - 2.f) * rcp_f(O * C);
}

// one lane of a converged warp: the guard of the single-thread MMA / commit instructions (see gemm_tc.cu)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                                   uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// D[tmem] += A[tmem] * B[smem desc], kind::f16 (operand formats from the instruction descriptor)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major SWIZZLE_64B descriptor (rows of 64 bytes, 8-row / 512-byte atoms)
__device__ __forceinline__ uint64_t kmajor_sw64_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// byte offset of 16-bit element (row, kk in 0..31) inside a [rows x 32] SWIZZLE_64B K-major slab
__device__ __forceinline__ uint32_t sw64_offset(int row, int kk) {
    return (uint32_t)(row * 64 + ((((kk >> 3) ^ ((row >> 1) & 3)) << 4) | ((kk & 7) << 1)));
}
// kind::f16 with fp16 operands (format 0)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// tcgen05.ld of N consecutive accumulator columns (one TMEM lane per thread) WITHOUT the wait: several loads are
// issued back to back and tmem_wait_ld() follows once.  The registers are defined only after that wait.
__device__ __forceinline__ void tmem_ld4_nw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols_nowait(uint32_t taddr, uint32_t* r) {
    static_assert(N % 4 == 0 && N >= 4 && N <= 32, "columns per warp");
    if constexpr (N >= 16) {
        tmem_ld16_nw(taddr, r);
        if constexpr (N > 16) tmem_ld_cols_nowait<N - 16>(taddr + 16, r + 16);
    } else if constexpr (N >= 8) {
        tmem_ld8_nw(taddr, r);
        if constexpr (N > 8) tmem_ld_cols_nowait<N - 8>(taddr + 8, r + 8);
    } else {
        tmem_ld4_nw(taddr, r);
    }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 gate warps; the last one also issues the step's MMAs (a few dozen uniform instructions once h has arrived) before
// it turns to its share of the gate work.  A 17th, issue-only warp (round 1 / 2) put 5 warps on one scheduler and capped
// every thread at 96 registers: the gate warps spilled, and their local-memory reloads missed L1 behind the streamed xg
// rows (ncu r02d: 12 % of all stall samples on four LDL).  16 warps get 128 registers.
constexpr int kThreads = 512;
constexpr int kIssueWarp = 15;

struct Params {
    const float* xg;            // [rows, 2048] permuted gate pre-activations (bias included): column dir*1024 + j*128 + 4u + gate
    const float* whh;           // [2, 1024, 256] fp32 W_hh, rows permuted like the xg columns
    const float* x_in;
    float* y_packed;
    float* y_utt;
    hi_t* y_hi;             // optional [rows, 512]: fp16(y), rows as y_utt when that is written, else as y_packed
    uint32_t* y_cross;      // optional: per 8-float block 8 x bf16(y - hi) then 8 x bf16(y)  (gemm_tc.cu kSplitAct)
    float* h_fin;
    float* c_fin;
    uint8_t* stage;             // [gridDim.x][kStageBytes] global staging images
    const int* len_sorted;
    const int* toff;
    const int* uoff;
    int B;
    int rows_per_chunk;
    int nchunks;
    long long* dbg;
};

constexpr int kMaxNB = 128;              // sequences per cluster: the accumulator takes TMEM columns [256, 256 + 2 NB)
constexpr int kStageBytes = kMaxNB * 128;    // per-CTA staging slot (largest NB)
constexpr float kLoScale = 2048.f;       // residuals are carried as fp16(lo * 2^11): |lo * 2^11| <= |x|, so they share x's range
constexpr float kLoUnscale = 1.f / 2048.f;

// HAS_RES: the layer has a residual input (layers 1-3, util.py:1284-1291).  LAST: the output goes out utterance-major
// (the encoder memory `enc` and the keys GEMM's split operand); otherwise packed time-major for the next layer.
template <int NB, bool HAS_RES, bool LAST>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kThreads, 1)
lstm_rec_tc3_kernel(Params p) {
    // 16 warps: warp = (TMEM lane quarter q = warp & 3, column group warp >> 2).  TMEM lane m of the CTA's
    // 128 gate rows is (unit m >> 2, gate m & 3): the four gates of a unit sit in 4 adjacent lanes, so a
    // 4x4 transpose by warp shuffles gives every lane complete (i, f, g, o) pre-activations of P cells
    // (unit, batch row) - no shared-memory transposition and no block barrier in the gate phase.
    constexpr int CW = NB / 4;               // accumulator columns (batch rows) per warp
    constexpr int P = CW / 4;                // cells per lane: rows cg * CW + 4 b + (lane & 3)
    constexpr int kSlab = NB * 128;          // one 32-wide K range: [NB rows fp16(h_hi) | NB rows fp16(h_lo')] x 64 bytes
    constexpr int kTrStride = 40;            // floats per transposition row: 32 lanes + 8 (conflict-free 16-byte reads)
    static_assert(NB % 16 == 0 && NB <= kMaxNB, "NB");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* T = smem;                                             // 8 slabs x kSlab
    float* s_tr = reinterpret_cast<float*>(T + 8 * kSlab);         // [16 gate warps][CW][kTrStride] gate transposition
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(s_tr + 16 * CW * kTrStride);
    uint64_t* h_ready = mma_done + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
    int* s_len = reinterpret_cast<int*>(tmem_slot + 1);            // [NB]

    cg::cluster_group cluster = cg::this_cluster();
    const int j = (int)cluster.block_rank();
    const int cid = blockIdx.x / 8;
    const int dir = cid / p.nchunks;
    const int chunk = cid - dir * p.nchunks;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler (uniform registers)
    const int r0 = chunk * p.rows_per_chunk;
    const int nrows = min(p.rows_per_chunk, p.B - r0);
    uint8_t* stage = p.stage + (size_t)blockIdx.x * kStageBytes;

    // ---- one-time setup ------------------------------------------------------------------------
    for (int i = tid; i < (8 * kSlab) / 16; i += kThreads) reinterpret_cast<float4*>(T)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < kSlab / 16; i += kThreads) reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < NB) s_len[tid] = tid < nrows ? p.len_sorted[r0 + tid] : 0;
    if (tid == 0) {
        mbar_init(mma_done, 8);
        mbar_init(h_ready, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_alo = tmem_base + 128;
    const uint32_t tmem_d = tmem_base + 256;
    if (warp < 4) {
        // TMEM lane m <- row dir*1024 + j*128 + m of the permuted W_hh, split on the fly: columns [0, 128) fp16(w)
        // pairs, columns [128, 256) fp16((w - fp16(w)) * 2^11) pairs
        const int m = 32 * warp + lane;
        const float4* wrow = reinterpret_cast<const float4*>(p.whh + ((size_t)dir * kGates + j * 128 + m) * kEncH);
#pragma unroll 1
        for (int c0 = 0; c0 < kEncH; c0 += 32) {           // c0 < 128: hi columns, else the residual columns
            const bool lo = c0 >= kEncH / 2;
            const int f0 = lo ? c0 - kEncH / 2 : c0;        // first fp16 pair column = float 2 f0
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 v = __ldg(wrow + f0 / 2 + q);
                const float h0 = hi_part(v.x), h1 = hi_part(v.y), h2 = hi_part(v.z), h3 = hi_part(v.w);
                r[2 * q] = lo ? pack_hi2((v.x - h0) * kLoScale, (v.y - h1) * kLoScale) : pack_hi2(h0, h1);
                r[2 * q + 1] = lo ? pack_hi2((v.z - h2) * kLoScale, (v.w - h3) * kLoScale) : pack_hi2(h2, h3);
            }
            tmem_st32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
        }
        tmem_wait_st();
    }
    __threadfence();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int Lc = s_len[0];

    const int uu = 8 * (warp & 3) + (lane >> 2);      // unit of this lane within the CTA's 32
    const int jr = lane & 3;
    const int col0 = (warp >> 2) * CW + jr;           // batch row of cell b: col0 + 4 b
    const int ocol = dir * kEncH + 32 * j + uu;
    float c_reg[P], h_reg[P];
    // byte offset of (row col0 + 4 q, unit) in the fp16 hi rows of the slab image: the swizzle phase repeats every
    // 8 rows, so two bases (even / odd q) + 256 q
    const uint32_t img_e = sw64_offset(col0, uu), img_o = sw64_offset(col0 + 4, uu) - 256u;
    int uo[LAST ? P : 1];                             // LAST: first utterance-major row of the cell's sequence
#pragma unroll
    for (int q = 0; q < P; ++q) {
        c_reg[q] = 0.f;
        h_reg[q] = 0.f;
        if (LAST) uo[q] = col0 + 4 * q < nrows ? p.uoff[r0 + col0 + 4 * q] : 0;
    }
    // this warp's transposition rows (gate warps only), as a shared-window address: the 1024-byte alignment of the
    // dynamic shared memory hides the address space from the compiler, and generic LD / ST are slower than LDS / STS
    const uint32_t tr = smem_u32(s_tr) + (uint32_t)(warp * (CW * kTrStride) * 4);

    cluster.sync();

    const uint32_t t_base = smem_u32(T);
    constexpr uint32_t idesc_2n = idesc_f16(128, 2 * NB);
    constexpr uint32_t idesc_n = idesc_f16(128, NB);

    // rows are sorted by decreasing length: the active count follows t incrementally
    int nact = dir == 0 ? nrows : 0;
    int prev_t = 0, prev_row_t = 0, prev_nact = 0;
    float yv[P];
#pragma unroll
    for (int q = 0; q < P; ++q) yv[q] = 0.f;
    // word of this unit's bf16 pair inside its 8-float block of the cross operand, relative to the element index:
    // even units pack the two residuals (words 0-3), odd units the two values (words 4-7)
    const bool odd = (ocol & 1) != 0;
    const int cross_adj = (odd ? 4 : 0) + ((ocol & 7) >> 1) - (ocol & 7);
    // layer output of one step: y = h + residual input (util.py:1284-1291) and its operand split for the next
    // GEMM.  Called one step late (after the next step's MMAs are issued) so that it never delays them.  Element
    // indices fit 32 bits (rows x 512 < 2^32 for any batch the workspace holds).
    auto store_outputs = [&](int t, int row_t, int n_act) {
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = col0 + 4 * q;
            const float y = yv[q];
            // the lane 4 away holds the neighbouring unit of the same row
            const float hi = __half2float(__float2half_rn(y));          // |y| <= 1 + |x|: far inside fp16's range
            const float lo = y - hi;
            const float lo_n = __shfl_xor_sync(0xffffffffu, lo, 4);
            const float y_n = __shfl_xor_sync(0xffffffffu, y, 4);
            if (i < n_act) {
                const uint32_t e = (uint32_t)(row_t + i) * kEnc + ocol;
                uint32_t se = e;                                         // element index in the split operand
                if (LAST) {
                    se = (uint32_t)(uo[q] + t) * kEnc + ocol;            // rows as `enc`
                    p.y_utt[se] = y;
                    if (p.y_packed) p.y_packed[e] = y;
                } else {
                    p.y_packed[e] = y;
                }
                p.y_hi[se] = __float2half_rn(hi);
                const __nv_bfloat162 pk = odd ? __floats2bfloat162_rn(y_n, y) : __floats2bfloat162_rn(lo, lo_n);
                p.y_cross[(int)se + cross_adj] = *reinterpret_cast<const uint32_t*>(&pk);
            }
        }
    };

    int toff_t = Lc > 0 ? p.toff[dir == 0 ? 0 : Lc - 1] : 0;
    for (int s = 0; s < Lc; ++s) {
        const int t = dir == 0 ? s : Lc - 1 - s;
        if (dir == 0) { while (nact > 0 && s_len[nact - 1] <= t) --nact; }
        else { while (nact < nrows && s_len[nact] > t) ++nact; }
        const int row_t = toff_t + r0;
        if (s + 1 < Lc) toff_t = p.toff[dir == 0 ? s + 1 : Lc - 2 - s];      // next step's row base, a step ahead

        {
        // this step's loads first (they do not depend on h: issued while the exchange is still in flight), then the
        // issue warp starts the MMAs, then everybody stores the previous step's outputs under them
        float4 xg4[P];
        float xres[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = col0 + 4 * q;
            xg4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            xres[q] = 0.f;
            if (i < nact) {
                const uint32_t r = (uint32_t)(row_t + i);
                if (HAS_RES) xres[q] = __ldg(p.x_in + r * kEnc + ocol);
                xg4[q] = __ldg(reinterpret_cast<const float4*>(p.xg + (size_t)r * (2 * kGates) + dir * kGates + j * 128) + uu);
            }
        }

        if (warp == kIssueWarp) {
            // all 32 lanes wait (uniform control flow keeps descriptors and TMEM addresses in uniform registers),
            // one elected lane issues
            if (s > 0) mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
            tc_fence_after();
            if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[s * 8 + 0] = clock64();
            const uint64_t d0 = kmajor_sw64_desc(t_base);
            if (elect_one()) {
                // Per 16 values of k two MMAs, each reading 8 TMEM columns of A: W_hi against both row blocks of the
                // tile at once (N = 2 NB: accumulator columns [0, NB) += w_hi h_hi, [NB, 2 NB) += w_hi h_lo'), and
                // W_lo' against the h_hi rows into [NB, 2 NB).  The full-magnitude sum sees 16 truncating
                // accumulations, the 2^-11-sized cross sum its own 32.
#pragma unroll
                for (int kh = 0; kh < 16; ++kh) {
                    const uint64_t db = d0 + (uint64_t)(((kh >> 1) * kSlab + (kh & 1) * 32) >> 4);
                    umma_bf16_ts(tmem_d, tmem_base + (uint32_t)(8 * kh), db, idesc_2n, kh > 0 ? 1u : 0u);
                    umma_bf16_ts(tmem_d + (uint32_t)NB, tmem_alo + (uint32_t)(8 * kh), db, idesc_n, 1u);
                }
                umma_commit_mc(mma_done, (uint16_t)0xFF);
            }
            __syncwarp();
            if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[s * 8 + 1] = clock64();
        }
        if (s > 0) store_outputs(prev_t, prev_row_t, prev_nact);

        mbar_wait(mma_done, (uint32_t)(s & 1));
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 2] = clock64();
        {
            // D_hi + 2^-11 D_lo, read in two column halves (both accumulators of a half in flight, one wait) to
            // stay inside the 96 registers 17 warps leave per thread
            constexpr int CA = 4 * ((P + 1) / 2), CB = CW - CA;
            const uint32_t dcol = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * CW);
            float d[CW];
            {
                uint32_t dh[CA], dl[CA];
                tmem_ld_cols_nowait<CA>(dcol, dh);
                tmem_ld_cols_nowait<CA>(dcol + (uint32_t)NB, dl);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < CA; ++c) {
                    asm volatile("" : "+r"(dh[c]), "+r"(dl[c]));      // defined only after the wait above
                    d[c] = fmaf(__uint_as_float(dl[c]), kLoUnscale, __uint_as_float(dh[c]));
                }
            }
            if constexpr (CB > 0) {
                uint32_t dh[CB], dl[CB];
                tmem_ld_cols_nowait<CB>(dcol + (uint32_t)CA, dh);
                tmem_ld_cols_nowait<CB>(dcol + (uint32_t)(NB + CA), dl);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    asm volatile("" : "+r"(dh[c]), "+r"(dl[c]));
                    d[CA + c] = fmaf(__uint_as_float(dl[c]), kLoUnscale, __uint_as_float(dh[c]));
                }
            }
            tc_fence_before();
            // 4 x 4 transposition through this warp's shared-memory rows: lane L (unit L >> 2, gate L & 3) writes its
            // gate of row c to tr[c][L]; the lane of cell (unit, row 4 q + jr) reads gates i, f, g, o as one float4
            // (20 STS + 5 LDS.128 per lane and step where warp shuffles took 30 SHFL + 60 FSEL)
#pragma unroll
            for (int c = 0; c < CW; ++c)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(tr + (uint32_t)((c * kTrStride + lane) * 4)), "f"(d[c]) : "memory");
            __syncwarp();
#pragma unroll
            for (int q = 0; q < P; ++q) {
                float di, df, dg, dO;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(di), "=f"(df), "=f"(dg), "=f"(dO)
                             : "r"(tr + (uint32_t)(((4 * q + jr) * kTrStride + 4 * (lane >> 2)) * 4)));
                const int i = col0 + 4 * q;
                if (i < nact) {
                    float c, hh;
                    lstm_cell(xg4[q].x + di, xg4[q].y + df, xg4[q].z + dg, xg4[q].w + dO, c_reg[q], c, hh);
                    c_reg[q] = c;
                    h_reg[q] = hh;
                    const float hi = __half2float(__float2half_rn(hh));      // |h| < 1
                    // image of this CTA's 32-unit slab of the next B operand, in global staging: fp16(h_hi) in
                    // row i, fp16(h_lo * 2^11) in row NB + i (same swizzle phase: NB is a multiple of 8)
                    uint8_t* img = stage + ((q & 1) ? img_o : img_e) + 256 * q;
                    *reinterpret_cast<__half*>(img) = __float2half_rn(hi);
                    *reinterpret_cast<__half*>(img + NB * 64) = __float2half_rn((hh - hi) * kLoScale);
                    yv[q] = hh + xres[q];
                }
            }
        }
        }
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 4] = clock64();
        if (s + 1 < Lc) {
            // the image stores (generic proxy) must be performed before the bulk copy (async proxy) of thread 0 reads
            // them: every writer fences generic -> async for global memory, then the CTA barrier orders the copy after
            // all of them.  (A __threadfence() in front of it, as up to round 2, costs a second MEMBAR.GPU per step.)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncthreads();
            if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 5] = clock64();
            if (tid == 0) {
                mbar_expect_tx(h_ready, 8u * (uint32_t)kSlab);
                bulk_g2s_multicast(T + j * kSlab, stage, (uint32_t)kSlab, h_ready, (uint16_t)0xFF);
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 6] = clock64();
            }
        }
        prev_t = t; prev_row_t = row_t; prev_nact = nact;
    }
    if (Lc > 0) store_outputs(prev_t, prev_row_t, prev_nact);

#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int i = col0 + 4 * q;
        if (i < nrows) {
            p.h_fin[(size_t)(r0 + i) * kEnc + ocol] = h_reg[q];
            p.c_fin[(size_t)(r0 + i) * kEnc + ocol] = c_reg[q];
        }
    }
    tc_fence_before();
    cluster.sync();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int NB, bool HAS_RES, bool LAST>
static int launch_mode(const Params& p, cudaStream_t st) {
    const size_t smem = 8 * (size_t)(NB * 128) + 1024 + 64 + NB * 4 + 16 * (size_t)(NB / 4) * 40 * 4;
    ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&lstm_rec_tc3_kernel<NB, HAS_RES, LAST>), smem));
    lstm_rec_tc3_kernel<NB, HAS_RES, LAST><<<2 * p.nchunks * 8, kThreads, smem, st>>>(p);
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}
template <int NB>
static int launch(const Params& p, cudaStream_t st) {
    if (p.y_utt) return launch_mode<NB, true, true>(p, st);
    return p.x_in ? launch_mode<NB, true, false>(p, st) : launch_mode<NB, false, false>(p, st);
}

}  // namespace rec3

size_t rec3_stage_bytes_per_cta() { return rec3::kStageBytes; }

int launch_lstm_recurrence_tc3(asr_handle* h, int layer, const float* xg, const float* x_in, float* y_packed,
                               float* y_utt, float* h_fin, float* c_fin, cudaStream_t st, hi_t* split_hi,
                               float* split_lo) {
    const BatchMeta& m = h->meta;
    rec3::Params p{};
    p.y_hi = split_hi;
    p.y_cross = reinterpret_cast<uint32_t*>(split_lo);
    p.xg = xg;
    p.whh = h->w.enc_w_hh[layer];
    p.x_in = x_in;
    p.y_packed = y_packed;
    p.y_utt = y_utt;
    p.h_fin = h_fin;
    p.c_fin = c_fin;
    p.stage = reinterpret_cast<uint8_t*>(h->ws.rec_stage);
    p.len_sorted = m.d_len_sorted;
    p.toff = m.d_toff;
    p.uoff = m.d_uoff_sorted;
    p.B = m.B;
    if (!split_hi || !split_lo || !y_packed || (y_utt && !x_in)) { set_error("recurrence: missing output buffer"); return ASR_ERR_ARG; }
    // <= 15 clusters of 8 CTAs are co-resident on a B200 (measured): aim at one round, i.e. at most
    // 7 chunks per direction, with the smallest operand tile that holds the chunk (up to 7 x 128 = 896
    // sequences in one round; the step time grows with NB through the gate phase and the exchange).
    // h->rec_chunks < 7 (asr_set_recurrence_chunks) trades a longer step for fewer SMs.
    const int chunks = h->rec_chunks < 1 ? 1 : (h->rec_chunks > 7 ? 7 : h->rec_chunks);
    int rows = (m.B + chunks - 1) / chunks;
    int NB = rows <= 16 ? 16 : rows <= 32 ? 32 : rows <= 64 ? 64 : rows <= 80 ? 80 : rows <= 96 ? 96 : 128;
    if (rows > rec3::kMaxNB) rows = rec3::kMaxNB;
    p.rows_per_chunk = rows;
    p.nchunks = (m.B + rows - 1) / rows;
    if ((size_t)2 * p.nchunks * 8 * rec3::kStageBytes > h->ws.rec_stage_ctas * 8192) {
        set_error("recurrence staging too small");
        return ASR_ERR_CAPACITY;
    }
    const bool want_dbg = h->rec_timeline;      // asr_stage_timing(h, 2): clock64 timeline of cluster 0 to stderr
    long long* dbg = nullptr;
    if (want_dbg) { ASR_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * 4096)); ASR_CUDA(cudaMemset(dbg, 0, sizeof(long long) * 8 * 4096)); }
    p.dbg = dbg;
    switch (NB) {
        case 16: ASR_TRY(rec3::launch<16>(p, st)); break;
        case 32: ASR_TRY(rec3::launch<32>(p, st)); break;
        case 64: ASR_TRY(rec3::launch<64>(p, st)); break;
        case 80: ASR_TRY(rec3::launch<80>(p, st)); break;
        case 96: ASR_TRY(rec3::launch<96>(p, st)); break;
        default: ASR_TRY(rec3::launch<128>(p, st)); break;
    }
    if (dbg) {
        std::vector<long long> hb(8 * 4096);
        ASR_CUDA(cudaStreamSynchronize(st));
        ASR_CUDA(cudaMemcpy(hb.data(), dbg, sizeof(long long) * hb.size(), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        const int L = m.len_sorted[0];
        double acc[6] = {};
        int cnt = 0;
        for (int s2 = 2; s2 + 1 < L; ++s2, ++cnt) {
            const long long* a = &hb[s2 * 8];
            acc[0] += (double)(a[1] - a[0]);
            acc[1] += (double)(a[2] - a[1]);
            acc[2] += (double)(a[4] - a[2]);
            acc[3] += (double)(a[5] - a[4]);
            acc[4] += (double)(a[6] - a[5]);
            acc[5] += (double)(hb[(s2 + 1) * 8] - a[6]);
        }
        fprintf(stderr, "[rec_tc3 dbg] layer %d NB=%d rows/chunk=%d chunks=%d steps=%d cycles/step: issue %.0f mma+commit %.0f "
                        "ld+gates %.0f fences %.0f bulk-issue %.0f exchange %.0f\n",
                layer, NB, rows, p.nchunks, L, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / cnt, acc[5] / cnt);
    }
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
