// Tensor-core LSTM recurrence, weights fully resident in TENSOR MEMORY.
//
// An earlier engine kept W_hi in shared memory and re-read its 128 KB through the UMMA operand path
// every step (measured: the 64 MMAs of a step were bound by ~64 B/clk of shared-memory operand
// fetch, ~3.9k cycles).  Here the whole recurrent weight slice of a CTA lives in TMEM:
//     columns [  0,128) : W_hi  = fp16(W_hh slice), 2 per column             128 lanes x 256 fp16
//     columns [128,384) : W cross, per 8 values of k [8 x bf16(w) | 8 x bf16(w - w_hi)]   (4 bytes per k)
//     columns [384,512) : fp32 accumulator D[128 gate cols, NB rows]
// (the split-precision operands of gemm_tc.cu: kSplitWeight on the A side, kSplitAct on the h side), so
// shared memory only holds the h operand tiles, which lets one cluster carry up to NB = 80 sequences
// (one round of <= 15 clusters covers 2 directions x 512 sequences).  Per step 48 MMAs, all kind::f16:
//     D += W_hi (TMEM, fp16) * h_hi^T                          16 x (K = 16 values of k)
//     D += W_x  (TMEM, bf16) * [bf16(h_lo) | bf16(h)]^T        32 x (K = 16 = 8 values of k): w*h_lo + w_lo*h
// (the tf32 form issued 80: an MMA costs ~65 cycles here whatever its shape).  The cross terms are 2^-11
// of the product, which bf16's 2^-9 relative accuracy carries to ~2^-20.
// Exchange of h_{t+1}: each CTA writes the image of its 32-unit slice
// ([cross | hi] rows, already in the swizzled UMMA layouts) to global staging and multicasts it
// with one cp.async.bulk into the tiles of all 8 CTAs; mbarriers only, no cluster barrier per step.
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
__device__ __forceinline__ void lstm_cell(float gi, float gf, float gg, float go, float c_prev, float& c, float& h) {
    const float A = 1.f + __expf(-fmaxf(gi, -40.f));
    const float Bf = 1.f + __expf(-fmaxf(gf, -40.f));
    const float G = __expf(2.f * fminf(gg, 20.f)) + 1.f;
    const float O = 1.f + __expf(-fmaxf(go, -40.f));
    const float r = rcp_f(A * G);
    const float si = r * G;                       // 1 / A
    const float tg = 1.f - 2.f * (r * A);         // 1 - 2 / G
    c = rcp_f(Bf) * c_prev + si * tg;
    const float C = __expf(2.f * fminf(c, 20.f)) + 1.f;
    h = (C - 2.f) * rcp_f(O * C);
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
// shared memory of this CTA -> shared memory of cluster CTA `rank` (same offsets), completion on the
// destination CTA's mbarrier
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
// D[tmem] += A[tmem, bf16] * B[smem desc, bf16]
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
// byte offset of bf16 element (row, kk in 0..31) inside a [rows x 32 bf16] SWIZZLE_64B K-major slab
__device__ __forceinline__ uint32_t sw64_offset(int row, int kk) {
    return (uint32_t)(row * 64 + ((((kk >> 3) ^ ((row >> 1) & 3)) << 4) | ((kk & 7) << 1)));
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with fp16 operands (format 0)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of byte `b` (0..127) of row `row` inside a [rows x 128 B] SWIZZLE_128B K-major slab
__device__ __forceinline__ uint32_t sw128_byte(int row, int b) {
    return (uint32_t)(row * 128 + ((((b >> 4) ^ (row & 7)) << 4) | (b & 15)));
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t* r) { tmem_ld8(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t* r) { tmem_ld16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t* r) { tmem_ld32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<10>(uint32_t taddr, uint32_t* r) {
    tmem_ld8(taddr, r);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                 : "=r"(r[8]), "=r"(r[9])
                 : "r"(taddr + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <>
__device__ __forceinline__ void tmem_ld_cols<24>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
                 : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_cols<20>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19])
                 : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kThreads = 544;        // 16 gate warps + 1 MMA-issue warp
constexpr int kIssueWarp = 16;

struct Params {
    const float* xg;            // [rows, 2048] permuted gate pre-activations (bias included)
    const hi_t* whh_hi;         // [2, 1024, 256] permuted, fp16(W_hh)                       (gemm_tc.cu kSplitWeight)
    const uint32_t* whh_x;      // [2, 1024, 256] words: per 8 k, 8 x bf16(w) then 8 x bf16(w - hi)
    const float* x_in;
    int xchg_dsmem;         // 1: exchange h through distributed shared memory, 0: through global staging
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

constexpr int kMaxNB = 128;              // sequences per cluster: the accumulator takes TMEM columns [384, 384 + NB)
constexpr int kStageBytes = kMaxNB * 192;    // per-CTA staging slot (largest NB)

template <int NB>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kThreads, 1)
lstm_rec_tc3_kernel(Params p) {
    // 16 warps: warp = (TMEM lane quarter q = warp & 3, column group warp >> 2).  TMEM lane m of the CTA's
    // 128 gate rows is (unit m >> 2, gate m & 3): the four gates of a unit sit in 4 adjacent lanes, so a
    // 4x4 transpose by warp shuffles gives every lane complete (i, f, g, o) pre-activations of P cells
    // (unit, batch row) - no shared-memory transposition and no block barrier in the gate phase.
    constexpr int CW = NB / 4;               // accumulator columns (batch rows) per warp
    constexpr int P = CW / 4;                // cells per lane: rows cg * CW + 4 b + (lane & 3)
    constexpr int kX = NB * 128;             // bytes of one cross slab: 32 values of k x [bf16(h_lo) | bf16(h)]
    constexpr int kHi = NB * 64;             // bytes of one fp16 hi slab
    constexpr int kSlab = kX + kHi;          // [cross | hi] of one 32-wide K range
    static_assert(NB % 16 == 0, "NB");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* T = smem;                                             // 8 slabs x kSlab
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(T + 8 * kSlab);
    uint64_t* h_ready = mma_done + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
    int* s_len = reinterpret_cast<int*>(tmem_slot + 1);            // [NB]

    cg::cluster_group cluster = cg::this_cluster();
    const int j = (int)cluster.block_rank();
    const int cid = blockIdx.x / 8;
    const int dir = cid / p.nchunks;
    const int chunk = cid - dir * p.nchunks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_ax = tmem_base + 128;
    const uint32_t tmem_d = tmem_base + 384;
    if (warp < 4) {
        const int m = 32 * warp + lane;                                          // TMEM lane = (unit m >> 2, gate m & 3)
        const size_t wrow = (size_t)dir * kGates + j * 128 + (m & 3) * 32 + (m >> 2);
        const uint32_t* whi = reinterpret_cast<const uint32_t*>(p.whh_hi + wrow * kEncH);     // 128 words of fp16 pairs
#pragma unroll 1
        for (int c0 = 0; c0 < kEncH / 2; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 v = *reinterpret_cast<const uint4*>(whi + c0 + 4 * q);
                r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
            tmem_st32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
        }
        const uint32_t* wx = p.whh_x + wrow * kEncH;                                          // 256 words
#pragma unroll 1
        for (int c0 = 0; c0 < kEncH; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 v = *reinterpret_cast<const uint4*>(wx + c0 + 4 * q);
                r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
            tmem_st32(tmem_ax + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
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
#pragma unroll
    for (int q = 0; q < P; ++q) { c_reg[q] = 0.f; h_reg[q] = 0.f; }

    cluster.sync();

    const uint32_t t_base = smem_u32(T);
    constexpr uint32_t idesc_h = idesc_f16(128, NB);
    constexpr uint32_t idesc_b = idesc_bf16(128, NB);

    // rows are sorted by decreasing length: the active count follows t incrementally
    int nact = dir == 0 ? nrows : 0;
    int prev_t = 0, prev_row_t = 0, prev_nact = 0;
    float yv[P];
#pragma unroll
    for (int q = 0; q < P; ++q) yv[q] = 0.f;
    // layer output of one step: y = h + residual input (util.py:1284-1291), and optionally its operand split
    // for the next GEMM.  Called one step late (after the next step's MMAs are issued) so that it never delays them.
    auto store_outputs = [&](int t, int row_t, int n_act) {
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = col0 + 4 * q;
            const bool act = i < n_act;
            const float y = yv[q];
            // operand split for the consuming GEMM: the lane 4 away holds the neighbouring unit of the
            // same row; even units pack the two residuals, odd units the two values (bf16 pairs)
            const float hi = hi_part(y);
            const float lo = y - hi;
            const float lo_n = __shfl_xor_sync(0xffffffffu, lo, 4);
            const float y_n = __shfl_xor_sync(0xffffffffu, y, 4);
            if (act) {
                const size_t row = (size_t)(row_t + i);
                const size_t urow = p.y_utt ? (size_t)(p.uoff[r0 + i] + t) : 0;
                if (p.y_packed) p.y_packed[row * kEnc + ocol] = y;
                if (p.y_utt) p.y_utt[urow * kEnc + ocol] = y;
                if (p.y_hi) {
                    const size_t srow = p.y_utt ? urow : row;      // last layer: rows as `enc`
                    p.y_hi[srow * kEnc + ocol] = __float2half_rn(hi);
                    const bool odd = (ocol & 1) != 0;
                    const __nv_bfloat162 pk = odd ? __floats2bfloat162_rn(y_n, y) : __floats2bfloat162_rn(lo, lo_n);
                    // block of 8 floats -> 8 words: words 0-3 residual pairs, words 4-7 value pairs
                    uint32_t* blk = p.y_cross + srow * kEnc + (size_t)(ocol & ~7);
                    blk[(odd ? 4 : 0) + ((ocol & 7) >> 1)] = *reinterpret_cast<const uint32_t*>(&pk);
                }
            }
        }
    };

    for (int s = 0; s < Lc; ++s) {
        if (warp == kIssueWarp) {
            if (lane == 0) {
                if (s > 0) mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
                tc_fence_after();
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 0] = clock64();
                const uint64_t dX0 = kmajor_sw128_desc(t_base);
                const uint64_t dHi0 = kmajor_sw64_desc(t_base + kX);
                // every MMA consumes 8 TMEM columns of A and 32 bytes of a B row.  The cross terms go first: the
                // tensor core truncates every accumulation, and while the accumulator only holds the 2^-11-sized
                // cross sum those 32 truncations are 2^-11 smaller; only the 16 hi MMAs truncate at full magnitude
                // (measured on the GEMM engine: 3x less error than interleaving, gemm_tc.cu).
#pragma unroll
                for (int kb = 0; kb < 32; ++kb) {        // 8 values of k each: w * h_lo + w_lo * h
                    const uint64_t adv = (uint64_t)(((kb >> 2) * kSlab + (kb & 3) * 32) >> 4);
                    umma_bf16_ts(tmem_d, tmem_ax + (uint32_t)(8 * kb), dX0 + adv, idesc_b, kb > 0 ? 1u : 0u);
                }
#pragma unroll
                for (int kh = 0; kh < 16; ++kh) {        // 16 values of k each
                    const uint64_t adv = (uint64_t)(((kh >> 1) * kSlab + (kh & 1) * 32) >> 4);
                    umma_bf16_ts(tmem_d, tmem_base + (uint32_t)(8 * kh), dHi0 + adv, idesc_h, 1u);
                }
                umma_commit_mc(mma_done, (uint16_t)0xFF);
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 1] = clock64();
            }
            __syncwarp();
        }

        const int t = dir == 0 ? s : Lc - 1 - s;
        if (dir == 0) { while (nact > 0 && s_len[nact - 1] <= t) --nact; }
        else { while (nact < nrows && s_len[nact] > t) ++nact; }
        const int row_t = p.toff[t] + r0;

        if (warp < kIssueWarp) {
        // gate warps: the issue warp above does nothing else, so it never arrives late at the step's barrier
        // the stores of the previous step first (registers only), then this step's loads: their latency is
        // covered by the MMAs in flight (measured: loading the residual inside the deferred store instead makes
        // the gate warps arrive ~500 cycles late at mma_done)
        if (s > 0) store_outputs(prev_t, prev_row_t, prev_nact);
        float xi[P], xf[P], xgg[P], xo[P], xres[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = col0 + 4 * q;
            xi[q] = xf[q] = xgg[q] = xo[q] = xres[q] = 0.f;
            if (i < nact) {
                if (p.x_in) xres[q] = __ldg(p.x_in + (size_t)(row_t + i) * kEnc + ocol);
                const float* g = p.xg + (size_t)(row_t + i) * (2 * kGates) + dir * kGates + j * 128 + uu;
                xi[q] = __ldg(g);
                xf[q] = __ldg(g + 32);
                xgg[q] = __ldg(g + 64);
                xo[q] = __ldg(g + 96);
            }
        }

        mbar_wait(mma_done, (uint32_t)(s & 1));
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 2] = clock64();
        {
            uint32_t d[CW];
            tmem_ld_cols<CW>(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * CW), d);
            tc_fence_before();
            const bool o1 = (jr & 1) != 0, o2 = (jr & 2) != 0;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                // 4x4 transpose across the 4 lanes of a unit: in = this lane's gate for rows 4q..4q+3,
                // out = gates i, f, g, o of row 4q + jr
                const float v0 = __uint_as_float(d[4 * q]), v1 = __uint_as_float(d[4 * q + 1]);
                const float v2 = __uint_as_float(d[4 * q + 2]), v3 = __uint_as_float(d[4 * q + 3]);
                const float ra = __shfl_xor_sync(0xffffffffu, o1 ? v0 : v1, 1);
                const float rb = __shfl_xor_sync(0xffffffffu, o1 ? v2 : v3, 1);
                const float x0 = o1 ? ra : v0, x1 = o1 ? v1 : ra;
                const float x2 = o1 ? rb : v2, x3 = o1 ? v3 : rb;
                const float ua = __shfl_xor_sync(0xffffffffu, o2 ? x0 : x2, 2);
                const float ub = __shfl_xor_sync(0xffffffffu, o2 ? x1 : x3, 2);
                const float di = o2 ? ua : x0, df = o2 ? ub : x1, dg = o2 ? x2 : ua, dO = o2 ? x3 : ub;
                const int i = col0 + 4 * q;
                if (i < nact) {
                    const float gi = xi[q] + di;
                    const float gf = xf[q] + df;
                    const float gg = xgg[q] + dg;
                    const float go = xo[q] + dO;
                    float c, hh;
                    lstm_cell(gi, gf, gg, go, c_reg[q], c, hh);
                    c_reg[q] = c;
                    h_reg[q] = hh;
                    const float hi = hi_part(hh);
                    const float lo = hh - hi;
                    // image of this CTA's 32-unit slab of the next B operand: into global staging, or (all
                    // 8 CTAs' MMAs of this step are complete - mma_done counts 8 commits) straight into
                    // the slab's place in the local operand tile
                    uint8_t* img = p.xchg_dsmem ? T + j * kSlab : stage;
                    const int xb = (uu >> 3) * 32 + (uu & 7) * 2;       // 8-k block: 16 bytes of bf16(h_lo), 16 of bf16(h)
                    *reinterpret_cast<__nv_bfloat16*>(img + sw128_byte(i, xb)) = __float2bfloat16_rn(lo);
                    *reinterpret_cast<__nv_bfloat16*>(img + sw128_byte(i, xb + 16)) = __float2bfloat16_rn(hh);
                    *reinterpret_cast<__half*>(img + kX + sw64_offset(i, uu)) = __float2half_rn(hi);
                    yv[q] = hh + xres[q];
                }
            }
        }
        }
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 4] = clock64();
        if (s + 1 < Lc) {
            if (p.xchg_dsmem) {
                // the local slab is in place; push it to the 7 peers with one bulk copy each (issued by 7
                // different warps), each completing on the receiver's h_ready: no global round trip and
                // no membar.gl on the step's critical path
                fence_proxy_async();
                __syncthreads();
                if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 5] = clock64();
                if (tid == 0) mbar_expect_tx(h_ready, 7u * (uint32_t)kSlab);
                if (lane == 0 && warp >= 1 && warp <= 7) {
                    const uint32_t peer = (uint32_t)((j + warp) & 7);
                    const uint32_t src = smem_u32(T + j * kSlab);
                    bulk_s2s(mapa_u32(src, peer), src, (uint32_t)kSlab, mapa_u32(smem_u32(h_ready), peer));
                }
                if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 6] = clock64();
            } else {
                __threadfence();
                fence_proxy_async();
                __syncthreads();
                if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 5] = clock64();
                if (tid == 0) {
                    mbar_expect_tx(h_ready, 8u * (uint32_t)kSlab);
                    bulk_g2s_multicast(T + j * kSlab, stage, (uint32_t)kSlab, h_ready, (uint16_t)0xFF);
                    if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 6] = clock64();
                }
            }
        }
        prev_t = t; prev_row_t = row_t; prev_nact = nact;
    }
    if (Lc > 0 && warp < kIssueWarp) store_outputs(prev_t, prev_row_t, prev_nact);

#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int i = col0 + 4 * q;
        if (i < nrows && warp < kIssueWarp) {
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

template <int NB>
static int launch(const Params& p, cudaStream_t st) {
    const size_t smem = 8 * (size_t)(NB * 192) + 1024 + 64 + NB * 4;
    ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&lstm_rec_tc3_kernel<NB>), smem));
    lstm_rec_tc3_kernel<NB><<<2 * p.nchunks * 8, kThreads, smem, st>>>(p);
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}

}  // namespace rec3

size_t rec3_stage_bytes_per_cta() { return rec3::kStageBytes; }

int launch_lstm_recurrence_tc3(asr_handle* h, int layer, const float* xg, const float* x_in, float* y_packed,
                               float* y_utt, float* h_fin, float* c_fin, cudaStream_t st, hi_t* split_hi,
                               float* split_lo) {
    const BatchMeta& m = h->meta;
    rec3::Params p{};
    p.xchg_dsmem = 0;      // measured: 7 DSMEM bulk copies per CTA per step (10.9 -> 17.9 ms per batch) lose to one multicast from L2
    p.y_hi = split_hi;
    p.y_cross = reinterpret_cast<uint32_t*>(split_lo);
    p.xg = xg;
    p.whh_hi = h->w.enc_w_hh_hi16[layer];
    p.whh_x = reinterpret_cast<const uint32_t*>(h->w.enc_w_hh_x[layer]);
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
    // <= 15 clusters of 8 CTAs are co-resident on a B200 (measured): aim at one round, i.e. at most
    // 7 chunks per direction, with the smallest operand tile that holds the chunk (up to 7 x 128 = 896
    // sequences in one round; the step time grows with NB only through the gate phase and the exchange).
    // h->rec_chunks < 7 (asr_set_recurrence_chunks) trades a longer step for fewer SMs.  Measured at 512 x 10 s
    // (ms per batch, SMs): 7 chunks 7.2 on 112; 6 chunks (96 sequences per cluster) 8.0 on 96; 4 chunks (128 per
    // cluster) 10.8 - 13.6 on 64: the gate phase is issue-bound and grows faster than the chunk, so 7 stays the
    // default and the knob is for experiments only.
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
