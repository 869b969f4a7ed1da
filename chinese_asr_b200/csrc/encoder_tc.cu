// Tensor-core LSTM recurrence (tcgen05 / TMEM, 3xTF32): the step-by-step h_{t-1} * W_hh^T product of
// nn.LSTM (reference util.py:1259) for one direction of one layer, for a chunk of NB sequences.
//
// Same cluster decomposition as encoder.cu (8 CTAs, CTA j owns the 4 gates of hidden units
// [32j, 32j+32) = 128 gate columns), but the per-step product runs on the 5th-gen tensor cores with
// the operands "swapped" so that the 128 gate columns fill the MMA M dimension and the batch rows
// are the (small) N dimension:
//        D[128 gate cols, rows] = W_slice[128, 256] * h[rows, 256]^T
//   * W_hi (tf32-rounded W_hh slice) is STATIONARY IN SHARED MEMORY (128 KB, K-major SWIZZLE_128B);
//     W_lo (the tf32 residual) is STATIONARY IN TENSOR MEMORY (128 lanes x 256 columns) and is fed
//     as the TMEM A operand.  Both are loaded once per layer.
//   * the B operand tile holds [h_hi rows | h_lo rows] (2*NB rows x 256, UMMA K-major layout), so per
//     K-step one N=2*NB MMA gives W_hi*h_hi and W_hi*h_lo in adjacent accumulator column blocks and
//     one N=NB MMA (A from TMEM) adds W_lo*h_hi: 64 tcgen05.mma per step, fp32 accumulation in TMEM.
//   * exchange of the new h: every CTA writes the 2*NB x 32 image of its slice (already split
//     into hi / lo and already in the swizzled operand layout) to a small global staging buffer and
//     issues ONE multicast bulk copy (cp.async.bulk ... .multicast::cluster) that lands it in the
//     operand tiles of all 8 CTAs and signals their `h_ready` mbarriers (complete_tx).
//   * no cluster-wide barrier in the loop: "every CTA's MMAs of this step are done" is a multicast
//     tcgen05.commit onto an mbarrier with count 8; "h_{t+1} is complete" is the tx-count of
//     `h_ready`.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "asr_internal.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace asr {

using namespace tcx;

// fast activations: ex2.approx + approximate division, absolute error < 1e-6 (the 3xTF32 products
// bound the overall accuracy of this kernel at ~1e-5)
__device__ __forceinline__ float sigmoid_tc(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_tc(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

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

struct RecTcParams {
    const float* xg;        // [rows, 2048] permuted gate pre-activations (bias included)
    const float* whh_hi;    // [2, 1024, 256] permuted, rn_tf32(W_hh)
    const float* whh_lo;    // [2, 1024, 256] permuted, rn_tf32(W_hh - hi)
    const float* x_in;      // residual input [rows, 512] packed, or nullptr
    float* y_packed;
    float* y_utt;
    float* h_fin;
    float* c_fin;
    float* stage;           // [gridDim.x][2048] global staging images
    const int* len_sorted;
    const int* toff;
    const int* uoff;
    int B;
    int nchunks;
    long long* dbg;         // optional [steps][8] clock64 timeline of CTA 0 (nullptr = off)
};

constexpr int kSlabA = 128 * 128;          // bytes: 128 rows x 32 fp32

template <int NB>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1)
lstm_rec_tc_kernel(RecTcParams p) {
    constexpr int kSlabB = 2 * NB * 128;    // bytes: [NB hi rows | NB lo rows] x 32 fp32
    constexpr int P = NB / 8;               // rows per gate thread (all 8 warps run the gate phase)
    constexpr int HC = NB / 2;              // accumulator columns read per warp
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* A_hi = smem;                                   // 8 slabs x 16 KB
    uint8_t* Bt = A_hi + 8 * kSlabA;                        // 8 slabs x kSlabB
    float* red = reinterpret_cast<float*>(Bt + 8 * kSlabB); // [NB][128]
    uint64_t* mma_done = reinterpret_cast<uint64_t*>(red + NB * 128);
    uint64_t* h_ready = mma_done + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
    int* s_len = reinterpret_cast<int*>(tmem_slot + 1);     // [NB]

    cg::cluster_group cluster = cg::this_cluster();
    const int j = (int)cluster.block_rank();
    const int cid = blockIdx.x / 8;
    const int dir = cid / p.nchunks;
    const int chunk = cid - dir * p.nchunks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = chunk * NB;
    const int nrows = min(NB, p.B - r0);
    float* stage = p.stage + (size_t)blockIdx.x * 2048;     // this CTA's slice image (kSlabB bytes used)

    // ---- one-time setup ------------------------------------------------------------------------
    for (int i = tid; i < (8 * kSlabB) / 16; i += 256)
        reinterpret_cast<float4*>(Bt)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < kSlabB / 16; i += 256)
        reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const float* wsrc = p.whh_hi + ((size_t)dir * kGates + j * 128) * kEncH;
        for (int idx = tid; idx < 128 * 64; idx += 256) {
            const int m = idx >> 6, k = (idx & 63) << 2;
            const float4 v = *reinterpret_cast<const float4*>(wsrc + (size_t)m * kEncH + k);
            *reinterpret_cast<float4*>(A_hi + (k >> 5) * kSlabA + sw128_offset(m, k & 31)) = v;
        }
    }
    if (tid < NB) s_len[tid] = tid < nrows ? p.len_sorted[r0 + tid] : 0;
    if (tid == 0) {
        mbar_init(mma_done, 8);        // one multicast commit from each CTA of the cluster
        mbar_init(h_ready, 1);         // one arrive.expect_tx by this CTA + 8 multicast bulk copies
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base + 256;               // accumulator columns [256, 256 + 4*NB): two (hi | lo) sets
    if (warp < 4) {
        // W_lo slice -> TMEM lanes (gate column m = 32*warp + lane), columns [0, 256)
        const float* wsrc = p.whh_lo + ((size_t)dir * kGates + j * 128 + 32 * warp + lane) * kEncH;
#pragma unroll 1
        for (int c0 = 0; c0 < kEncH; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(wsrc + c0 + 4 * q);
                r[4 * q] = __float_as_uint(v.x); r[4 * q + 1] = __float_as_uint(v.y);
                r[4 * q + 2] = __float_as_uint(v.z); r[4 * q + 3] = __float_as_uint(v.w);
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

    const int uu = lane;                                   // hidden unit 32j + uu (gate threads)
    float c_reg[P], h_reg[P];
#pragma unroll
    for (int q = 0; q < P; ++q) { c_reg[q] = 0.f; h_reg[q] = 0.f; }

    cluster.sync();      // every CTA's barriers / tiles are initialised before any multicast arrives

    const uint64_t dA0 = kmajor_sw128_desc(smem_u32(A_hi)), dB0 = kmajor_sw128_desc(smem_u32(Bt));
    constexpr uint32_t idesc2 = idesc_tf32(128, 2 * NB);
    constexpr uint32_t idesc1 = idesc_tf32(128, NB);

    for (int s = 0; s < Lc; ++s) {
        const int t = dir == 0 ? s : Lc - 1 - s;
        int nact = 0;
        for (int i = 0; i < nrows; ++i) nact += (s_len[i] > t) ? 1 : 0;
        const int row_t = p.toff[t] + r0;

        // (a) one thread issues the 64 MMAs of this step once h_t has landed
        if (warp == 4) {
            if (lane == 0) {
                if (s > 0) mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
                tc_fence_after();
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 0] = clock64();
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) {
                    // descriptors differ from the slab-0 ones by compile-time constants only
                    const uint64_t dA = dA0 + (uint64_t)(((kk >> 2) * kSlabA + (kk & 3) * 32) >> 4);
                    const uint64_t dB = dB0 + (uint64_t)(((kk >> 2) * kSlabB + (kk & 3) * 32) >> 4);
                    // two independent accumulators (even / odd K-steps) halve the dependent-MMA chain
                    const uint32_t dacc = tmem_d + (uint32_t)((kk & 1) * 2 * NB);
                    umma_tf32_ss(dacc, dA, dB, idesc2, kk > 1 ? 1u : 0u);            // W_hi * [h_hi | h_lo]
                    umma_tf32_ts(dacc, tmem_base + (uint32_t)(8 * kk), dB, idesc1, 1u);     // + W_lo * h_hi
                }
                umma_commit_mc(mma_done, (uint16_t)0xFF);
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 1] = clock64();
            }
            __syncwarp();
        }

        // (b) prefetch this step's input-projection pre-activations
        float xi[P], xf[P], xgg[P], xo[P], xres[P], yv[P];
        const int ocol = dir * kEncH + 32 * j + uu;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = warp + 8 * q;
            xi[q] = xf[q] = xgg[q] = xo[q] = xres[q] = yv[q] = 0.f;
            if (i < nact) {
                if (p.x_in) xres[q] = __ldg(p.x_in + (size_t)(row_t + i) * kEnc + ocol);
                const float* g = p.xg + (size_t)(row_t + i) * (2 * kGates) + dir * kGates + j * 128 + uu;
                xi[q] = __ldg(g);
                xf[q] = __ldg(g + 32);
                xgg[q] = __ldg(g + 64);
                xo[q] = __ldg(g + 96);
            }
        }

        // (c) all 8 CTAs' MMAs of this step are complete: the accumulator is final and every
        //     operand tile of the cluster may be overwritten
        mbar_wait(mma_done, (uint32_t)(s & 1));
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 2] = clock64();
        {
            // accumulator rows (gate column m = 32*(warp&3) + lane); warps w and w+4 share a TMEM lane
            // quarter and split the columns.  D = cols[n] (W_hi h_hi + W_lo h_hi) + cols[NB+n] (W_hi h_lo)
            uint32_t d1[HC], d2[HC], d3[HC], d4[HC];
            const int qd = warp & 3, half = warp >> 2;
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * qd) << 16) + (uint32_t)(half * HC);
            if (HC == 16) {
                tmem_ld16(taddr, d1); tmem_ld16(taddr + NB, d2);
                tmem_ld16(taddr + 2 * NB, d3); tmem_ld16(taddr + 3 * NB, d4);
            } else {
                tmem_ld8(taddr, d1); tmem_ld8(taddr + NB, d2);
                tmem_ld8(taddr + 2 * NB, d3); tmem_ld8(taddr + 3 * NB, d4);
            }
            const int m = 32 * qd + lane;
#pragma unroll
            for (int n = 0; n < HC; ++n)
                red[(half * HC + n) * 128 + m] = (__uint_as_float(d1[n]) + __uint_as_float(d3[n])) +
                                                 (__uint_as_float(d2[n]) + __uint_as_float(d4[n]));
        }
        tc_fence_before();
        __syncthreads();
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 3] = clock64();
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = warp + 8 * q;
            if (i < nact) {
                const float* rr = red + i * 128 + uu;
                const float gi = xi[q] + rr[0];
                const float gf = xf[q] + rr[32];
                const float gg = xgg[q] + rr[64];
                const float go = xo[q] + rr[96];
                const float c = sigmoid_tc(gf) * c_reg[q] + sigmoid_tc(gi) * tanh_tc(gg);
                const float hh = sigmoid_tc(go) * tanh_tc(c);
                c_reg[q] = c;
                h_reg[q] = hh;
                // slice image in the operand layout: hi row i, lo row NB + i, column uu of K-slab j
                const float hi = rn_tf32(hh);
                const float lo = rn_tf32(hh - hi);
                const uint32_t off = sw128_offset(i, uu) >> 2;
                stage[off] = hi;
                stage[NB * 32 + off] = lo;
                yv[q] = hh + xres[q];                  // residual add (util.py:1284-1291)
            }
        }
        if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 4] = clock64();
        if (s + 1 < Lc) {
            __threadfence();           // the image is visible at L2 ...
            fence_proxy_async();       // ... and ordered before the async-proxy bulk read
            __syncthreads();
            if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[s * 8 + 5] = clock64();
            if (tid == 0) {
                mbar_expect_tx(h_ready, 8u * (uint32_t)kSlabB);
                bulk_g2s_multicast(Bt + j * kSlabB, stage, (uint32_t)kSlabB, h_ready, (uint16_t)0xFF);
                if (p.dbg && blockIdx.x == 0) p.dbg[s * 8 + 6] = clock64();
            }
        }
        // layer output stores are off the critical path: issued after the exchange has been started
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = warp + 8 * q;
            if (i < nact) {
                const size_t row = (size_t)(row_t + i);
                if (p.y_packed) p.y_packed[row * kEnc + ocol] = yv[q];
                if (p.y_utt) p.y_utt[(size_t)(p.uoff[r0 + i] + t) * kEnc + ocol] = yv[q];
            }
        }
    }

#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int i = warp + 8 * q;
        if (i < nrows) {
            const int ocol = dir * kEncH + 32 * j + uu;
            p.h_fin[(size_t)(r0 + i) * kEnc + ocol] = h_reg[q];
            p.c_fin[(size_t)(r0 + i) * kEnc + ocol] = c_reg[q];
        }
    }
    tc_fence_before();
    cluster.sync();                                     // nobody exits while peers may still signal it
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int NB>
static int launch_tc(const RecTcParams& p, cudaStream_t st) {
    const size_t smem = 8 * (size_t)kSlabA + 8 * (size_t)(2 * NB * 128) + (size_t)NB * 128 * 4 + 1024 + 64 + NB * 4;
    static bool attr = false;
    if (!attr) {
        ASR_CUDA(cudaFuncSetAttribute(lstm_rec_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    lstm_rec_tc_kernel<NB><<<2 * p.nchunks * 8, 256, smem, st>>>(p);
    ASR_CHECK_LAUNCH();
    return ASR_OK;
}

int launch_lstm_recurrence_tc(asr_handle* h, int layer, const float* xg, const float* x_in, float* y_packed,
                              float* y_utt, float* h_fin, float* c_fin, cudaStream_t st) {
    const BatchMeta& m = h->meta;
    RecTcParams p{};
    p.xg = xg;
    p.whh_hi = h->w.enc_w_hh_hi[layer];
    p.whh_lo = h->w.enc_w_hh_lo[layer];
    p.x_in = x_in;
    p.y_packed = y_packed;
    p.y_utt = y_utt;
    p.h_fin = h_fin;
    p.c_fin = c_fin;
    p.stage = h->ws.rec_stage;
    p.len_sorted = m.d_len_sorted;
    p.toff = m.d_toff;
    p.uoff = m.d_uoff_sorted;
    p.B = m.B;
    const int NB = m.B > 128 ? 32 : 16;
    p.nchunks = (m.B + NB - 1) / NB;
    if ((size_t)2 * p.nchunks * 8 > h->ws.rec_stage_ctas) { set_error("recurrence staging too small"); return ASR_ERR_CAPACITY; }
    static const bool want_dbg = getenv("ASR_B200_REC_DBG") != nullptr;
    long long* dbg = nullptr;
    if (want_dbg) { ASR_CUDA(cudaMalloc(&dbg, sizeof(long long) * 8 * 4096)); ASR_CUDA(cudaMemset(dbg, 0, sizeof(long long) * 8 * 4096)); }
    p.dbg = dbg;
    if (NB == 32) ASR_TRY(launch_tc<32>(p, st));
    else ASR_TRY(launch_tc<16>(p, st));
    if (dbg) {
        std::vector<long long> hbuf(8 * 4096);
        ASR_CUDA(cudaStreamSynchronize(st));
        ASR_CUDA(cudaMemcpy(hbuf.data(), dbg, sizeof(long long) * hbuf.size(), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        const int L = m.len_sorted[0];
        double acc[7] = {};
        int cnt = 0;
        for (int s2 = 2; s2 + 1 < L; ++s2, ++cnt) {
            const long long* a = &hbuf[s2 * 8];
            acc[0] += (double)(a[1] - a[0]);         // MMA issue
            acc[1] += (double)(a[2] - a[1]);         // MMA exec + multicast commit
            acc[2] += (double)(a[3] - a[2]);         // TMEM ld + red + sync
            acc[3] += (double)(a[4] - a[3]);         // gates + stores
            acc[4] += (double)(a[5] - a[4]);         // fences + sync
            acc[5] += (double)(a[6] - a[5]);         // bulk issue
            acc[6] += (double)(hbuf[(s2 + 1) * 8] - a[6]);   // exchange latency until next h_ready
        }
        fprintf(stderr, "[rec_tc dbg] layer %d NB=%d steps=%d cycles/step: issue %.0f mma+commit %.0f ld+red %.0f gates %.0f fences %.0f bulk-issue %.0f exchange %.0f\n",
                layer, NB, L, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt);
    }
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
