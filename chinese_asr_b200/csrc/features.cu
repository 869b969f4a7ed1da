// Feature kernels: get_log_mel (reference data.py:167-280) and per-utterance CMVN (main.py:37).
//
//   F1 logmel_kernel   : pre-emphasis (data.py:201-202) -> Hann-400 placed at offset 56 of a
//                        512-sample frame, hop 160 (torch.stft center=False, data.py:205-209) ->
//                        512-point real FFT (256-point complex radix-4 Stockham in shared memory +
//                        split) -> power (data.py:221) -> sparse mel filterbank (data.py:222; the
//                        [257,80] matrix has 504 non-zeros) -> zero->eps, log (data.py:223-224).
//                        One warp per frame; each PCM sample is read from HBM once per frame it
//                        belongs to (3.2 frames overlap -> L1/L2 hits), mel rows written coalesced.
//   F2 delta_cmvn_kernel: 9-tap identity/delta/delta-delta correlation with zero padding
//                        (data.py:157-162), drop T%3 tail frames and stack 3 frames -> [L,720]
//                        with column c*240 + j*80 + m (data.py:244-249), then (x-mean)/(std+1e-6)
//                        per column with the unbiased std (main.py:37).
// Both are HBM/L2-bound streaming kernels: algorithmic bytes per utterance 4N + 4*720*L.
#include <cuda_bf16.h>
#include <math.h>

#include <vector>

#include "asr_internal.cuh"

namespace asr {

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// complex add / subtract as ONE packed fp32x2 instruction each (FADD2 / FFMA2 on sm_100): the butterflies are add-bound
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }        // a * (-i)

constexpr int kLmFrames = 32;                         // frames per CTA: 16 pairs, two per warp
constexpr int kLmSpan = (kLmFrames - 1) * kHop + kWinOff + kWin;      // pre-emphasised samples a CTA touches (5416)
constexpr int kLmSpanPad = (kLmSpan + 7) & ~7;
constexpr int kLmBuf = 512 + 32;                      // complex points per warp buffer: index i lives at i + (i >> 4)
constexpr size_t kLmSmem = sizeof(float) * (kLmSpanPad + kWin) + sizeof(float2) * (512 + 8 * kLmBuf);

// PCM sample as the float32 soundfile.read(dtype='float32') hands the reference (data.py:111): float32 files as
// stored, 16-bit files as x / 32768 (exact: a power of two), so both sample types give bit-identical frames.
__device__ __forceinline__ float pcm_sample(const float* x, long long i) { return x[i]; }
__device__ __forceinline__ float pcm_sample(const short* x, long long i) { return (float)x[i] * (1.0f / 32768.0f); }

// 8 consecutive samples starting at a 16-byte aligned address (+ the one after them) -> 9 floats
__device__ __forceinline__ void pcm_load9(const short* x, float (&v)[9]) {
    const uint4 r = *reinterpret_cast<const uint4*>(x);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = (float)(short)(w[i] & 0xffffu) * (1.0f / 32768.0f);
        v[2 * i + 1] = (float)(short)(w[i] >> 16) * (1.0f / 32768.0f);
    }
    v[8] = (float)x[8] * (1.0f / 32768.0f);
}
__device__ __forceinline__ void pcm_load9(const float* x, float (&v)[9]) {
    const float4 a = *reinterpret_cast<const float4*>(x), b = *reinterpret_cast<const float4*>(x + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = x[8];
}

// forward 8-point DFT in registers (radix-2 decimation in time; w = e^(-2 pi i / 8))
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    const float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), a3 = cmul_negi(csub(v[2], v[6]));
    const float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]), a6 = cadd(v[3], v[7]), a7 = cmul_negi(csub(v[3], v[7]));
    const float2 b0 = cadd(a0, a2), b1 = cadd(a1, a3), b2 = csub(a0, a2), b3 = csub(a1, a3);
    const float2 b4 = cadd(a4, a6), b5 = cadd(a5, a7), b6 = csub(a4, a6), b7 = csub(a5, a7);
    const float2 t5 = make_float2(h * (b5.x + b5.y), h * (b5.y - b5.x));          // b5 * (1 - i) / sqrt 2
    const float2 t6 = cmul_negi(b6);
    const float2 t7 = make_float2(h * (b7.y - b7.x), -h * (b7.x + b7.y));         // b7 * (-1 - i) / sqrt 2
    v[0] = cadd(b0, b4); v[4] = csub(b0, b4);
    v[1] = cadd(b1, t5); v[5] = csub(b1, t5);
    v[2] = cadd(b2, t6); v[6] = csub(b2, t6);
    v[3] = cadd(b3, t7); v[7] = csub(b3, t7);
}

// One CTA = 32 consecutive frames of one utterance (grid: frame chunks x utterances).
//   1. the chunk's PCM span is read ONCE with 16-byte loads (8 int16 or 4 + 4 float samples), pre-emphasised
//      (data.py:201-202) and kept in shared memory (the round-1 kernel fetched every sample ~3.2 times, scalar);
//   2. each warp transforms TWO real frames per complex FFT: z = frame_a + i frame_b, 512-point Stockham radix-8 in
//      three passes (64 butterflies per pass = 2 per lane, 16 points in registers; the first pass reads straight from
//      the pre-emphasised samples x window, torch.stft center=False with the Hann-400 at offset 56, data.py:205-209),
//      then X_a[k] = (Z[k] + conj Z[-k]) / 2, X_b[k] = (Z[k] - conj Z[-k]) / 2i - no 512-point twiddle pass;
//   3. power (data.py:221), sparse mel filterbank (data.py:222), zero -> eps, log (data.py:223-224) for both frames.
template <typename S>
__global__ void __launch_bounds__(256, 3)
logmel_kernel(const S* __restrict__ pcm, const long long* __restrict__ pcm_off,
              const int* __restrict__ frame_off, const float* __restrict__ g_window,
              const float2* __restrict__ g_tw512, const int* __restrict__ mel_start,
              const int* __restrict__ mel_len, const float* __restrict__ mel_w, int mel_maxw,
              float preemph, float* __restrict__ mel_out) {
    extern __shared__ __align__(16) uint8_t lm_smem[];
    float* s_pe = reinterpret_cast<float*>(lm_smem);                   // [kLmSpanPad] pre-emphasised samples
    float* s_win = s_pe + kLmSpanPad;                                  // [400]
    float2* s_tw = reinterpret_cast<float2*>(s_win + kWin);            // [512] exp(-2 pi i k / 512)
    float2* s_buf = s_tw + 512;                                        // [8 warps][kLmBuf]

    const int u = blockIdx.y;
    const int f0 = frame_off[u];
    const int T = frame_off[u + 1] - f0;
    const int t0 = blockIdx.x * kLmFrames;
    if (t0 >= T) return;
    const int nf = min(kLmFrames, T - t0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long x0 = pcm_off[u] + (long long)t0 * kHop;            // first sample of the chunk
    const int span = (nf - 1) * kHop + kWinOff + kWin;                 // pe[i], i < span, is read below

    for (int i = tid; i < kWin; i += 256) s_win[i] = g_window[i];
    for (int i = tid; i < 512; i += 256) s_tw[i] = g_tw512[i];
    // pe[i] = x[i + 1] - preemph * x[i]  (two roundings, like the reference's tensor expression)
    if ((reinterpret_cast<uintptr_t>(pcm + x0) & 15) == 0) {        // 16-byte aligned chunk start: vector loads
        for (int g = tid; g * 8 < span; g += 256) {
            float v[9];
            pcm_load9(pcm + x0 + 8 * g, v);          // reads x[8 g + 8] at most: inside the utterance (frame layout)
            float pe[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pe[i] = __fsub_rn(v[i + 1], __fmul_rn(preemph, v[i]));
            *reinterpret_cast<float4*>(s_pe + 8 * g) = make_float4(pe[0], pe[1], pe[2], pe[3]);
            *reinterpret_cast<float4*>(s_pe + 8 * g + 4) = make_float4(pe[4], pe[5], pe[6], pe[7]);
        }
    } else {
        for (int i = tid; i < span; i += 256)
            s_pe[i] = __fsub_rn(pcm_sample(pcm, x0 + i + 1), __fmul_rn(preemph, pcm_sample(pcm, x0 + i)));
    }
    __syncthreads();

    // mels of this lane: rounds 0, 1 take mel lane / lane + 32; the 16 widest bands (mels 64..79, up to ~30 bins) are
    // split in two halves over lane pairs (round 2: mel 64 + lane / 2, half lane & 1) and summed by one shuffle
    int m_st[3], m_n[3], m_w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int m = r < 2 ? lane + 32 * r : 64 + (lane >> 1);
        const int st = mel_start[m], n = mel_len[m];
        const int n0 = (n + 1) >> 1;
        const bool second = r == 2 && (lane & 1);
        m_st[r] = st + (second ? n0 : 0);
        m_n[r] = r < 2 ? n : (second ? n - n0 : n0);
        m_w[r] = m * mel_maxw + (second ? n0 : 0);
    }
    // Warp buffer of 512 complex points, point i at i + (i >> 4): one pad per 16 keeps the contiguous 8-byte accesses
    // of a half-warp inside one 128-byte bank window and spreads the strided stores of passes 1 and 2 (ncu r02d: with a
    // pad per 8 every access of the transform took twice its ideal wavefronts).  Every access below is (a
    // lane-dependent base) + (a compile-time function of i):
    //   reads of passes 2, 3      point j + 64 i            -> rd[b] + 68 i
    //   stores of pass 1 (Ns = 1)  point 8 j + i             -> 8 j + (j >> 1) + i
    //   stores of pass 2 (Ns = 8)  point 64 g + k + 8 i      -> 68 g + k + 8 i + (i >> 1)     (j = 8 g + k)
    //   stores of pass 3 (Ns = 64) point j + 64 i            -> rd[b] + 68 i
    float2* buf = s_buf + warp * kLmBuf;
    float2* rd[2];
    float2* w1[2];
    float2* w2[2];
    int tws2[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int j = lane + 32 * b;
        rd[b] = buf + j + (j >> 4);
        w1[b] = buf + 8 * j + (j >> 1);
        w2[b] = buf + 68 * (j >> 3) + (j & 7);
        tws2[b] = (j & 7) * 8;
    }
    // split: Z[k] at zk + 34 mm, Z[512 - k] at zc - 34 mm for k = lane + 32 mm (k = 0 pairs with itself)
    const float2* zk = buf + lane + (lane >> 4);
    const float2* zc = buf + (512 - lane) + ((512 - lane) >> 4);

    for (int pr = warp; 2 * pr < nf; pr += 8) {
        const int ta = 2 * pr;                                          // frames ta, ta + 1 of the chunk
        const bool has_b = ta + 1 < nf;
        const float* pa = s_pe + ta * kHop;
        const float* pb = pa + (has_b ? kHop : 0);
        // ---- pass 1 (Ns = 1): inputs straight from the samples, no twiddles ----------------------------------
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lane + 32 * b;
            float2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = j + 64 * i;
                v[i] = make_float2(0.f, 0.f);
                if (n >= kWinOff && n < kWinOff + kWin) {
                    const float w = s_win[n - kWinOff];
                    v[i] = make_float2(pa[n] * w, has_b ? pb[n] * w : 0.f);
                }
            }
            fft8(v);
#pragma unroll
            for (int i = 0; i < 8; ++i) w1[b][i] = v[i];
        }
        __syncwarp();
        // ---- passes 2, 3 (Ns = 8, 64): in place - all 16 points of a lane are loaded before any is stored --------
#pragma unroll
        for (int pass = 1; pass < 3; ++pass) {
            float2 v[2][8];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[b][i] = rd[b][68 * i];
            }
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                // twiddles exp(-2 pi i * i k / (8 Ns)) = tw512[i * k * 64 / Ns], k = j mod Ns
                const int tws = pass == 1 ? tws2[b] : lane + 32 * b;
#pragma unroll
                for (int i = 1; i < 8; ++i) v[b][i] = cmul(v[b][i], s_tw[i * tws]);
                fft8(v[b]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (pass == 1) w2[b][8 * i + (i >> 1)] = v[b][i];
                    else rd[b][68 * i] = v[b][i];
                }
            }
            __syncwarp();
        }
        // ---- split into the two real spectra, power --------------------------------------------------------
        float pw_a[9], pw_b[9];
#pragma unroll
        for (int mm = 0; mm < 9; ++mm) {
            const int k = lane + 32 * mm;
            pw_a[mm] = 0.f; pw_b[mm] = 0.f;
            if (k <= 256) {
                const float2 z = zk[34 * mm];
                const float2 c = k == 0 ? z : zc[-34 * mm];
                const float ar = 0.5f * (z.x + c.x), ai = 0.5f * (z.y - c.y);
                const float br = 0.5f * (z.y + c.y), bi = 0.5f * (c.x - z.x);
                pw_a[mm] = ar * ar + ai * ai;
                pw_b[mm] = br * br + bi * bi;
            }
        }
        __syncwarp();
        float* power = reinterpret_cast<float*>(buf);                   // [2][264]
#pragma unroll
        for (int mm = 0; mm < 9; ++mm) {
            const int k = lane + 32 * mm;
            if (k <= 256) { power[k] = pw_a[mm]; power[264 + k] = pw_b[mm]; }
        }
        __syncwarp();
        // ---- mel filterbank, log ------------------------------------------------------------------------------
        float* out_a = mel_out + (size_t)(f0 + t0 + ta) * kMel;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* w = mel_w + m_w[r];
            const float* p0 = power + m_st[r];
            float acc_a = 0.f, acc_b = 0.f;
            for (int i = 0; i < m_n[r]; ++i) {
                const float wi = __ldg(w + i);
                acc_a = fmaf(p0[i], wi, acc_a);
                acc_b = fmaf(p0[264 + i], wi, acc_b);
            }
            int m = lane + 32 * r;
            bool writer = true;
            if (r == 2) {
                acc_a += __shfl_xor_sync(0xffffffffu, acc_a, 1);
                acc_b += __shfl_xor_sync(0xffffffffu, acc_b, 1);
                m = 64 + (lane >> 1);
                writer = (lane & 1) == 0;
            }
            if (acc_a == 0.f) acc_a = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
            if (acc_b == 0.f) acc_b = 1.1920928955078125e-07f;
            if (writer) {
                out_a[m] = logf(acc_a);
                if (has_b) out_a[kMel + m] = logf(acc_b);
            }
        }
        __syncwarp();                                                   // `power` is read before the next pair's pass 1
    }
}

// ---------------------------------------------------------------------------------------------
struct Taps { float w[27]; };

constexpr int kSubRows = 32;                       // stacked rows per staged sub-chunk
constexpr int kSubFrames = 3 * kSubRows + 8;       // log-mel frames they touch (9 taps, stride 3)
constexpr int kStatChunks = 4;                     // partial CMVN sums per utterance

// Stage log-mel frames [3 g0 - 4, 3 g0 - 4 + kSubFrames) of one utterance in shared memory (zero outside
// [0, T): the correlation's zero padding, data.py:157-162) and hand every (row g, column group c)
// value of rows [g0, g1) to `sink(g, c, value)`; thread (x = j*80 + m, ty) owns column c*240 + x of
// rows g = g0 + ty, g0 + ty + 4, ...  Each mel value is read from global memory once per sub-chunk
// (coalesced float4) instead of once per tap.
template <typename Sink>
__device__ __forceinline__ void delta_rows(const float* __restrict__ mel_u, int T, int g0, int g1,
                                           const Taps& taps, float* s_mel, Sink&& sink) {
    const int x = threadIdx.x, ty = threadIdx.y;
    const int j = x / kMel, m = x - j * kMel;
    const int tid = ty * 240 + x;
    const int tb = 3 * g0 - 4;
    __syncthreads();                                // previous sub-chunk fully consumed
    for (int i = tid; i < kSubFrames * (kMel / 4); i += 960) {
        const int fr = i / (kMel / 4), q = i - fr * (kMel / 4);
        const int tt = tb + fr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tt >= 0 && tt < T) v = __ldg(reinterpret_cast<const float4*>(mel_u + (size_t)tt * kMel) + q);
        reinterpret_cast<float4*>(s_mel)[i] = v;
    }
    __syncthreads();
    for (int g = g0 + ty; g < g1; g += 4) {
        const float* base = s_mel + (3 * (g - g0) + j) * kMel + m;     // frame t - 4 of this row / slot
        // c = 0: identity tap (i = 4); c = 1: taps 2..6; c = 2: taps 0..8  (zero weights skipped as in
        // the reference's dense 9-tap kernels they are exact zeros)
        sink(g, 0, taps.w[4] * base[4 * kMel]);
        float a1 = 0.f;
#pragma unroll
        for (int i = 2; i < 7; ++i) a1 = fmaf(taps.w[9 + i], base[i * kMel], a1);
        sink(g, 1, a1);
        float a2 = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) a2 = fmaf(taps.w[18 + i], base[i * kMel], a2);
        sink(g, 2, a2);
    }
}

// pass 1: partial sums of every feature column over a quarter of the utterance's rows
__global__ void __launch_bounds__(960)
feat_stats_kernel(const float* __restrict__ mel, const int* __restrict__ frame_off,
                  const int* __restrict__ featrow_off, Taps taps, double2* __restrict__ partial) {
    __shared__ __align__(16) float s_mel[kSubFrames * kMel];
    __shared__ double s_red[2][4][240];
    const int u = blockIdx.x, ch = blockIdx.y;
    const int x = threadIdx.x, ty = threadIdx.y;
    const int f0 = frame_off[u];
    const int T = frame_off[u + 1] - f0;
    const int L = featrow_off[u + 1] - featrow_off[u];
    const int per = (L + kStatChunks - 1) / kStatChunks;
    const int ga = min(L, ch * per), gb = min(L, ga + per);
    double sum[3] = {0.0, 0.0, 0.0}, sq[3] = {0.0, 0.0, 0.0};
    for (int g0 = ga; g0 < gb; g0 += kSubRows) {
        delta_rows(mel + (size_t)f0 * kMel, T, g0, min(gb, g0 + kSubRows), taps, s_mel,
                   [&](int, int c, float v) { sum[c] += (double)v; sq[c] += (double)v * (double)v; });
    }
    for (int c = 0; c < 3; ++c) {
        __syncthreads();
        s_red[0][ty][x] = sum[c];
        s_red[1][ty][x] = sq[c];
        __syncthreads();
        if (ty == 0) {
            const double a = (s_red[0][0][x] + s_red[0][1][x]) + (s_red[0][2][x] + s_red[0][3][x]);
            const double b = (s_red[1][0][x] + s_red[1][1][x]) + (s_red[1][2][x] + s_red[1][3][x]);
            partial[((size_t)u * kStatChunks + ch) * kFeat + c * 240 + x] = make_double2(a, b);
        }
    }
}

// pass 2: recompute the taps from the staged log-mel rows, normalise, write the [L, 720] rows once
__global__ void __launch_bounds__(960)
feat_write_kernel(const float* __restrict__ mel, const int* __restrict__ frame_off,
                  const int* __restrict__ featrow_off, Taps taps, int normalise, float eps,
                  const double2* __restrict__ partial, const int* __restrict__ out_rowmap,
                  float* __restrict__ out, hi_t* __restrict__ split_hi, uint32_t* __restrict__ split_x) {
    __shared__ __align__(16) float s_mel[kSubFrames * kMel];
    __shared__ float s_mean[kFeat], s_den[kFeat];
    const int u = blockIdx.x;
    const int x = threadIdx.x, ty = threadIdx.y;
    const int f0 = frame_off[u];
    const int T = frame_off[u + 1] - f0;
    const int r0 = featrow_off[u];
    const int L = featrow_off[u + 1] - r0;
    const int g0 = blockIdx.y * kSubRows;
    if (g0 >= L) return;
    if (normalise && ty < 3) {
        // (x - mean) / (std + eps) per column with the unbiased std (eps 1e-6 main.py:37, 1e-7 data.py:517)
        const int col = ty * 240 + x;
        double tsum = 0.0, tsq = 0.0;
        for (int ch = 0; ch < kStatChunks; ++ch) {
            const double2 p = partial[((size_t)u * kStatChunks + ch) * kFeat + col];
            tsum += p.x;
            tsq += p.y;
        }
        const double mean = tsum / (double)L;
        double var = (tsq - tsum * mean) / (double)(L - 1);
        if (var < 0.0) var = 0.0;
        s_mean[col] = (float)mean;
        s_den[col] = (float)sqrt(var) + eps;
    }
    // (delta_rows starts with a barrier: s_mean / s_den are visible before the first sink call)
    delta_rows(mel + (size_t)f0 * kMel, T, g0, min(L, g0 + kSubRows), taps, s_mel,
               [&](int g, int c, float v) {
                   const int col = c * 240 + x;
                   const int row = out_rowmap ? out_rowmap[r0 + g] : (r0 + g);
                   const float val = normalise ? (v - s_mean[col]) / s_den[col] : v;
                   if (out) out[(size_t)row * kFeat + col] = val;
                   if (split_hi) {
                       // the A operand of the layer-0 input GEMM, already split (gemm_tc.cu kSplitAct): fp16 hi and,
                       // per 8 values, 8 x bf16(x - hi) then 8 x bf16(x) - no separate split pass over the features
                       const float hi = hi_part(val);
                       split_hi[(size_t)row * kFeat + col] = __float2half_rn(hi);
                       __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(split_x + (size_t)row * kFeat + (col & ~7));
                       xb[col & 7] = __float2bfloat16_rn(val - hi);
                       xb[8 + (col & 7)] = __float2bfloat16_rn(val);
                   }
               });
}

// ---------------------------------------------------------------------------------------------
// AudioLoader.batch_audio (data.py:513-518): instance normalisation of features that already exist
// ([sum L, 720], utterance-major), eps = 1e-7.  Same two-pass shape as above: fixed-order double
// partial sums per quarter of the rows (deterministic), then one read + one write of every row.
__global__ void __launch_bounds__(720)
cmvn_stats_kernel(const float* __restrict__ x, const int* __restrict__ featrow_off,
                  double2* __restrict__ partial) {
    const int u = blockIdx.x, ch = blockIdx.y, col = threadIdx.x;
    const int r0 = featrow_off[u];
    const int L = featrow_off[u + 1] - r0;
    const int per = (L + kStatChunks - 1) / kStatChunks;
    const int ga = min(L, ch * per), gb = min(L, ga + per);
    double s = 0.0, q = 0.0;
    const float* p = x + (size_t)(r0 + ga) * kFeat + col;
    int g = ga;
    for (; g + 4 <= gb; g += 4, p += 4 * kFeat) {       // 4 independent loads in flight per thread
        const float a = p[0], b = p[kFeat], c = p[2 * kFeat], d = p[3 * kFeat];
        s += (double)a; q += (double)a * a;
        s += (double)b; q += (double)b * b;
        s += (double)c; q += (double)c * c;
        s += (double)d; q += (double)d * d;
    }
    for (; g < gb; ++g, p += kFeat) { const float a = p[0]; s += (double)a; q += (double)a * a; }
    partial[((size_t)u * kStatChunks + ch) * kFeat + col] = make_double2(s, q);
}

__global__ void __launch_bounds__(720)
cmvn_apply_kernel(const float* __restrict__ x, const int* __restrict__ featrow_off, float eps,
                  const double2* __restrict__ partial, float* __restrict__ out) {
    const int u = blockIdx.x, col = threadIdx.x;
    const int r0 = featrow_off[u];
    const int L = featrow_off[u + 1] - r0;
    const int g0 = blockIdx.y * kSubRows;
    if (g0 >= L) return;
    double tsum = 0.0, tsq = 0.0;
    for (int ch = 0; ch < kStatChunks; ++ch) {
        const double2 pr = partial[((size_t)u * kStatChunks + ch) * kFeat + col];
        tsum += pr.x;
        tsq += pr.y;
    }
    const double mean_d = tsum / (double)L;
    double var = (tsq - tsum * mean_d) / (double)(L - 1);      // L == 1 -> NaN, like torch.std
    if (var < 0.0) var = 0.0;
    const float mean = (float)mean_d, den = (float)sqrt(var) + eps;
    const int g1 = min(L, g0 + kSubRows);
    for (int g = g0; g < g1; ++g) {
        const size_t i = (size_t)(r0 + g) * kFeat + col;
        out[i] = (x[i] - mean) / den;
    }
}

int launch_cmvn(asr_handle* h, const float* d_in, const int* d_featrow_off, int B, int max_rows_per_utt,
                float eps, float* d_out, cudaStream_t st) {
    cmvn_stats_kernel<<<dim3(B, kStatChunks), kFeat, 0, st>>>(d_in, d_featrow_off,
                                                              reinterpret_cast<double2*>(h->ws.feat_partial));
    ASR_CHECK_LAUNCH();
    cmvn_apply_kernel<<<dim3(B, (max_rows_per_utt + kSubRows - 1) / kSubRows), kFeat, 0, st>>>(
        d_in, d_featrow_off, eps, reinterpret_cast<const double2*>(h->ws.feat_partial), d_out);
    ASR_CHECK_LAUNCH();
    h->launches += 2;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int upload(asr_handle* h, T** dst, const std::vector<T>& v) {
    ASR_CUDA(cudaMalloc(dst, v.size() * sizeof(T)));
    h->weight_allocs.push_back(*dst);
    ASR_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ASR_OK;
}

int build_feature_consts(asr_handle* h, const asr_feature_consts* fc) {
    FeatureConsts& c = h->fc;
    c.preemph = fc->preemphasis;
    for (int i = 0; i < 27; ++i) c.taps[i] = fc->taps[i];
    std::vector<float> win(fc->window, fc->window + kWin);
    ASR_TRY(upload(h, &c.window, win));
    const double pi = 3.14159265358979323846;
    std::vector<float2> t512(512);
    for (int k = 0; k < 512; ++k)
        t512[k] = make_float2((float)cos(-2.0 * pi * k / 512.0), (float)sin(-2.0 * pi * k / 512.0));
    ASR_TRY(upload(h, &c.tw512, t512));
    // CSR-like band storage of the [257, 80] filterbank
    std::vector<int> start(kMel, 0), len(kMel, 0);
    int maxw = 1;
    for (int m = 0; m < kMel; ++m) {
        int first = -1, last = -1;
        for (int j = 0; j < kBins; ++j)
            if (fc->mel_fb[j * kMel + m] != 0.f) { if (first < 0) first = j; last = j; }
        if (first >= 0) { start[m] = first; len[m] = last - first + 1; }
        if (len[m] > maxw) maxw = len[m];
    }
    std::vector<float> w((size_t)kMel * maxw, 0.f);
    for (int m = 0; m < kMel; ++m)
        for (int i = 0; i < len[m]; ++i) w[(size_t)m * maxw + i] = fc->mel_fb[(start[m] + i) * kMel + m];
    c.mel_maxw = maxw;
    ASR_TRY(upload(h, &c.mel_start, start));
    ASR_TRY(upload(h, &c.mel_len, len));
    ASR_TRY(upload(h, &c.mel_w, w));
    return ASR_OK;
}

int launch_logmel(asr_handle* h, const void* d_pcm, int format, const long long* d_pcm_off,
                  const int* d_frame_off, int B, int max_frames_per_utt, float* d_mel, cudaStream_t st) {
    if (max_frames_per_utt <= 0 || B <= 0) return ASR_OK;
    const FeatureConsts& c = h->fc;
    const dim3 grid((max_frames_per_utt + kLmFrames - 1) / kLmFrames, B);
    if (format == ASR_PCM_S16) {
        ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&logmel_kernel<short>), kLmSmem));
        logmel_kernel<short><<<grid, 256, kLmSmem, st>>>(static_cast<const short*>(d_pcm), d_pcm_off, d_frame_off, c.window,
                                                         c.tw512, c.mel_start, c.mel_len, c.mel_w, c.mel_maxw, c.preemph,
                                                         d_mel);
    } else {
        ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&logmel_kernel<float>), kLmSmem));
        logmel_kernel<float><<<grid, 256, kLmSmem, st>>>(static_cast<const float*>(d_pcm), d_pcm_off, d_frame_off, c.window,
                                                         c.tw512, c.mel_start, c.mel_len, c.mel_w, c.mel_maxw, c.preemph,
                                                         d_mel);
    }
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

int launch_delta_cmvn(asr_handle* h, const float* d_mel, const int* d_frame_off,
                      const int* d_featrow_off, int B, int max_rows_per_utt, int normalise, float eps,
                      const int* out_rowmap, float* d_out, cudaStream_t st, hi_t* split_hi, float* split_x) {
    Taps t;
    for (int i = 0; i < 27; ++i) t.w[i] = h->fc.taps[i];
    dim3 block(240, 4);
    if (normalise) {
        feat_stats_kernel<<<dim3(B, kStatChunks), block, 0, st>>>(d_mel, d_frame_off, d_featrow_off, t,
                                                                  reinterpret_cast<double2*>(h->ws.feat_partial));
        ASR_CHECK_LAUNCH();
        h->launches++;
    }
    feat_write_kernel<<<dim3(B, (max_rows_per_utt + kSubRows - 1) / kSubRows), block, 0, st>>>(
        d_mel, d_frame_off, d_featrow_off, t, normalise, eps, reinterpret_cast<const double2*>(h->ws.feat_partial),
        out_rowmap, d_out, split_hi, reinterpret_cast<uint32_t*>(split_x));
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
