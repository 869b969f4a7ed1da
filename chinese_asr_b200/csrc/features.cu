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

constexpr int kFramesPerCta = 8;      // one warp per frame
constexpr int kFrameRounds = 4;       // frames per warp: the twiddle / window tables are staged once per 32 frames

// PCM sample as the float32 soundfile.read(dtype='float32') hands the reference (data.py:111): float32 files as
// stored, 16-bit files as x / 32768 (exact: a power of two), so both sample types give bit-identical frames.
__device__ __forceinline__ float pcm_sample(const float* x, int i) { return x[i]; }
__device__ __forceinline__ float pcm_sample(const short* x, int i) { return (float)x[i] * (1.0f / 32768.0f); }

template <typename S>
__global__ void __launch_bounds__(256)
logmel_kernel(const S* __restrict__ pcm, const long long* __restrict__ pcm_off,
              const int* __restrict__ frame_off, int B, int total_frames,
              const float* __restrict__ g_window, const float2* __restrict__ g_tw256,
              const float2* __restrict__ g_tw512, const int* __restrict__ mel_start,
              const int* __restrict__ mel_len, const float* __restrict__ mel_w, int mel_maxw,
              float preemph, float* __restrict__ mel_out) {
    __shared__ float2 s_tw256[256];
    __shared__ float2 s_tw512[257];
    __shared__ float s_win[kWin];
    __shared__ float2 s_a[kFramesPerCta][256];
    __shared__ float2 s_b[kFramesPerCta][256 + 2];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += 256) s_tw256[i] = g_tw256[i];
    for (int i = tid; i < 257; i += 256) s_tw512[i] = g_tw512[i];
    for (int i = tid; i < kWin; i += 256) s_win[i] = g_window[i];
    __syncthreads();

    for (int round = 0; round < kFrameRounds; ++round) {
    const int gf = (blockIdx.x * kFrameRounds + round) * kFramesPerCta + warp;
    if (gf >= total_frames) return;
    __syncwarp();                       // the previous frame's mel pass has finished reading `power`

    // utterance of this frame: largest u with frame_off[u] <= gf
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (frame_off[mid] <= gf) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int t = gf - frame_off[u];
    const S* x = pcm + pcm_off[u] + (long long)t * kHop;

    float2* a = s_a[warp];
    float2* b = s_b[warp];

    // windowed, pre-emphasised frame packed as z[n] = xw[2n] + i xw[2n+1]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = lane + 32 * i;
        const int s = 2 * n;
        float2 z = make_float2(0.f, 0.f);
        if (s >= kWinOff && s < kWinOff + kWin) {   // kWinOff, kWin even -> both samples inside
            const float x0 = pcm_sample(x, s), x1 = pcm_sample(x, s + 1), x2 = pcm_sample(x, s + 2);
            const float y0 = __fsub_rn(x1, __fmul_rn(preemph, x0));
            const float y1 = __fsub_rn(x2, __fmul_rn(preemph, x1));
            z.x = y0 * s_win[s - kWinOff];
            z.y = y1 * s_win[s + 1 - kWinOff];
        }
        a[n] = z;
    }
    __syncwarp();

    // 256-point complex FFT, radix-4 Stockham autosort, 4 passes (Ns = 1, 4, 16, 64)
    float2* src = a;
    float2* dst = b;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int Ns = 1 << (2 * pass);
        const int tw_stride = 64 >> (2 * pass);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int j = lane + 32 * i;
            const int k = j & (Ns - 1);
            float2 v0 = src[j];
            float2 v1 = src[j + 64];
            float2 v2 = src[j + 128];
            float2 v3 = src[j + 192];
            if (pass > 0) {
                v1 = cmul(v1, s_tw256[k * tw_stride]);
                v2 = cmul(v2, s_tw256[2 * k * tw_stride]);
                v3 = cmul(v3, s_tw256[3 * k * tw_stride]);
            }
            const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y);
            const float2 d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y);
            const float2 d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
            dst[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);      // v0 - i v1 - v2 + i v3
            dst[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
            dst[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);  // v0 + i v1 - v2 - i v3
        }
        __syncwarp();
        float2* tmp = src; src = dst; dst = tmp;
    }
    // result in `src` (== a after 4 passes); split into the 257 real-FFT bins, store power in dst
    float* power = reinterpret_cast<float*>(dst);
    for (int kbin = lane; kbin <= 256; kbin += 32) {
        const float2 zk = src[kbin & 255];
        const float2 zc = src[(256 - kbin) & 255];
        // E = (zk + conj(zc)) / 2 ; O = (zk - conj(zc)) / (2i)
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
        const float2 wo = cmul(s_tw512[kbin], o);
        const float re = e.x + wo.x, im = e.y + wo.y;
        power[kbin] = re * re + im * im;
    }
    __syncwarp();

    float* out = mel_out + (size_t)gf * kMel;
    for (int m = lane; m < kMel; m += 32) {
        const int st = mel_start[m], n = mel_len[m];
        const float* w = mel_w + m * mel_maxw;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(power[st + i], w[i], acc);
        if (acc == 0.f) acc = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
        out[m] = logf(acc);
    }
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
    std::vector<float2> t256(256), t512(257);
    for (int k = 0; k < 256; ++k)
        t256[k] = make_float2((float)cos(-2.0 * pi * k / 256.0), (float)sin(-2.0 * pi * k / 256.0));
    for (int k = 0; k <= 256; ++k)
        t512[k] = make_float2((float)cos(-2.0 * pi * k / 512.0), (float)sin(-2.0 * pi * k / 512.0));
    ASR_TRY(upload(h, &c.tw256, t256));
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
                  const int* d_frame_off, int B, int total_frames, float* d_mel, cudaStream_t st) {
    if (total_frames <= 0) return ASR_OK;
    const FeatureConsts& c = h->fc;
    const int per_cta = kFramesPerCta * kFrameRounds;
    const int grid = (total_frames + per_cta - 1) / per_cta;
    if (format == ASR_PCM_S16)
        logmel_kernel<short><<<grid, 256, 0, st>>>(static_cast<const short*>(d_pcm), d_pcm_off, d_frame_off, B,
                                                   total_frames, c.window, c.tw256, c.tw512, c.mel_start,
                                                   c.mel_len, c.mel_w, c.mel_maxw, c.preemph, d_mel);
    else
        logmel_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(d_pcm), d_pcm_off, d_frame_off, B,
                                                   total_frames, c.window, c.tw256, c.tw512, c.mel_start,
                                                   c.mel_len, c.mel_w, c.mel_maxw, c.preemph, d_mel);
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
