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
#include <math.h>

#include <vector>

#include "asr_internal.cuh"

namespace asr {

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

constexpr int kFramesPerCta = 8;

__global__ void __launch_bounds__(256)
logmel_kernel(const float* __restrict__ pcm, const long long* __restrict__ pcm_off,
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

    const int gf = blockIdx.x * kFramesPerCta + warp;
    if (gf >= total_frames) return;

    // utterance of this frame: largest u with frame_off[u] <= gf
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (frame_off[mid] <= gf) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int t = gf - frame_off[u];
    const float* x = pcm + pcm_off[u] + (long long)t * kHop;

    float2* a = s_a[warp];
    float2* b = s_b[warp];

    // windowed, pre-emphasised frame packed as z[n] = xw[2n] + i xw[2n+1]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = lane + 32 * i;
        const int s = 2 * n;
        float2 z = make_float2(0.f, 0.f);
        if (s >= kWinOff && s < kWinOff + kWin) {   // kWinOff, kWin even -> both samples inside
            const float x0 = x[s], x1 = x[s + 1], x2 = x[s + 2];
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

// ---------------------------------------------------------------------------------------------
struct Taps { float w[27]; };

__global__ void __launch_bounds__(960)
delta_cmvn_kernel(const float* __restrict__ mel, const int* __restrict__ frame_off,
                  const int* __restrict__ featrow_off, Taps taps, int normalise,
                  const int* __restrict__ out_rowmap, float* __restrict__ out) {
    __shared__ double s_sum[4][240];
    __shared__ double s_sq[4][240];
    const int u = blockIdx.x;
    const int c = blockIdx.y;
    const int x = threadIdx.x;          // 0..239 = j*80 + m
    const int ty = threadIdx.y;         // 0..3
    const int j = x / kMel, m = x - j * kMel;
    const int f0 = frame_off[u];
    const int T = frame_off[u + 1] - f0;
    const int r0 = featrow_off[u];
    const int L = featrow_off[u + 1] - r0;
    const int col = c * 240 + x;
    const float* tw = taps.w + c * 9;
    const int i_lo = (c == 0) ? 4 : (c == 1 ? 2 : 0);
    const int i_hi = (c == 0) ? 5 : (c == 1 ? 7 : 9);

    // the 9-tap correlation is recomputed in the second pass (its log-mel input is L2-resident)
    // instead of writing raw features and reading them back: the output is written exactly once
    auto tap = [&](int g) {
        const int t = 3 * g + j;
        float acc = 0.f;
        for (int i = i_lo; i < i_hi; ++i) {
            const int tt = t + i - 4;
            if (tt >= 0 && tt < T) acc = fmaf(tw[i], __ldg(mel + (size_t)(f0 + tt) * kMel + m), acc);
        }
        return acc;
    };
    if (!normalise) {
        for (int g = ty; g < L; g += 4) {
            const int row = out_rowmap ? out_rowmap[r0 + g] : (r0 + g);
            out[(size_t)row * kFeat + col] = tap(g);
        }
        return;
    }
    double sum = 0.0, sq = 0.0;
    for (int g = ty; g < L; g += 4) {
        const float acc = tap(g);
        sum += (double)acc;
        sq += (double)acc * (double)acc;
    }
    s_sum[ty][x] = sum;
    s_sq[ty][x] = sq;
    __syncthreads();
    const double tsum = s_sum[0][x] + s_sum[1][x] + s_sum[2][x] + s_sum[3][x];
    const double tsq = s_sq[0][x] + s_sq[1][x] + s_sq[2][x] + s_sq[3][x];
    const double mean = tsum / (double)L;
    double var = (tsq - tsum * mean) / (double)(L - 1);     // unbiased, torch.std default
    if (var < 0.0) var = 0.0;
    const float meanf = (float)mean;
    const float denom = (float)sqrt(var) + 1e-6f;
    for (int g = ty; g < L; g += 4) {
        const int row = out_rowmap ? out_rowmap[r0 + g] : (r0 + g);
        out[(size_t)row * kFeat + col] = (tap(g) - meanf) / denom;
    }
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

int launch_logmel(asr_handle* h, const float* d_pcm, const long long* d_pcm_off,
                  const int* d_frame_off, int B, int total_frames, float* d_mel, cudaStream_t st) {
    if (total_frames <= 0) return ASR_OK;
    const FeatureConsts& c = h->fc;
    const int grid = (total_frames + kFramesPerCta - 1) / kFramesPerCta;
    logmel_kernel<<<grid, 256, 0, st>>>(d_pcm, d_pcm_off, d_frame_off, B, total_frames, c.window,
                                        c.tw256, c.tw512, c.mel_start, c.mel_len, c.mel_w,
                                        c.mel_maxw, c.preemph, d_mel);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

int launch_delta_cmvn(asr_handle* h, const float* d_mel, const int* d_frame_off,
                      const int* d_featrow_off, int B, int normalise, const int* out_rowmap,
                      float* d_out, cudaStream_t st) {
    Taps t;
    for (int i = 0; i < 27; ++i) t.w[i] = h->fc.taps[i];
    dim3 grid(B, 3), block(240, 4);
    delta_cmvn_kernel<<<grid, block, 0, st>>>(d_mel, d_frame_off, d_featrow_off, t, normalise,
                                              out_rowmap, d_out);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
