// Row movers of the encoder stage: packing caller-supplied features into the packed time-major order the
// encoder works in (pack_sequence, encoder.py:47-53), and exporting results in the reference's padded layouts
// (pad_packed_sequence + un-sort, encoder.py:63-72).  The recurrence itself is in encoder_tc3.cu.
#include "asr_internal.cuh"

namespace asr {

__global__ void pack_rows_kernel(const float4* __restrict__ src, const int* __restrict__ rowmap,
                                 long long rows, int w4, float4* __restrict__ dst) {
    const long long total = rows * w4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w4;
        const int c = (int)(i - r * w4);
        dst[i] = src[(long long)rowmap[r] * w4 + c];
    }
}

int launch_pack_rows(asr_handle* h, const float* src, const int* rowmap, int64_t rows, int width,
                     float* dst, cudaStream_t st) {
    if (rows <= 0) return ASR_OK;
    const int w4 = width / 4;
    long long total = rows * w4;
    int grid = (int)((total + 255) / 256);
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    pack_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(src), rowmap, rows, w4,
                                           reinterpret_cast<float4*>(dst));
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// utterance-major (sorted order) -> reference layout [Lmax, B, W] in ORIGINAL utterance order
__global__ void export_padded_kernel(const float* __restrict__ src, const int* __restrict__ uoff,
                                     const int* __restrict__ len, const int* __restrict__ order,
                                     int W, int B, const float* __restrict__ pad_row,
                                     float* __restrict__ dst) {
    const int t = blockIdx.x, r = blockIdx.y;
    float* d = dst + ((size_t)t * B + order[r]) * W;
    if (t < len[r]) {
        const float* s = src + (size_t)(uoff[r] + t) * W;
        for (int c = threadIdx.x; c < W; c += blockDim.x) d[c] = s[c];
    } else {
        for (int c = threadIdx.x; c < W; c += blockDim.x) d[c] = pad_row ? pad_row[c] : 0.f;
    }
}

int launch_export_padded(asr_handle* h, const float* src_utt, int width, float* dst, int Lmax, int B,
                         const float* pad_row, cudaStream_t st) {
    const BatchMeta& m = h->meta;
    dim3 grid(Lmax, B);
    export_padded_kernel<<<grid, 128, 0, st>>>(src_utt, m.d_uoff_sorted, m.d_len_sorted, m.d_order,
                                               width, B, pad_row, dst);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

__global__ void export_packed_padded_kernel(const float* __restrict__ src,
                                            const int* __restrict__ toff,
                                            const int* __restrict__ len,
                                            const int* __restrict__ order, int W, int B,
                                            float* __restrict__ dst) {
    const int t = blockIdx.x, r = blockIdx.y;
    float* d = dst + ((size_t)t * B + order[r]) * W;
    if (t < len[r]) {
        const float* s = src + (size_t)(toff[t] + r) * W;
        for (int c = threadIdx.x; c < W; c += blockDim.x) d[c] = s[c];
    } else {
        for (int c = threadIdx.x; c < W; c += blockDim.x) d[c] = 0.f;
    }
}

int launch_export_packed_padded(asr_handle* h, const float* src_packed, int width, float* dst,
                                cudaStream_t st) {
    const BatchMeta& m = h->meta;
    dim3 grid(m.Lmax, m.B);
    export_packed_padded_kernel<<<grid, 128, 0, st>>>(src_packed, m.d_toff, m.d_len_sorted,
                                                      m.d_order, width, m.B, dst);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

__global__ void unsort_rows_kernel(const float* __restrict__ src, const int* __restrict__ order,
                                   int W, float* __restrict__ dst) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < W; c += blockDim.x)
        dst[(size_t)order[r] * W + c] = src[(size_t)r * W + c];
}

int launch_unsort_rows(asr_handle* h, const float* src, int width, float* dst, cudaStream_t st) {
    unsort_rows_kernel<<<h->meta.B, 128, 0, st>>>(src, h->meta.d_order, width, dst);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
