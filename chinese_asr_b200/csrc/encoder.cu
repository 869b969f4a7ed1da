// Encoder recurrence: 4 residual bidirectional LSTM layers on length-sorted, packed sequences.
// Replaces RNNEncoder.forward / RNN_RES.forward (encoder.py:36-81, util.py:1223-1324), i.e. the
// cuDNN/ATen nn.LSTM time loop, pack_sequence / pad_packed_sequence and the residual add.
//
// Layout: activations live in PACKED TIME-MAJOR order (row(t, r) = toff[t] + r, r = rank in the
// descending-length order), the same order nn.utils.rnn.pack_sequence produces, so the per-step
// rows of all active sequences are contiguous.  Padding never exists in memory.
//
// lstm_rec_kernel: one thread-block CLUSTER of 8 CTAs owns (direction, chunk of <= 8*P sequences).
//   * W_hh of one direction is 1024x256 fp32 = 1 MB: it is split over the 8 CTAs (CTA j owns the
//     4 gates of hidden units [32j, 32j+32)) and held STATIONARY IN REGISTERS for the whole
//     sequence (128 floats per thread) - it is read from HBM/L2 exactly once per layer.
//   * each step: h_{t-1} (rows x 256) is read from local shared memory (broadcast LDS.128),
//     the 2 K-halves are reduced through shared memory, the gate non-linearities + cell update
//     run in the same kernel (c stays in a register), and the new 32-wide slice of h is written
//     into the shared memory of all 8 CTAs through DSMEM, followed by one cluster barrier.
//   * x-projection pre-activations (xg, from the input GEMM) are prefetched at the top of the
//     step so their HBM/L2 latency overlaps the recurrent mat-vec.
//   * residual add (util.py:1284-1291) and the final (h_n, c_n) extraction (encoder.py:67-72)
//     are fused into the store epilogue; the last layer writes utterance-major memory for the
//     attention kernels.
// This kernel is latency-bound by construction (4 * Lmax dependent steps); see DESIGN.md.
#include <cooperative_groups.h>

#include "asr_internal.cuh"

namespace cg = cooperative_groups;

namespace asr {

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

struct RecParams {
    const float* xg;        // [rows, 2048] permuted gate pre-activations (bias included)
    const float* whh;       // [2, 1024, 256] permuted
    const float* x_in;      // residual input [rows, 512] packed, or nullptr
    float* y_packed;        // [rows, 512] packed, or nullptr
    float* y_utt;           // [rows, 512] utterance-major sorted, or nullptr
    float* h_fin;           // [B, 512]
    float* c_fin;           // [B, 512]
    const int* len_sorted;  // [B]
    const int* toff;        // [Lmax + 1]
    const int* uoff;        // [B + 1]
    int B;
    int rows_per_chunk;
    int nchunks;
};

template <int P>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1)
lstm_rec_kernel(RecParams p) {
    constexpr int RT = 8 * P;
    extern __shared__ __align__(16) float smem[];
    float* hbuf = smem;                        // [2][RT][256]
    float* red = smem + 2 * RT * 256;          // [2][RT][128]
    __shared__ int s_len[RT];

    cg::cluster_group cluster = cg::this_cluster();
    const int j = (int)cluster.block_rank();           // 0..7: hidden units [32j, 32j+32)
    const int cid = blockIdx.x / 8;
    const int dir = cid / p.nchunks;
    const int chunk = cid - dir * p.nchunks;
    const int tid = threadIdx.x;
    const int col = tid & 127;                          // local gate column: gate*32 + uu
    const int kh = tid >> 7;                            // K half

    const int r0 = chunk * p.rows_per_chunk;
    const int nrows = min(p.rows_per_chunk, p.B - r0);

    // stationary recurrent weights: row (j*128 + col) of the permuted [1024, 256] matrix
    float w[128];
    {
        const float4* src = reinterpret_cast<const float4*>(
            p.whh + ((size_t)dir * kGates + j * 128 + col) * kEncH + kh * 128);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float4 v = src[i];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
    for (int i = tid; i < 2 * RT * 256; i += 256) hbuf[i] = 0.f;
    if (tid < RT) s_len[tid] = tid < nrows ? p.len_sorted[r0 + tid] : 0;
    __syncthreads();
    const int Lc = s_len[0];

    // gate-phase ownership: pair q -> row i = (tid >> 5) + 8*q, unit uu = tid & 31
    const int uu = tid & 31;
    const int irow = tid >> 5;
    float c_reg[P], h_reg[P];
#pragma unroll
    for (int q = 0; q < P; ++q) { c_reg[q] = 0.f; h_reg[q] = 0.f; }

    cluster.sync();

    int cur = 0;
    for (int s = 0; s < Lc; ++s) {
        const int t = dir == 0 ? s : Lc - 1 - s;
        int nact = 0;
        for (int i = 0; i < nrows; ++i) nact += (s_len[i] > t) ? 1 : 0;
        const int row_t = p.toff[t] + r0;               // packed row of chunk row 0 at time t

        // prefetch this step's input-projection pre-activations (consumed after the mat-vec)
        float xi[P], xf[P], xgg[P], xo[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = irow + 8 * q;
            xi[q] = xf[q] = xgg[q] = xo[q] = 0.f;
            if (i < nact) {
                const float* g = p.xg + (size_t)(row_t + i) * (2 * kGates) + dir * kGates + j * 128 + uu;
                xi[q] = __ldg(g);
                xf[q] = __ldg(g + 32);
                xgg[q] = __ldg(g + 64);
                xo[q] = __ldg(g + 96);
            }
        }

        // recurrent mat-vec: red[kh][i][col] = sum_k w[k] * h[i][kh*128 + k]
        // packed fp32x2 FMAs (FFMA2, sm_100): two rows per iteration -> 4 independent chains
        const float* hc = hbuf + cur * RT * 256 + kh * 128;
        for (int i = 0; i < nact; i += 2) {
            const int i1 = min(i + 1, RT - 1);
            const float4* hv0 = reinterpret_cast<const float4*>(hc + i * 256);
            const float4* hv1 = reinterpret_cast<const float4*>(hc + i1 * 256);
            float2 a01 = make_float2(0.f, 0.f), a23 = a01, b01 = a01, b23 = a01;
#pragma unroll
            for (int k4 = 0; k4 < 32; ++k4) {
                const float4 v = hv0[k4];
                const float4 u = hv1[k4];
                const float2 w01 = make_float2(w[4 * k4], w[4 * k4 + 1]);
                const float2 w23 = make_float2(w[4 * k4 + 2], w[4 * k4 + 3]);
                a01 = __ffma2_rn(w01, make_float2(v.x, v.y), a01);
                a23 = __ffma2_rn(w23, make_float2(v.z, v.w), a23);
                b01 = __ffma2_rn(w01, make_float2(u.x, u.y), b01);
                b23 = __ffma2_rn(w23, make_float2(u.z, u.w), b23);
            }
            red[(kh * RT + i) * 128 + col] = (a01.x + a01.y) + (a23.x + a23.y);
            if (i + 1 < nact) red[(kh * RT + i + 1) * 128 + col] = (b01.x + b01.y) + (b23.x + b23.y);
        }
        __syncthreads();

        // gate non-linearities + cell update for (row i, unit 32j+uu); broadcast h through DSMEM
        const int nxt = cur ^ 1;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int i = irow + 8 * q;
            if (i < nact) {
                const float* ra = red + (0 * RT + i) * 128 + uu;
                const float* rb = red + (1 * RT + i) * 128 + uu;
                const float gi = xi[q] + (ra[0] + rb[0]);
                const float gf = xf[q] + (ra[32] + rb[32]);
                const float gg = xgg[q] + (ra[64] + rb[64]);
                const float go = xo[q] + (ra[96] + rb[96]);
                const float c = sigmoid_acc(gf) * c_reg[q] + sigmoid_acc(gi) * tanhf(gg);
                const float hh = sigmoid_acc(go) * tanhf(c);
                c_reg[q] = c;
                h_reg[q] = hh;
                float* slot = hbuf + nxt * RT * 256 + i * 256 + 32 * j + uu;
#pragma unroll
                for (int rk = 0; rk < 8; ++rk) *cluster.map_shared_rank(slot, rk) = hh;
                const size_t row = (size_t)(row_t + i);
                const int ocol = dir * kEncH + 32 * j + uu;
                float y = hh;
                if (p.x_in) y += p.x_in[row * kEnc + ocol];
                if (p.y_packed) p.y_packed[row * kEnc + ocol] = y;
                if (p.y_utt) p.y_utt[(size_t)(p.uoff[r0 + i] + t) * kEnc + ocol] = y;
            }
        }
        cluster.sync();
        cur = nxt;
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int i = irow + 8 * q;
        if (i < nrows) {
            const int ocol = dir * kEncH + 32 * j + uu;
            p.h_fin[(size_t)(r0 + i) * kEnc + ocol] = h_reg[q];
            p.c_fin[(size_t)(r0 + i) * kEnc + ocol] = c_reg[q];
        }
    }
}

int launch_lstm_recurrence(asr_handle* h, int layer, const float* xg, const float* x_in,
                           float* y_packed, float* y_utt, float* h_fin, float* c_fin,
                           cudaStream_t st) {
    const BatchMeta& m = h->meta;
    RecParams p{};
    p.xg = xg;
    p.whh = h->w.enc_w_hh[layer];
    p.x_in = x_in;
    p.y_packed = y_packed;
    p.y_utt = y_utt;
    p.h_fin = h_fin;
    p.c_fin = c_fin;
    p.len_sorted = m.d_len_sorted;
    p.toff = m.d_toff;
    p.uoff = m.d_uoff_sorted;
    p.B = m.B;
    // ~16 clusters of 8 CTAs fit on the 148 SMs (2 per GPC): aim at 8 chunks per direction
    int rpc = (m.B + 7) / 8;
    if (rpc < 1) rpc = 1;
    if (rpc > 16) rpc = 16;
    p.rows_per_chunk = rpc;
    p.nchunks = (m.B + rpc - 1) / rpc;
    const int grid = 2 * p.nchunks * 8;
    if (rpc <= 8) {
        const size_t smem = (size_t)8 * 3072;
        lstm_rec_kernel<1><<<grid, 256, smem, st>>>(p);
    } else {
        const size_t smem = (size_t)16 * 3072;
        static bool attr_set = false;
        if (!attr_set) {
            ASR_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        lstm_rec_kernel<2><<<grid, 256, smem, st>>>(p);
    }
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
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
