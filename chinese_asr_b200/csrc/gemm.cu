// fp32 GEMM C[M,N] = A[M,K] * W[N,K]^T with fused epilogues (bias / temperature / LSTM cell).
//
// Replaces the reference's cuBLAS/MKL calls: nn.LSTM input projections (util.py:1259), attention
// keys (attention.py:77), nn.LSTMCell (util.py:1650-1661 via decoder.py:114) and the vocabulary
// projection (decoder.py:133).  Both operands are K-major (activations [rows, K], nn.Linear / LSTM
// weights [out, in]), so no transposes are materialised.
//
// The A operand is described by up to three K-segments with optional row gathers (AOperand): the
// decoder cell reads [embedding[token] | ctx[src] | h[src]] in place, which is what removes the
// reference's per-step reorder-by-backpointer copies (model.py:913-925) and the k-fold tiling of
// the decoder state (model.py:660-669).
//
// This file is the CUDA-core fp32 path (exact fp32 FMA accumulation, needed for token-exact
// parity with the fp32 reference); the tcgen05 3xTF32 path lives in gemm_tc.cu.
#include "asr_internal.cuh"

namespace asr {

constexpr int BK = 16;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
sgemm_kernel(AOperand A, const float* __restrict__ W, int M, int N, int K, GemmEpilogue epi) {
    if (epi.stop_flag && *epi.stop_flag >= 0) return;
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    constexpr int A_SLOTS = (BM * BK / 4) / 256;   // float4 loads per thread per tile
    constexpr int B_SLOTS = (BN * BK / 4) / 256;
    static_assert(A_SLOTS >= 1 && B_SLOTS >= 1, "tile too small for 256 threads");
    static_assert((BN / TN) * (BM / TM) == 256, "thread layout");

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    // per-thread load slots
    const float* a_ptr[A_SLOTS][3];
    int a_row[A_SLOTS], a_kq[A_SLOTS];
    bool a_ok[A_SLOTS];
#pragma unroll
    for (int s = 0; s < A_SLOTS; ++s) {
        const int li = tid + s * 256;
        a_row[s] = li >> 2;
        a_kq[s] = li & 3;
        const int row = m0 + a_row[s];
        a_ok[s] = row < M;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            a_ptr[s][g] = nullptr;
            if (g < A.nseg && a_ok[s]) {
                const int r = A.seg[g].rowidx ? A.seg[g].rowidx[row] : row;
                const int kstart = g == 0 ? 0 : A.seg[g - 1].kend;
                a_ptr[s][g] = A.seg[g].base + (size_t)r * A.seg[g].ld - kstart;
            }
        }
    }
    const float* b_ptr[B_SLOTS];
    int b_row[B_SLOTS], b_kq[B_SLOTS];
#pragma unroll
    for (int s = 0; s < B_SLOTS; ++s) {
        const int li = tid + s * 256;
        b_row[s] = li >> 2;
        b_kq[s] = li & 3;
        const int n = n0 + b_row[s];
        b_ptr[s] = n < N ? W + (size_t)n * K : nullptr;
    }

    float4 a_reg[A_SLOTS], b_reg[B_SLOTS];
    auto load_tile = [&](int k0) {
        int g = 0;
        if (A.nseg > 1 && k0 >= A.seg[0].kend) g = 1;
        if (A.nseg > 2 && k0 >= A.seg[1].kend) g = 2;
#pragma unroll
        for (int s = 0; s < A_SLOTS; ++s) {
            const float* p = g == 0 ? a_ptr[s][0] : (g == 1 ? a_ptr[s][1] : a_ptr[s][2]);
            a_reg[s] = a_ok[s] ? *reinterpret_cast<const float4*>(p + k0 + 4 * a_kq[s])
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int s = 0; s < B_SLOTS; ++s) {
            b_reg[s] = b_ptr[s] ? *reinterpret_cast<const float4*>(b_ptr[s] + k0 + 4 * b_kq[s])
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int s = 0; s < A_SLOTS; ++s) {
            As[buf][4 * a_kq[s] + 0][a_row[s]] = a_reg[s].x;
            As[buf][4 * a_kq[s] + 1][a_row[s]] = a_reg[s].y;
            As[buf][4 * a_kq[s] + 2][a_row[s]] = a_reg[s].z;
            As[buf][4 * a_kq[s] + 3][a_row[s]] = a_reg[s].w;
        }
#pragma unroll
        for (int s = 0; s < B_SLOTS; ++s) {
            Bs[buf][4 * b_kq[s] + 0][b_row[s]] = b_reg[s].x;
            Bs[buf][4 * b_kq[s] + 1][b_row[s]] = b_reg[s].y;
            Bs[buf][4 * b_kq[s] + 2][b_row[s]] = b_reg[s].z;
            Bs[buf][4 * b_kq[s] + 3][b_row[s]] = b_reg[s].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int ntiles = K / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_tile((t + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (t + 1 < ntiles) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------
    const int col0 = n0 + tx * TN;
    if (epi.kind == Epi::kLstmCell) {
        // columns are gate-interleaved: n = 4*u + {i,f,g,o}
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int row = m0 + ty * TM + i;
            if (row >= M) continue;
            const int crow = epi.c_rowidx ? epi.c_rowidx[row] : row;
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const int n = col0 + j;
                if (n >= N) continue;
                const int u = n >> 2;
                const float gi = acc[i][j] + epi.bias[n];
                const float gf = acc[i][j + 1] + epi.bias[n + 1];
                const float gg = acc[i][j + 2] + epi.bias[n + 2];
                const float go = acc[i][j + 3] + epi.bias[n + 3];
                const float cp = epi.c_prev[(size_t)crow * epi.H + u];
                const float c = sigmoidf_acc(gf) * cp + sigmoidf_acc(gi) * tanhf(gg);
                const float hh = sigmoidf_acc(go) * tanhf(c);
                epi.c_out[(size_t)row * epi.H + u] = c;
                epi.h_out[(size_t)row * epi.H + u] = hh;
            }
        }
        return;
    }
    const bool scale = epi.kind == Epi::kBiasScale;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + ty * TM + i;
        if (row >= M) continue;
        float* crow = epi.C + (size_t)row * epi.ldc;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const int n = col0 + j;
            if (n + 3 < N) {
                float4 v;
                v.x = acc[i][j] + epi.bias[n];
                v.y = acc[i][j + 1] + epi.bias[n + 1];
                v.z = acc[i][j + 2] + epi.bias[n + 2];
                v.w = acc[i][j + 3] + epi.bias[n + 3];
                if (scale) { v.x /= epi.scale; v.y /= epi.scale; v.z /= epi.scale; v.w /= epi.scale; }
                *reinterpret_cast<float4*>(crow + n) = v;
            } else {
                for (int jj = 0; jj < 4; ++jj) {
                    if (n + jj < N) {
                        float v = acc[i][j + jj] + epi.bias[n + jj];
                        if (scale) v /= epi.scale;
                        crow[n + jj] = v;
                    }
                }
            }
        }
    }
}

int launch_gemm(const AOperand& A, const float* W, int M, int N, int K, const GemmEpilogue& epi,
                cudaStream_t st, int64_t* launches) {
    if (M <= 0) return ASR_OK;
    if (K % BK != 0) { set_error("gemm: K=%d not a multiple of %d", K, BK); return ASR_ERR_ARG; }
    for (int g = 0; g < A.nseg; ++g)
        if (A.seg[g].kend % BK != 0 || A.seg[g].ld % 4 != 0) {
            set_error("gemm: segment %d not aligned", g);
            return ASR_ERR_ARG;
        }
    if (epi.kind != Epi::kLstmCell && epi.ldc % 4 != 0) { set_error("gemm: ldc"); return ASR_ERR_ARG; }
    if (M >= 2048) {
        dim3 grid((N + 127) / 128, (M + 127) / 128);
        sgemm_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(A, W, M, N, K, epi);
    } else {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(A, W, M, N, K, epi);
    }
    ASR_CHECK_LAUNCH();
    if (launches) ++*launches;
    return ASR_OK;
}

}  // namespace asr
