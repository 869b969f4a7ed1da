// Decoder-side kernels of the greedy / beam loop (reference model.py:503-602, 604-987,
// attention.py:91-95, decoder.py:94-137).  The two GEMMs of a step (LSTM cell, vocabulary
// projection) are in gemm.cu; this file holds everything between and after them:
//
//   attention_kernel   : query projection h*W_hidden, additive scores v.tanh(keys + q), masked
//                        softmax over time and the context reduction, for all k beams of an
//                        utterance at once.  The encoder memory and keys of an utterance are read
//                        ONCE per step and shared by its k beams (the reference tiles them k times
//                        and re-gathers them every step, model.py:660-669, 913-916).  Long
//                        utterances are split over several CTAs (flash-decoding style partial
//                        softmax), the last CTA to finish combines the partials.
//   beam_merge_kernel  : per utterance: logsumexp of its k rows from the vocabulary GEMM's per-tile
//                        (max, sum) partials (model.py:835), log-prob + beam score (model.py:836) of the
//                        per-tile top candidates, exact top-2k over k*V (model.py:860-867), then
//                        EOS/finished bookkeeping (model.py:876-889), early stop flag (model.py:897-901),
//                        active-set selection, back-pointers and history (model.py:904-929) and the gather
//                        of the next step's cell-GEMM operand.  The logits themselves never reach HBM.  No
//                        host sync: the stop decision is a device flag every later kernel checks.
//   beam_finalise_kernel: parse_finished_tensors + second-pass LM rescoring + un-finished
//                        fallback (model.py:708-765, 945-987), n-gram LM lookups on device.
//   greedy_pick_kernel : argmax / score / length bookkeeping of the greedy loop (model.py:554-578).
#include <math_constants.h>

#include "asr_internal.cuh"

namespace asr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// tanh(x) = 1 - 2 / (1 + e^{2x}) with ex2.approx / rcp-based division: absolute error < 1e-6
// (libdevice tanhf costs ~3x the instructions; tanh.approx.f32 is only 2^-11 accurate)
__device__ __forceinline__ float tanh_acc(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}
// r = 1 / (1 + 2^x) on the MUFU pipe (ex2.approx + rcp.approx); 2^x = inf -> 0, 2^x = 0 -> 1
// (measured: replacing the rcp by a seed + 3 Newton steps on the FMA pipe is slower - the kernel then
// becomes issue-bound - so both transcendental steps stay on the MUFU pipe)
__device__ __forceinline__ float rcp1p_ex2(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// init
struct InitParams {
    const float* h0; const float* c0;
    float* dh; float* dc; float* dctx;
    int* src_row; float* beam_score; int* tok_hist; int* prev_hist; int* top_done;
    int* ctrl; int* att_ticket;
    float* g_accum; int* g_finished; int* g_len;
    int B, k, R;
};

__global__ void decode_init_kernel(InitParams p) {
    const int r = blockIdx.x;
    const int u = r / p.k;
    for (int c = threadIdx.x; c < kDecH; c += blockDim.x) {
        p.dh[(size_t)r * kDecH + c] = p.h0[(size_t)u * kEnc + c];
        p.dc[(size_t)r * kDecH + c] = p.c0[(size_t)u * kEnc + c];
        p.dctx[(size_t)r * kEnc + c] = 0.f;
    }
    if (threadIdx.x == 0) {
        p.src_row[r] = r;
        p.beam_score[r] = 0.f;
        p.tok_hist[r] = kSos;
        p.prev_hist[r] = r;
        if (r % p.k == 0) {
            p.top_done[u] = 0;
            p.att_ticket[u] = 0;
            p.g_accum[u] = 0.f;
            p.g_finished[u] = 0;
            p.g_len[u] = 0;
        }
        if (r == 0) {
            p.ctrl[0] = -1;   // stop step
            p.ctrl[1] = 0;
            p.ctrl[2] = 0;    // ticket
            p.ctrl[3] = 0;    // steps run
        }
    }
}

int decode_init(asr_handle* h, int k, int max_len, bool greedy, cudaStream_t st) {
    Workspace& w = h->ws;
    const int B = h->meta.B, R = B * k;
    InitParams p{w.h0, w.c0, w.dh[0], w.dc[0], w.dctx[0], w.src_row, w.beam_score, w.tok_hist,
                 w.prev_hist, w.top_done, w.ctrl, w.att_ticket, w.g_accum, w.g_finished, w.g_len,
                 B, k, R};
    ASR_CUDA(cudaMemsetAsync(w.tok_hist, 0, sizeof(int) * (size_t)(max_len + 1) * R, st));
    decode_init_kernel<<<R, 128, 0, st>>>(p);
    ASR_CHECK_LAUNCH();
    h->launches++;
    if (!greedy) {
        ASR_CUDA(cudaMemsetAsync(w.fin_row, 0xff, sizeof(int) * (size_t)max_len * B * k, st));
    }
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
// attention
struct AttnParams {
    const float* q;         // [R, 128] query projection h_new * W_hidden (GEMM engine)
    const float* keys;      // [rows, 128] utterance-major sorted
    const float* keys_exp;  // [rows, 128] 2^(key * 2 log2 e) = e^(2 key)      (keys_exp_kernel)
    const int* keys_big;    // [B] != 0: some |key| of the utterance is outside the product form's range
    const float* enc;       // [rows, 512]
    const float* v;         // [128]
    const int* uoff;        // [B + 1]
    float* ctx_out;         // [R, 512]
    hi_t* split_hi;         // optional: fp16 hi / bf16 cross split of ctx written at column 512 of the
    float* split_lo;        //   [R, split_ld] A operand of the vocabulary GEMM
    int split_ld;
    float* part;            // [B, S, k, 514]
    int* ticket;            // [B]
    float* raw_score;       // [R, score_ld] or nullptr (alignment export)
    float* align_out;       // [Lmax, B] original order for this step, or nullptr
    const int* order;       // [B]
    const int* ctrl;
    int k, S, B, Lmax, sc_ld;
    long long score_ld;
};

__device__ __forceinline__ void write_ctx_split(const AttnParams& p, int row, int tid, float2 o) {
    if (!p.split_hi) return;
    // thread `tid` owns ctx[2 tid], ctx[2 tid + 1].  hi = fp16(x); the cross operand holds, per
    // 8-float block, 8 x bf16(x - hi) then 8 x bf16(x) (gemm_tc.cu: kSplitAct)
    float2 hi;
    uint32_t xl, xx;
    hi.x = hi_part(o.x);
    hi.y = hi_part(o.y);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xl) : "f"(o.y - hi.y), "f"(o.x - hi.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xx) : "f"(o.y), "f"(o.x));
    reinterpret_cast<uint32_t*>(p.split_hi + (size_t)row * p.split_ld + kDecH)[tid] = pack_hi2(hi.x, hi.y);
    uint32_t* cross = reinterpret_cast<uint32_t*>(p.split_lo + (size_t)row * p.split_ld + kDecH) + (tid >> 2) * 8 + (tid & 3);
    cross[0] = xl;
    cross[4] = xx;
}

// e^(2 (key + q)) = e^(2 key) * e^(2 q): with the per-frame factor precomputed once per batch (here)
// and the per-beam factor once per step (phase A of the attention kernel), one score element costs
// FFMA, MUFU.RCP, FFMA instead of FADD, MUFU.EX2, FADD, MUFU.RCP, FFMA - half the MUFU work that bounds
// the kernel.  Both factors are kept inside 2^+-60 so the product cannot overflow or hit inf * 0; an
// utterance (or a step's queries) with a pre-activation outside that range (|x| > 20.8, where tanh has
// long been +-1 to fp32) takes the exact sum-then-exp path instead.
constexpr float kAttScale = 2.885390081777927f;       // 2 * log2(e)
constexpr float kAttRange = 60.f;

__global__ void keys_exp_kernel(const float* __restrict__ keys, const int* __restrict__ uoff,
                                float* __restrict__ keys_exp, int* __restrict__ keys_big) {
    const int u = blockIdx.x;
    const long long beg = (long long)uoff[u] * (kAtt / 4), end = (long long)uoff[u + 1] * (kAtt / 4);
    bool big = false;
    for (long long i = beg + blockIdx.y * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.y * blockDim.x) {
        float4 x = reinterpret_cast<const float4*>(keys)[i];
        x.x *= kAttScale; x.y *= kAttScale; x.z *= kAttScale; x.w *= kAttScale;
        big |= !(fabsf(x.x) <= kAttRange && fabsf(x.y) <= kAttRange && fabsf(x.z) <= kAttRange && fabsf(x.w) <= kAttRange);
        reinterpret_cast<float4*>(keys_exp)[i] = make_float4(exp2f(x.x), exp2f(x.y), exp2f(x.z), exp2f(x.w));
    }
    if (__syncthreads_or(big) && threadIdx.x == 0) atomicOr(keys_big + u, 1);
}

int launch_keys_exp(asr_handle* h, cudaStream_t st) {
    Workspace& w = h->ws;
    const BatchMeta& m = h->meta;
    ASR_CUDA(cudaMemsetAsync(w.keys_big, 0, sizeof(int) * (size_t)m.B, st));
    const int slices = std::max(1, std::min(8, (4 * kNumSMs + m.B - 1) / m.B));
    keys_exp_kernel<<<dim3(m.B, slices), 256, 0, st>>>(w.keys, m.d_uoff_sorted, w.keys_exp, w.keys_big);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

template <int K>
__global__ void __launch_bounds__(256, K <= 8 ? 4 : 2)
attention_kernel(AttnParams p) {
    if (p.ctrl[0] >= 0) return;
    extern __shared__ __align__(16) float sm[];
    float* s_q = sm;                        // [K][128]
    float* s_sc = sm + K * kAtt;            // [K][sc_ld]
    __shared__ float s_m[K], s_sum[K];
    __shared__ int s_last;

    const int u = blockIdx.x, sp = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.k;
    const int row0 = p.uoff[u];
    const int L = p.uoff[u + 1] - row0;
    const int Lc = (L + p.S - 1) / p.S;
    const int lbeg = min(L, sp * Lc), lend = min(L, lbeg + Lc);
    const int nl = lend - lbeg;

    // ---- phase A: the query projection q = h * W_hidden was computed by the GEMM engine --------
    bool q_big = false;
    for (int i = tid; i < k * kAtt; i += 256) {
        const float qv = p.q[(size_t)u * k * kAtt + i];
        s_q[i] = qv;
        q_big |= !(fabsf(qv * kAttScale) <= kAttRange);
    }
    const bool product_form = !__syncthreads_or(q_big) && p.keys_big[u] == 0;

    // ---- phase B: scores e[l][kb] = sum_d v_d tanh(key[l][d] + q[kb][d]) ----------------------
    // one warp per frame, lane owns 4 of the 128 attention dims for all K beams; the K partial
    // sums are reduced with a transposed butterfly (K/2 + K/4 + .. shuffles instead of 5 K)
    // tanh(x) = 1 - 2 r with r = 1 / (1 + 2^(x * 2 log2 e)):  e = V - 2 * sum_d v_d r_d, V = sum_d v_d.
    // Keys and queries are pre-scaled by 2 log2(e) (once per frame / per beam), so one element costs
    // FADD, MUFU.EX2, FADD, MUFU.RCP, FFMA - the kernel is bound by the 16 lanes/clk MUFU pipe.
    {
        constexpr float kScale = kAttScale;
        const float4 v4 = *reinterpret_cast<const float4*>(p.v + 4 * lane);
        const float vsum = (v4.x + v4.y) + (v4.z + v4.w);
        float4 q4[K];
#pragma unroll
        for (int kb = 0; kb < K; ++kb) {
            q4[kb] = kb < k ? *reinterpret_cast<const float4*>(s_q + kb * kAtt + 4 * lane)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            q4[kb].x *= kScale; q4[kb].y *= kScale; q4[kb].z *= kScale; q4[kb].w *= kScale;
            if (product_form) {      // per-beam factor e^(2 q)
                q4[kb].x = exp2f(q4[kb].x); q4[kb].y = exp2f(q4[kb].y);
                q4[kb].z = exp2f(q4[kb].z); q4[kb].w = exp2f(q4[kb].w);
            }
        }
        constexpr int kGroup = 32 / K;                 // lanes that end up holding the same beam
        const float* kbase = product_form ? p.keys_exp : p.keys;
        for (int l = lbeg + warp; l < lend; l += 8) {
            float4 key = __ldg(reinterpret_cast<const float4*>(kbase + (size_t)(row0 + l) * kAtt) + lane);
            float e[K];
            if (product_form) {
#pragma unroll
                for (int kb = 0; kb < K; ++kb) {
                    float t = v4.x * rcp_approx(fmaf(key.x, q4[kb].x, 1.f));
                    t = fmaf(v4.y, rcp_approx(fmaf(key.y, q4[kb].y, 1.f)), t);
                    t = fmaf(v4.z, rcp_approx(fmaf(key.z, q4[kb].z, 1.f)), t);
                    t = fmaf(v4.w, rcp_approx(fmaf(key.w, q4[kb].w, 1.f)), t);
                    e[kb] = fmaf(-2.f, t, vsum);
                }
            } else {
                key.x *= kScale; key.y *= kScale; key.z *= kScale; key.w *= kScale;
#pragma unroll
                for (int kb = 0; kb < K; ++kb) {
                    float t = v4.x * rcp1p_ex2(key.x + q4[kb].x);
                    t = fmaf(v4.y, rcp1p_ex2(key.y + q4[kb].y), t);
                    t = fmaf(v4.z, rcp1p_ex2(key.z + q4[kb].z), t);
                    t = fmaf(v4.w, rcp1p_ex2(key.w + q4[kb].w), t);
                    e[kb] = fmaf(-2.f, t, vsum);
                }
            }
            int off = 16;
#pragma unroll
            for (int n = K; n > 1; n >>= 1) {
                const int half = n >> 1;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    const float send = up ? e[j] : e[j + half];
                    const float keep = up ? e[j + half] : e[j];
                    e[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
                off >>= 1;
            }
            for (; off > 0; off >>= 1) e[0] += __shfl_xor_sync(0xffffffffu, e[0], off);
            const int kb = lane / kGroup;
            if ((lane % kGroup) == 0 && kb < k) s_sc[kb * p.sc_ld + (l - lbeg)] = e[0];
        }
    }
    __syncthreads();

    // ---- phase C: local softmax statistics -----------------------------------------------------
    for (int kb = warp; kb < k; kb += 8) {
        float* sc = s_sc + kb * p.sc_ld;
        float m = -CUDART_INF_F;
        for (int l = lane; l < nl; l += 32) m = fmaxf(m, sc[l]);
        m = warp_max(m);
        float sum = 0.f;
        for (int l = lane; l < nl; l += 32) {
            const float e = sc[l];
            if (p.raw_score) p.raw_score[(size_t)(u * k + kb) * p.score_ld + lbeg + l] = e;
            const float pe = expf(e - m);
            sc[l] = pe;
            sum += pe;
        }
        sum = warp_sum(sum);
        if (lane == 0) { s_m[kb] = m; s_sum[kb] = sum; }
    }
    __syncthreads();

    // ---- phase D: context partial: ctx[kb][c] = sum_l p[l][kb] * enc[l][c] ---------------------
    // HBM-bound (the encoder memory of the batch does not fit in L2): each half-CTA streams half of
    // the frames with 8 independent 16-byte loads in flight per thread (thread = 4 of the 512
    // columns, all K beams), packed fp32x2 FMAs; the two halves are summed through shared memory.
    float2 acc[K];
    {
        float* s_ctx = s_sc + K * p.sc_ld;                 // [K][512]
        const int cg4 = tid & 127, hf = tid >> 7;
        const int nh = (((nl + 1) >> 1) + 3) & ~3;         // frames per half, multiple of 4
        const int f0 = min(nl, hf * nh), f1 = min(nl, f0 + nh);
        float2 a01[K], a23[K];
#pragma unroll
        for (int kb = 0; kb < K; ++kb) { a01[kb] = make_float2(0.f, 0.f); a23[kb] = a01[kb]; }
        const float4* encp = reinterpret_cast<const float4*>(p.enc + (size_t)(row0 + lbeg) * kEnc) + cg4;
        int l = f0;
        for (; l + 8 <= f1; l += 8) {
            float4 e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = __ldg(encp + (size_t)(l + i) * (kEnc / 4));
#pragma unroll
            for (int kb = 0; kb < K; ++kb) {
                if (kb < k) {
                    const float4 w0 = *reinterpret_cast<const float4*>(s_sc + kb * p.sc_ld + l);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_sc + kb * p.sc_ld + l + 4);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        a01[kb] = __ffma2_rn(make_float2(w[i], w[i]), make_float2(e[i].x, e[i].y), a01[kb]);
                        a23[kb] = __ffma2_rn(make_float2(w[i], w[i]), make_float2(e[i].z, e[i].w), a23[kb]);
                    }
                }
            }
        }
        for (; l < f1; ++l) {
            const float4 e0 = __ldg(encp + (size_t)l * (kEnc / 4));
#pragma unroll
            for (int kb = 0; kb < K; ++kb) {
                if (kb < k) {
                    const float pe = s_sc[kb * p.sc_ld + l];
                    a01[kb] = __ffma2_rn(make_float2(pe, pe), make_float2(e0.x, e0.y), a01[kb]);
                    a23[kb] = __ffma2_rn(make_float2(pe, pe), make_float2(e0.z, e0.w), a23[kb]);
                }
            }
        }
        if (hf == 1) {
#pragma unroll
            for (int kb = 0; kb < K; ++kb)
                if (kb < k) *reinterpret_cast<float4*>(s_ctx + kb * kEnc + 4 * cg4) = make_float4(a01[kb].x, a01[kb].y, a23[kb].x, a23[kb].y);
        }
        __syncthreads();
        if (hf == 0) {
#pragma unroll
            for (int kb = 0; kb < K; ++kb)
                if (kb < k) {
                    float4 o = *reinterpret_cast<float4*>(s_ctx + kb * kEnc + 4 * cg4);
                    o.x += a01[kb].x; o.y += a01[kb].y; o.z += a23[kb].x; o.w += a23[kb].y;
                    *reinterpret_cast<float4*>(s_ctx + kb * kEnc + 4 * cg4) = o;
                }
        }
        __syncthreads();
#pragma unroll
        for (int kb = 0; kb < K; ++kb)
            acc[kb] = kb < k ? *reinterpret_cast<float2*>(s_ctx + kb * kEnc + 2 * tid) : make_float2(0.f, 0.f);
    }

    if (p.S == 1) {
#pragma unroll
        for (int kb = 0; kb < K; ++kb) {
            if (kb < k) {
                const float inv = s_sum[kb];
                float2 o = make_float2(acc[kb].x / inv, acc[kb].y / inv);
                reinterpret_cast<float2*>(p.ctx_out + (size_t)(u * k + kb) * kEnc)[tid] = o;
                write_ctx_split(p, u * k + kb, tid, o);
            }
        }
        if (p.align_out) {       // greedy only (k == 1)
            const float inv = s_sum[0];
            for (int l = tid; l < p.Lmax; l += 256)
                p.align_out[(size_t)l * p.B + p.order[u]] = l < L ? s_sc[l] / inv : 0.f;
        }
        return;
    }

    // ---- phase E: write partials; the last CTA of the utterance combines -----------------------
    float* part = p.part + ((size_t)(u * p.S + sp) * k) * 514;
#pragma unroll
    for (int kb = 0; kb < K; ++kb) {
        if (kb < k) {
            reinterpret_cast<float2*>(part + (size_t)kb * 514 + 2)[tid] = acc[kb];
            if (tid == 0) { part[(size_t)kb * 514] = s_m[kb]; part[(size_t)kb * 514 + 1] = s_sum[kb]; }
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int t = atomicAdd(p.ticket + u, 1);
        s_last = (t == p.S - 1);
        if (s_last) p.ticket[u] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* pu = p.part + (size_t)u * p.S * k * 514;
    for (int kb = 0; kb < k; ++kb) {
        float M = -CUDART_INF_F;
        for (int s = 0; s < p.S; ++s) M = fmaxf(M, __ldcg(pu + ((size_t)s * k + kb) * 514));
        float denom = 0.f;
        float2 o = make_float2(0.f, 0.f);
        for (int s = 0; s < p.S; ++s) {
            const float* ps = pu + ((size_t)s * k + kb) * 514;
            const float ms = __ldcg(ps);
            const float wgt = ms == -CUDART_INF_F ? 0.f : expf(ms - M);
            denom = fmaf(wgt, __ldcg(ps + 1), denom);
            const float2 c = __ldcg(reinterpret_cast<const float2*>(ps + 2) + tid);
            o.x = fmaf(wgt, c.x, o.x);
            o.y = fmaf(wgt, c.y, o.y);
        }
        o.x /= denom; o.y /= denom;
        reinterpret_cast<float2*>(p.ctx_out + (size_t)(u * k + kb) * kEnc)[tid] = o;
        write_ctx_split(p, u * k + kb, tid, o);
        if (p.align_out && kb == 0) {
            const float* rs = p.raw_score + (size_t)(u * k) * p.score_ld;
            for (int l = tid; l < p.Lmax; l += 256)
                p.align_out[(size_t)l * p.B + p.order[u]] = l < L ? expf(__ldcg(rs + l) - M) / denom : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming (single pass) attention: the two-phase kernel above runs its MUFU-bound score phase and its
// HBM-bound context phase one after the other, and since every CTA of the grid starts at the same time
// the whole GPU alternates between the two limits.  Here warps 0-3 of a CTA produce softmax numerators
// chunk by chunk (64 frames) while warps 4-7 stream the encoder rows of the previous chunk, so both
// limits are worked on at once.  Softmax is "online": numerators are taken against the running
// maximum M and the accumulators are rescaled by exp(M_old - M_new) when a chunk raises it
// (same value as softmax-then-sum up to fp32 rounding of the rescale factors).
// Hand-off through named barriers: full[b] (producers arrive, consumers sync), empty[b] (the reverse).
constexpr int kAttChunk = 32;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void att_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void att_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void att_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void att_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 28)) __trap();
    }
}
// contiguous global -> shared bulk copy (TMA engine, no registers), completion on an mbarrier
__device__ __forceinline__ void att_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

constexpr int kAttRows = 8;        // encoder rows per ring stage (16 KB)
constexpr int kAttStages = 2;      // the consumers drain a stage much faster than HBM delivers one: two keep a copy in flight
constexpr int kAttBufs = 4;        // numerator chunks the producers may run ahead of the consumers

// per-lane asynchronous copy of 16 bytes global -> shared (LDGSTS): the producers' key prefetch, no registers held
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared memory of the streaming kernel (floats): encoder ring | per-warp key ring (aliased by the queries during
// the prologue) | numerator buffers | per-buffer rescale factors | chunk maxima | sums | mbarriers
template <int K>
struct AttSmem {
    static constexpr int kRing = kAttStages * kAttRows * kEnc;
    static constexpr int kKeys = 4 * (kAttChunk / 4) * kAtt;            // 4 producer warps x 8 frames x 128
    static_assert(K * kAtt <= kKeys, "queries alias the key ring");
    static constexpr int kPStride = kAttChunk + 4;      // beam rows 36 floats apart: the K leader lanes of a warp hit K different banks
    static constexpr int kP = kAttBufs * K * kPStride;
    static constexpr int kSmall = kAttBufs * K + 2 * 4 * K + 4 * K;
    static constexpr size_t kBytes = sizeof(float) * (kRing + kKeys + kP + kSmall) + sizeof(uint64_t) * 2 * kAttStages + 16;
};

// FULL: k == K (every beam slot of the instantiation is live: no per-beam branches in the consumers' inner loop)
template <int K, bool FULL>
__global__ void __launch_bounds__(256, K <= 8 ? 4 : 2)
attention_stream_kernel(AttnParams p) {
    griddep_launch_dependents();
    constexpr int C = kAttChunk;
    constexpr int kGroup = 32 / K;                 // lanes that end up holding the same beam
    constexpr int FPW = C / 4;                     // frames per producer warp per chunk
    using S = AttSmem<K>;
    extern __shared__ __align__(128) uint8_t att_smem[];
    float* s_ring = reinterpret_cast<float*>(att_smem);                          // [stages][8][512]
    float* s_keys = s_ring + S::kRing;                                            // [4 warps][FPW][128]
    float* s_q = s_keys;                                                          // [K][128], prologue only
    float* s_p = s_keys + S::kKeys;                                               // [bufs][K][C + 4]
    constexpr int CP = S::kPStride;
    float* s_scale = s_p + S::kP;                                                 // [bufs][K]
    float* s_wmax = s_scale + kAttBufs * K;                                       // [2][4][K]
    float* s_wsum = s_wmax + 2 * 4 * K;                                           // [4][K]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wsum + 4 * K);                 // full[stages], empty[stages]

    const int u = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.k;
    const int row0 = p.uoff[u];
    const int nl = p.uoff[u + 1] - row0;
    const int nchunk = (nl + C - 1) / C;
    const int nstage = (nl + kAttRows - 1) / kAttRows;
    uint64_t* full_e = bars;
    uint64_t* empty_e = bars + kAttStages;
    // named barriers: 1 + b: numerators of buffer b published (producers arrive, consumers sync);
    // 1 + kAttBufs + b: buffer b drained (consumers arrive, producers sync); 1 + 2 kAttBufs: producers only

    if (tid == 0) {
        for (int i = 0; i < kAttStages; ++i) { att_mbar_init(&full_e[i], 1); att_mbar_init(&empty_e[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    griddep_wait();                                        // the query GEMM has finished: global memory may be touched
    if (p.ctrl[0] >= 0) return;
    bool q_big = false;
    for (int i = tid; i < k * kAtt; i += 256) {
        const float qv = p.q[(size_t)u * k * kAtt + i];
        s_q[i] = qv;
        q_big |= !(fabsf(qv * kAttScale) <= kAttRange);
    }
    const bool product_form = !__syncthreads_or(q_big) && p.keys_big[u] == 0;

    if (warp < 4) {
        // ---------------- producers: scores -> numerators ------------------------------------------
        const float4 v4 = *reinterpret_cast<const float4*>(p.v + 4 * lane);
        const float vsum = (v4.x + v4.y) + (v4.z + v4.w);
        float4 q4[K];
#pragma unroll
        for (int kb = 0; kb < K; ++kb) {
            q4[kb] = kb < k ? *reinterpret_cast<const float4*>(s_q + kb * kAtt + 4 * lane)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            q4[kb].x *= kAttScale; q4[kb].y *= kAttScale; q4[kb].z *= kAttScale; q4[kb].w *= kAttScale;
            if (product_form) {
                q4[kb].x = exp2f(q4[kb].x); q4[kb].y = exp2f(q4[kb].y);
                q4[kb].z = exp2f(q4[kb].z); q4[kb].w = exp2f(q4[kb].w);
            }
        }
        named_bar_sync(1 + 2 * kAttBufs, 128);             // every producer has its queries: the key ring may be filled
        // Key rows stream through a per-warp ring of FPW slots: lane copies (and later reads) its own 16 bytes of a
        // frame's 128-float row.  The slot of frame f is refilled with frame f of the NEXT chunk right after it was
        // consumed, i.e. one chunk (FPW frames of compute) ahead - the __ldg it replaces was one frame ahead and
        // left the producers waiting on HBM latency (ncu r02a: 18 % of all stall samples on that load).
        const float* kbase = product_form ? p.keys_exp : p.keys;
        float* my_keys = s_keys + (warp * FPW) * kAtt + 4 * lane;
        auto prefetch_key = [&](int c, int f) {
            if (c < nchunk) {
                const int l = min(c * C + warp + 4 * f, nl - 1);
                cp_async16(my_keys + f * kAtt, kbase + (size_t)(row0 + l) * kAtt + 4 * lane);
            }
            cp_async_commit();                             // one group per frame, empty past the last chunk
        };
#pragma unroll
        for (int f = 0; f < FPW; ++f) prefetch_key(0, f);
        const int my_kb = lane / kGroup;
        const bool leader = (lane % kGroup) == 0 && my_kb < k;
        float run_max = -CUDART_INF_F, run_sum = 0.f;      // leader lanes: state of beam my_kb
        for (int c = 0; c < nchunk; ++c) {
            const int buf = c % kAttBufs;
            const int c0 = c * C;
            float* pb = s_p + (buf * K + my_kb) * CP;      // this lane's beam row of the chunk
            if (c >= kAttBufs) named_bar_sync(1 + kAttBufs + buf, 256);      // buffer drained by the consumers
            float cmax = -CUDART_INF_F;
#pragma unroll 1
            for (int f = 0; f < FPW; ++f) {
                const int l = c0 + warp + 4 * f;
                float ev = -CUDART_INF_F;
                cp_async_wait<FPW - 1>();                  // the oldest group = this frame's key row has landed
                float4 key = *reinterpret_cast<const float4*>(my_keys + f * kAtt);
                if (l < nl) {                              // warp-uniform
                    float e[K];
                    if (product_form) {
                        if (K >= 2) {
                            // two beams per packed FFMA2: d = key * q + 1, then t += v * rcp(d)
#pragma unroll
                            for (int kb = 0; kb < K; kb += 2) {
                                const int k1 = kb + 1 < K ? kb + 1 : kb;
                                const float2 one = make_float2(1.f, 1.f);
                                const float2 dx = __ffma2_rn(make_float2(key.x, key.x), make_float2(q4[kb].x, q4[k1].x), one);
                                const float2 dy = __ffma2_rn(make_float2(key.y, key.y), make_float2(q4[kb].y, q4[k1].y), one);
                                const float2 dz = __ffma2_rn(make_float2(key.z, key.z), make_float2(q4[kb].z, q4[k1].z), one);
                                const float2 dw = __ffma2_rn(make_float2(key.w, key.w), make_float2(q4[kb].w, q4[k1].w), one);
                                float2 t = __fmul2_rn(make_float2(v4.x, v4.x), make_float2(rcp_approx(dx.x), rcp_approx(dx.y)));
                                t = __ffma2_rn(make_float2(v4.y, v4.y), make_float2(rcp_approx(dy.x), rcp_approx(dy.y)), t);
                                t = __ffma2_rn(make_float2(v4.z, v4.z), make_float2(rcp_approx(dz.x), rcp_approx(dz.y)), t);
                                t = __ffma2_rn(make_float2(v4.w, v4.w), make_float2(rcp_approx(dw.x), rcp_approx(dw.y)), t);
                                const float2 ee = __ffma2_rn(make_float2(-2.f, -2.f), t, make_float2(vsum, vsum));
                                e[kb] = ee.x;
                                e[k1] = ee.y;
                            }
                        } else {
                            float t = v4.x * rcp_approx(fmaf(key.x, q4[0].x, 1.f));
                            t = fmaf(v4.y, rcp_approx(fmaf(key.y, q4[0].y, 1.f)), t);
                            t = fmaf(v4.z, rcp_approx(fmaf(key.z, q4[0].z, 1.f)), t);
                            t = fmaf(v4.w, rcp_approx(fmaf(key.w, q4[0].w, 1.f)), t);
                            e[0] = fmaf(-2.f, t, vsum);
                        }
                    } else {
                        key.x *= kAttScale; key.y *= kAttScale; key.z *= kAttScale; key.w *= kAttScale;
#pragma unroll
                        for (int kb = 0; kb < K; ++kb) {
                            float t = v4.x * rcp1p_ex2(key.x + q4[kb].x);
                            t = fmaf(v4.y, rcp1p_ex2(key.y + q4[kb].y), t);
                            t = fmaf(v4.z, rcp1p_ex2(key.z + q4[kb].z), t);
                            t = fmaf(v4.w, rcp1p_ex2(key.w + q4[kb].w), t);
                            e[kb] = fmaf(-2.f, t, vsum);
                        }
                    }
                    int off = 16;
#pragma unroll
                    for (int n = K; n > 1; n >>= 1) {
                        const int half = n >> 1;
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int j = 0; j < half; ++j) {
                            const float send = up ? e[j] : e[j + half];
                            const float keep = up ? e[j + half] : e[j];
                            e[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                        off >>= 1;
                    }
                    for (; off > 0; off >>= 1) e[0] += __shfl_xor_sync(0xffffffffu, e[0], off);
                    ev = e[0];
                }
                // the slot has been read (its value feeds the arithmetic above): refill it one chunk ahead
                prefetch_key(c + 1, f);
                if (leader) pb[warp + 4 * f] = ev;         // raw score, turned into a numerator below
                cmax = fmaxf(cmax, ev);
            }
            if (leader) s_wmax[((c & 1) * 4 + warp) * K + my_kb] = cmax;
            named_bar_sync(1 + 2 * kAttBufs, 128);         // chunk maxima of the 4 producer warps
            if (leader) {
                const float* wm = s_wmax + (c & 1) * 4 * K + my_kb;
                const float new_max = fmaxf(run_max, fmaxf(fmaxf(wm[0], wm[K]), fmaxf(wm[2 * K], wm[3 * K])));
                const float scale = __expf(run_max - new_max);     // first chunk: exp(-inf) = 0 (accumulators are 0)
                run_max = new_max;
                float csum = 0.f;
#pragma unroll
                for (int f = 0; f < FPW; ++f) {
                    const float pe = __expf(pb[warp + 4 * f] - run_max);      // frames past the end: exp(-inf) = 0
                    pb[warp + 4 * f] = pe;
                    csum += pe;
                }
                run_sum = fmaf(run_sum, scale, csum);
                if (warp == 0) s_scale[buf * K + my_kb] = scale;
            }
            __threadfence_block();
            named_bar_arrive(1 + buf, 256);
        }
        cp_async_wait<0>();
        if (leader) s_wsum[warp * K + my_kb] = run_sum;
    } else {
        // ---------------- consumers: context accumulation -------------------------------------------
        // All shared-memory traffic of the inner loop goes through 32-bit shared-window addresses computed ONCE and
        // `ld.shared` with immediate offsets: with generic pointers the compiler rebuilt the window base (S2UR
        // SR_CgaCtaId + ULEA) and the row offset for every beam - 9 overhead instructions per 8 FFMA2 (ncu r02b:
        // FFMA2 were 34 % of the consumers' instructions, and the producers spent 43 % of their time waiting for them).
        const int cg4 = tid - 128;                         // 4 of the 512 encoder columns
        float2 a01[K], a23[K];
#pragma unroll
        for (int kb = 0; kb < K; ++kb) { a01[kb] = make_float2(0.f, 0.f); a23[kb] = a01[kb]; }
        const float* enc_u = p.enc + (size_t)row0 * kEnc;
        const uint32_t ring_u32 = smem_addr(s_ring) + (uint32_t)cg4 * 16u;
        const uint32_t p_u32 = smem_addr(s_p);
        const uint32_t scale_u32 = smem_addr(s_scale);
        const uint32_t full_u32 = smem_addr(full_e), empty_u32 = smem_addr(empty_e);
        auto issue = [&](int st) {                         // rows [8 st, 8 st + 8) of the utterance -> ring slot
            const int slot = st % kAttStages;
            const int rows = min(kAttRows, nl - st * kAttRows);
            const uint32_t bytes = (uint32_t)rows * kEnc * 4u;
            att_mbar_expect_tx(&full_e[slot], bytes);
            att_bulk_g2s(s_ring + slot * kAttRows * kEnc, enc_u + (size_t)st * kAttRows * kEnc, bytes, &full_e[slot]);
        };
        if (tid == 128) for (int st = 0; st < min(kAttStages, nstage); ++st) issue(st);
        for (int st = 0; st < nstage; ++st) {
            const int slot = st % kAttStages;
            const int l0 = st * kAttRows;
            const int buf = (l0 / C) % kAttBufs;
            const int lc = l0 % C;
            if (lc == 0) {
                named_bar_sync(1 + buf, 256);              // this chunk's numerators are ready
#pragma unroll
                for (int kb = 0; kb < K; ++kb) {
                    if (FULL || kb < k) {
                        float sc;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc) : "r"(scale_u32 + (uint32_t)((buf * K + kb) * 4)) : "memory");
                        a01[kb].x *= sc; a01[kb].y *= sc; a23[kb].x *= sc; a23[kb].y *= sc;
                    }
                }
            }
            {   // full[slot]: the stage's rows have landed
                const uint32_t bar = full_u32 + (uint32_t)slot * 8u, par = (uint32_t)((st / kAttStages) & 1);
                uint32_t done = 0;
                unsigned spins = 0;
                while (!done) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(done) : "r"(bar), "r"(par) : "memory");
                    if (!done && ++spins > (1u << 28)) __trap();
                }
            }
            const int rows = min(kAttRows, nl - l0);
            const uint32_t erow = ring_u32 + (uint32_t)(slot * kAttRows * kEnc * 4);
            const uint32_t wrow = p_u32 + (uint32_t)(((buf * K) * CP + lc) * 4);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {               // 4 rows at a time: 16 registers of encoder data
                float4 e[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (4 * hf + i < rows)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(e[i].x), "=f"(e[i].y), "=f"(e[i].z), "=f"(e[i].w)
                                     : "r"(erow + (uint32_t)((4 * hf + i) * kEnc * 4)) : "memory");
                }
                if (hf == 1) {
                    __syncwarp();
                    if (lane == 0)                          // this warp has read the slot
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_u32 + (uint32_t)slot * 8u) : "memory");
                }
#pragma unroll
                for (int kb = 0; kb < K; ++kb) {
                    if (FULL || kb < k) {
                        float4 w4;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w)
                                     : "r"(wrow + (uint32_t)((kb * CP + 4 * hf) * 4)) : "memory");
                        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            a01[kb] = __ffma2_rn(make_float2(w[i], w[i]), make_float2(e[i].x, e[i].y), a01[kb]);
                            a23[kb] = __ffma2_rn(make_float2(w[i], w[i]), make_float2(e[i].z, e[i].w), a23[kb]);
                        }
                    }
                }
            }
            if (lc + kAttRows == C || st + 1 == nstage) {              // last stage of the chunk
                const int c = l0 / C;
                if (c + kAttBufs < nchunk) named_bar_arrive(1 + kAttBufs + buf, 256);
            }
            if (tid == 128 && st + kAttStages < nstage) {
                att_mbar_wait(&empty_e[slot], (uint32_t)((st / kAttStages) & 1));   // all 4 warps have read it
                issue(st + kAttStages);
            }
        }
        __syncthreads();                                   // producers' sums
#pragma unroll
        for (int kb = 0; kb < K; ++kb) {
            if (kb < k) {
                const float inv = (s_wsum[kb] + s_wsum[K + kb]) + (s_wsum[2 * K + kb] + s_wsum[3 * K + kb]);
                const float4 o = make_float4(a01[kb].x / inv, a01[kb].y / inv, a23[kb].x / inv, a23[kb].y / inv);
                const int row = u * k + kb;
                *reinterpret_cast<float4*>(p.ctx_out + (size_t)row * kEnc + 4 * cg4) = o;
                if (p.split_hi) {
                    // operand split of ctx for the vocabulary GEMM (gemm_tc.cu kSplitAct): this thread
                    // owns half of an 8-float block
                    const float4 hi = make_float4(hi_part(o.x), hi_part(o.y), hi_part(o.z), hi_part(o.w));
                    *reinterpret_cast<uint2*>(p.split_hi + (size_t)row * p.split_ld + kDecH + 4 * cg4) =
                        make_uint2(pack_hi2(hi.x, hi.y), pack_hi2(hi.z, hi.w));
                    uint2 lo2, xx2;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2.x) : "f"(o.y - hi.y), "f"(o.x - hi.x));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2.y) : "f"(o.w - hi.w), "f"(o.z - hi.z));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xx2.x) : "f"(o.y), "f"(o.x));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xx2.y) : "f"(o.w), "f"(o.z));
                    uint32_t* blk = reinterpret_cast<uint32_t*>(p.split_lo + (size_t)row * p.split_ld + kDecH) + (cg4 >> 1) * 8;
                    *reinterpret_cast<uint2*>(blk + 2 * (cg4 & 1)) = lo2;
                    *reinterpret_cast<uint2*>(blk + 4 + 2 * (cg4 & 1)) = xx2;
                }
            }
        }
        return;
    }
    __syncthreads();                                       // matches the consumers' barrier above
}

template <int K>
static int launch_attention_stream(const AttnParams& p, int B, cudaStream_t st) {
    constexpr size_t smem = AttSmem<K>::kBytes;
    if (p.k == K) {
        ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attention_stream_kernel<K, true>), smem));
        ASR_CUDA(launch_kernel(attention_stream_kernel<K, true>, dim3(B), dim3(256), smem, st, true, p));
    } else {
        ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attention_stream_kernel<K, false>), smem));
        ASR_CUDA(launch_kernel(attention_stream_kernel<K, false>, dim3(B), dim3(256), smem, st, true, p));
    }
    return ASR_OK;
}

int launch_attention(asr_handle* h, int k, int step, int nxt, float* d_align_step, cudaStream_t st) {
    (void)step;
    Workspace& w = h->ws;
    const BatchMeta& m = h->meta;
    AttnParams p{};
    p.q = w.att_q;
    p.keys = w.keys;
    p.keys_exp = w.keys_exp;
    p.keys_big = w.keys_big;
    p.enc = w.enc;
    p.v = h->w.att_v;
    p.uoff = m.d_uoff_sorted;
    p.ctx_out = w.dctx[nxt];
    p.split_hi = w.dec_split_hi;
    p.split_lo = w.dec_split_lo;
    p.split_ld = kProjK;
    p.part = w.att_part;
    p.ticket = w.att_ticket;
    p.order = m.d_order;
    p.ctrl = w.ctrl;
    p.k = k;
    p.B = m.B;
    p.Lmax = m.Lmax;
    int S = (2 * kNumSMs + m.B - 1) / m.B;
    if (S > 8) S = 8;
    if (S < 1) S = 1;
    while (S > 1 && (m.Lmax + S - 1) / S < 16) --S;
    p.S = S;
    p.sc_ld = ((m.Lmax + S - 1) / S + 3) & ~3;
    p.score_ld = w.att_score_ld;
    p.align_out = d_align_step;
    p.raw_score = (d_align_step && S > 1) ? w.att_score : nullptr;
    const int K = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : 16));
    if (S == 1 && !d_align_step) {
        switch (K) {
            case 1: ASR_TRY(launch_attention_stream<1>(p, m.B, st)); break;
            case 4: ASR_TRY(launch_attention_stream<4>(p, m.B, st)); break;
            case 8: ASR_TRY(launch_attention_stream<8>(p, m.B, st)); break;
            default: ASR_TRY(launch_attention_stream<16>(p, m.B, st)); break;
        }
        ASR_CHECK_LAUNCH();
        h->launches++;
        return ASR_OK;
    }
    const size_t smem = sizeof(float) * ((size_t)K * kAtt + (size_t)K * p.sc_ld + (size_t)K * kEnc);
    dim3 grid(m.B, S);
#define ASR_LAUNCH_ATT(KK)                                                                      \
    do {                                                                                        \
        ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attention_kernel<KK>), smem)); \
        attention_kernel<KK><<<grid, 256, smem, st>>>(p);                                       \
    } while (0)
    if (smem > 200 * 1024) { set_error("attention: utterance too long (%d frames)", m.Lmax); return ASR_ERR_CAPACITY; }
    switch (K) {
        case 1: ASR_LAUNCH_ATT(1); break;
        case 4: ASR_LAUNCH_ATT(4); break;
        case 8: ASR_LAUNCH_ATT(8); break;
        default: ASR_LAUNCH_ATT(16); break;
    }
#undef ASR_LAUNCH_ATT
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
// per-utterance merge of the vocabulary GEMM's tile partials + beam bookkeeping
__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

constexpr int kCandCap = 1024;

struct BookParams {
    float* beam_score; int* src_row;            // [R]
    int* tok_hist; int* prev_hist;              // [max_len + 1, R]
    float* fin_score; int* fin_row;             // [max_len, B, k]
    float* tr_cand_s; int* tr_cand_b; int* tr_cand_t;   // [max_len, B, K]
    int* tr_bp; int* tr_tok;                    // [max_len, B, k]
    int* top_done; int* ctrl;
    int B, k, K, step;
    // gather of the next step's cell-GEMM operand [ctx[src] | h[src]] (already split, from the [h | ctx] rows the
    // cell epilogue / attention kernel wrote) - replaces a gather + split launch
    const hi_t* split_hi; const float* split_lo;    // [R, 1024] = [h | ctx]: fp16 hi, 4-byte cross words
    hi_t* next_hi; float* next_lo;                  // [R, 1024] = [ctx | h]
};

struct MergeParams {
    const uint2* part;        // [kVocabTiles, R, KP] (logit bits, token id)
    const float2* part_ms;    // [kVocabSums, R] (tile max, sum of exp(logit - max) over a column half)
    int KP;
    BookParams book;
};

// Beam bookkeeping of one step for utterance u from its merged, sorted top-K candidates (model.py:876-929):
// s_cs[j] = score, s_cf[j] = beam * V + token of rank j.
__device__ void beam_bookkeep(const BookParams& p, int u, const float* s_cs, const int* s_cf) {
    __shared__ int s_src[kMaxBeam];
    const int tid = threadIdx.x;
    const int k = p.k, K = p.K, R = p.B * k;
    if (tid < K) {
        const size_t o = ((size_t)p.step * p.B + u) * K + tid;
        p.tr_cand_s[o] = s_cs[tid];
        p.tr_cand_b[o] = s_cf[tid] / kVocab;
        p.tr_cand_t[o] = s_cf[tid] % kVocab;
    }
    if (tid < 32) {
        // one warp, lane j = rank j of the merged list (K <= 32)
        const int j = tid;
        const bool valid = j < K;
        const int f = valid ? s_cf[j] : 0;
        const int beam = f / kVocab, tok = f - beam * kVocab;
        const bool eos = valid && tok == kEos;
        // finished hypotheses: </s> among the top k (model.py:876-889)
        if (eos && j < k) {
            const size_t o = ((size_t)p.step * p.B + u) * k + j;
            p.fin_score[o] = s_cs[j];
            p.fin_row[o] = u * k + beam;
        }
        if (j == 0 && eos) p.top_done[u] = 1;
        // active set: first k non-EOS candidates in rank order (model.py:904-929)
        const unsigned live = __ballot_sync(0xffffffffu, valid && !eos);
        const int i = __popc(live & ((1u << j) - 1u));
        if (valid && !eos && i < k) {
            const int r = u * k + i;
            s_src[i] = u * k + beam;
            p.src_row[r] = u * k + beam;
            p.beam_score[r] = s_cs[j];
            p.tok_hist[(size_t)(p.step + 1) * R + r] = tok;
            p.prev_hist[(size_t)(p.step + 1) * R + r] = u * k + beam;
            const size_t o = ((size_t)p.step * p.B + u) * k + i;
            p.tr_bp[o] = beam;
            p.tr_tok[o] = tok;
        }
        // early stop (model.py:897-901): decided by the last utterance to arrive
        __threadfence();
        __syncwarp();
        if (j == 0) {
            const int t = atomicAdd(p.ctrl + 2, 1);
            if (t == p.B - 1) {
                __threadfence();
                int all = 1;
                for (int b = 0; b < p.B; ++b) all &= (*((volatile int*)p.top_done + b) != 0);
                p.ctrl[2] = 0;
                p.ctrl[3] = p.step + 1;
                if (all) p.ctrl[0] = p.step;
            }
        }
    }
    if (p.next_hi && k > 1) {
        __syncthreads();
        // k rows x 1024 values x (2-byte hi, 4-byte cross): 16-byte copies, halves swapped ([h | ctx] -> [ctx | h])
        static_assert(kProjK / 4 == 256, "one 16-byte cross column per thread");
        // thread `tid` owns one 16-byte column of the cross operand of every row (4 values of k) and, for
        // tid < 128, one 16-byte column of the hi operand (8 values of k); all loads of a group of 4 rows
        // are issued before the first store
        const int c4 = tid;
        const int s4 = c4 < kEnc / 4 ? c4 + kDecH / 4 : c4 - kEnc / 4;
        const bool has_hi = tid < kProjK / 8;
        const int c8 = tid & (kProjK / 8 - 1);
        const int s8 = c8 < kEnc / 8 ? c8 + kDecH / 8 : c8 - kEnc / 8;
        for (int i0 = 0; i0 < k; i0 += 4) {          // up to 8 loads in flight per thread
            uint4 vh[4];
            float4 vl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i0 + i < k) {
                    const size_t so = (size_t)s_src[i0 + i] * kProjK;
                    vl[i] = __ldcg(reinterpret_cast<const float4*>(p.split_lo + so + 4 * s4));
                    if (has_hi) vh[i] = __ldcg(reinterpret_cast<const uint4*>(p.split_hi + so + 8 * s8));
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i0 + i < k) {
                    const size_t d = (size_t)(u * k + i0 + i) * kProjK;
                    *reinterpret_cast<float4*>(p.next_lo + d + 4 * c4) = vl[i];
                    if (has_hi) *reinterpret_cast<uint4*>(p.next_hi + d + 8 * c8) = vh[i];
                }
            }
        }
    }
}

// One CTA per utterance.  The vocabulary GEMM's epilogue left, per 224-column tile and row, the tile's
// (max, sum exp) and its KP >= 2k largest logits with token ids; the top 2k of the utterance's k * V scores
// (model.py:860-867) are among those k * 23 * KP candidates: a logit outside its tile's top KP >= 2k has 2k
// larger-or-equal ones in its own row, and x -> (x - lse) + beam score is monotone.  Order: score descending,
// then beam * V + token ascending (torch.topk's tie order is unspecified; the oracle defines the same).
__global__ void __launch_bounds__(256)
beam_merge_kernel(MergeParams p) {
    griddep_launch_dependents();
    const BookParams& b = p.book;
    griddep_wait();                                        // the vocabulary GEMM has finished
    if (b.ctrl[0] >= 0) return;
    extern __shared__ __align__(16) uint8_t merge_smem[];
    __shared__ float s_lse[kMaxBeam], s_bs[kMaxBeam];
    __shared__ float s_gmax[64];
    __shared__ float s_thr;
    __shared__ int s_cnt;
    __shared__ int s_wcnt[8];
    __shared__ float s_cs[kCandCap];
    __shared__ int s_cf[kCandCap];
    __shared__ float s_top_s[32];
    __shared__ int s_top_f[32];
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = b.k, K = b.K, R = b.B * k, KP = p.KP;
    const int nrow = b.step == 0 ? 1 : k;                  // step 0: only the first beam (model.py:862)
    const int total = nrow * kVocabTiles * KP;
    float* s_lp = reinterpret_cast<float*>(merge_smem);    // [total] log-prob + beam score
    int* s_fl = reinterpret_cast<int*>(s_lp + total);      // [total] beam * V + token

    // log-sum-exp of every row from its tile partials (model.py:835)
    for (int r = warp; r < nrow; r += 8) {
        const int row = u * k + r;
        static_assert(kVocabSums <= 64, "two partials per lane");
        const float2 m0 = __ldcg(p.part_ms + (size_t)lane * R + row);
        const float2 m1 = lane + 32 < kVocabSums ? __ldcg(p.part_ms + (size_t)(lane + 32) * R + row) : make_float2(-CUDART_INF_F, 0.f);
        const float mx = warp_max(fmaxf(m0.x, m1.x));
        const float sum = warp_sum(m0.y * __expf(m0.x - mx) + m1.y * __expf(m1.x - mx));     // exp(-inf) = 0
        if (lane == 0) { s_lse[r] = mx + logf(sum); s_bs[r] = b.beam_score[row]; }
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // candidates: i = (tile * nrow + r) * KP + j, so a warp reads consecutive slots of consecutive rows
    float tmf = -CUDART_INF_F;
    for (int i = tid; i < total; i += 256) {
        const int j = i % KP, q = i / KP;
        const int r = q % nrow, tile = q / nrow;
        const uint2 c = __ldcg(p.part + ((size_t)tile * R + u * k + r) * KP + j);
        const float lp = __fadd_rn(__fsub_rn(__uint_as_float(c.x), s_lse[r]), s_bs[r]);      // model.py:835-836
        s_lp[i] = lp;
        s_fl[i] = r * kVocab + (int)c.y;
        tmf = fmaxf(tmf, lp);
    }
    // Threshold: the maxima of 64 groups of 4 threads are 64 distinct candidates (at least 23 K / 4 >= K groups
    // are non-empty), so their K-th largest T has at least K candidates >= T and usually few more.
    float gm = fmaxf(tmf, __shfl_xor_sync(0xffffffffu, tmf, 1));
    gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
    if ((tid & 3) == 0) s_gmax[tid >> 2] = gm;
    __syncthreads();
    {
        const int g = tid >> 2, part = tid & 3;
        int above = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int o = part * 16 + i;
            const float x = s_gmax[o];
            above += (x > gm || (x == gm && o < g)) ? 1 : 0;
        }
        above += __shfl_xor_sync(0xffffffffu, above, 1);
        above += __shfl_xor_sync(0xffffffffu, above, 2);
        if (part == 0 && above == K - 1) s_thr = gm;
    }
    __syncthreads();
    unsigned kth = ordered_key(s_thr);
    auto collect = [&]() {
        for (int i = tid; i < total; i += 256) {
            const float v = s_lp[i];
            if (ordered_key(v) >= kth) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < kCandCap) { s_cs[slot] = v; s_cf[slot] = s_fl[i]; }
            }
        }
    };
    collect();
    __syncthreads();
    if (s_cnt > kCandCap) {
        // Degenerate scores (more than kCandCap candidates above the threshold): exact K-th largest by
        // bisection over all candidates, then collect again (ties beyond the cap are truncated).
        kth = 0u;
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned cand = kth | (1u << bit);
            int c = 0;
            for (int i = tid; i < total; i += 256) c += ordered_key(s_lp[i]) >= cand ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            __syncthreads();
            if (lane == 0) s_wcnt[warp] = c;
            __syncthreads();
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_wcnt[w];
            if (tot >= K) kth = cand;
        }
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        collect();
        __syncthreads();
    }
    const int n = min(s_cnt, kCandCap);
    // rank by counting: order (score desc, beam * V + token asc)
    for (int i = tid; i < n; i += 256) {
        const float si = s_cs[i];
        const int fi = s_cf[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = s_cs[j];
            rank += (sj > si || (sj == si && s_cf[j] < fi)) ? 1 : 0;
        }
        if (rank < K) { s_top_s[rank] = si; s_top_f[rank] = fi; }
    }
    __syncthreads();
    beam_bookkeep(b, u, s_top_s, s_top_f);
}

int launch_beam_merge(asr_handle* h, int k, int step, cudaStream_t st) {
    Workspace& w = h->ws;
    const int KP = vocab_topk_slots(k);
    MergeParams p{w.topk_part, w.topk_ms, KP,
                  BookParams{w.beam_score, w.src_row, w.tok_hist, w.prev_hist, w.fin_score, w.fin_row, w.tr_cand_s,
                             w.tr_cand_b, w.tr_cand_t, w.tr_bp, w.tr_tok, w.top_done, w.ctrl, h->meta.B, k, 2 * k, step,
                             w.dec_split_hi, w.dec_split_lo, w.a_hi, w.a_lo}};
    const size_t smem = (size_t)k * kVocabTiles * KP * 8;          // k = 16: 94 KB
    ASR_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&beam_merge_kernel), smem));
    ASR_CUDA(launch_kernel(beam_merge_kernel, dim3(h->meta.B), dim3(256), smem, st, true, p));
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
// n-gram LM on device (tables from asr_set_lm; semantics of oracle NGramLM / kenlm .score)
struct LmDev {
    const float* uni_logp; const float* uni_bo;
    const long long* bi_keys; const float* bi_vals; long long bi_cap;
    const long long* tri_keys; const float* tri_vals; long long tri_cap;
    int vocab, skip_id;
    const int* id_map;
};

__device__ __forceinline__ long long hash_slot(long long key, long long cap) {
    unsigned long long z = (unsigned long long)key + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (long long)(z & (unsigned long long)(cap - 1));
}
__device__ long long lm_find(const long long* keys, long long cap, long long key) {
    long long s = hash_slot(key, cap);
    for (long long probes = 0; probes < cap; ++probes) {    // bounded: a corrupt table cannot hang
        const long long kk = keys[s];
        if (kk == key) return s;
        if (kk == -1) return -1;
        s = (s + 1) & (cap - 1);
    }
    return -1;
}
__device__ float lm_uni_or_bi(const LmDev& lm, int c, int w) {
    const long long s = lm_find(lm.bi_keys, lm.bi_cap, (long long)c * lm.vocab + w);
    if (s >= 0) return lm.bi_vals[2 * s];
    return __fadd_rn(lm.uni_bo[c], lm.uni_logp[w]);
}
__device__ float lm_word(const LmDev& lm, int c0, int c1, int nctx, int w) {
    if (nctx == 2) {
        const long long s = lm_find(lm.tri_keys, lm.tri_cap, ((long long)c0 * lm.vocab + c1) * lm.vocab + w);
        if (s >= 0) return lm.tri_vals[s];
        const long long b = lm_find(lm.bi_keys, lm.bi_cap, (long long)c0 * lm.vocab + c1);
        const float bo = b >= 0 ? lm.bi_vals[2 * b + 1] : 0.f;
        return __fadd_rn(bo, lm_uni_or_bi(lm, c1, w));
    }
    if (nctx == 1) return lm_uni_or_bi(lm, c1, w);
    return lm.uni_logp[w];
}
// total log10 probability of ids[0..n) with <s> context and </s> appended, float32 accumulation
__device__ float lm_score_ids(const LmDev& lm, const int* ids, int n) {
    int c0 = 0, c1 = kSos, nctx = 1;
    float total = 0.f;
    for (int i = 0; i <= n; ++i) {
        int w = i < n ? ids[i] : kEos;
        if (i < n && w == lm.skip_id) continue;
        if (lm.id_map) w = lm.id_map[w];
        total = __fadd_rn(total, lm_word(lm, c0, c1, nctx, w));
        c0 = c1; c1 = w; nctx = 2;
    }
    return total;
}

__global__ void lm_score_kernel(LmDev lm, const int* __restrict__ ids, const int* __restrict__ n,
                                int count, int max_n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = lm_score_ids(lm, ids + (size_t)i * max_n, n[i]);
}

static LmDev lm_dev(const asr_handle* h) {
    const LmTables& t = h->lm;
    return LmDev{t.uni_logp, t.uni_bo, t.bi_keys, t.bi_vals, t.bi_cap, t.tri_keys, t.tri_vals,
                 t.tri_cap, t.vocab, t.skip_id, t.id_map};
}

int launch_lm_score(asr_handle* h, const int* d_ids, const int* d_n, int n, int max_n,
                    float* d_scores, cudaStream_t st) {
    lm_score_kernel<<<(n + 127) / 128, 128, 0, st>>>(lm_dev(h), d_ids, d_n, n, max_n, d_scores);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
// finalisation
struct FinalParams {
    const float* fin_score; const int* fin_row;   // [max_len, B, k]
    const int* tok_hist; const int* prev_hist;    // [max_len + 1, R]
    const float* beam_score;                      // [R]
    const int* ctrl; const int* order;
    int* out_tokens; int* out_len; float* out_score; int* out_info;
    LmDev lm;
    int B, k, max_len, second_pass;
    double lm_weight, length_weight;
};

constexpr int kMaxDecodeLen = 64;

__device__ int backtrace(const FinalParams& p, int row, int len, int* toks) {
    // tokens at history positions 1..len along the back-pointer chain ending in `row` at `len`
    const int R = p.B * p.k;
    for (int pos = len; pos >= 1; --pos) {
        toks[pos - 1] = p.tok_hist[(size_t)pos * R + row];
        row = p.prev_hist[(size_t)pos * R + row];
    }
    return len;
}

__global__ void __launch_bounds__(256)
beam_finalise_kernel(FinalParams p) {
    __shared__ double s_best[256];
    __shared__ int s_idx[256];
    __shared__ int s_cnt[256];
    const int u = blockIdx.x, tid = threadIdx.x;
    const int k = p.k;
    const int steps = p.ctrl[3];
    const int n = steps * k;                      // entry e = l*k + j, (step, rank) order
    int cnt = 0;
    for (int e = tid; e < n; e += 256) {
        const int l = e / k, j = e - l * k;
        cnt += p.fin_row[((size_t)l * p.B + u) * k + j] >= 0 ? 1 : 0;
    }
    s_cnt[tid] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) s_cnt[tid] += s_cnt[tid + o]; __syncthreads(); }
    const int nfin = s_cnt[0];
    __syncthreads();

    int toks[kMaxDecodeLen];
    double best = -CUDART_INF;
    int best_e = 0x7fffffff;
    if (nfin > 0) {
        for (int e = tid; e < n; e += 256) {
            const int l = e / k, j = e - l * k;
            const size_t o = ((size_t)l * p.B + u) * k + j;
            const int row = p.fin_row[o];
            if (row < 0) continue;
            double sc = (double)p.fin_score[o];
            if (p.second_pass && nfin > 1) {
                backtrace(p, row, l, toks);
                const float lm = lm_score_ids(p.lm, toks, l);
                sc = __dadd_rn(__dadd_rn(sc, __dmul_rn(p.lm_weight, (double)lm)),
                               __dmul_rn(p.length_weight, (double)l));      // model.py:758-759
            }
            if (sc > best || (sc == best && e < best_e)) { best = sc; best_e = e; }
        }
    }
    s_best[tid] = best;
    s_idx[tid] = best_e;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            const double b2 = s_best[tid + o];
            const int e2 = s_idx[tid + o];
            if (b2 > s_best[tid] || (b2 == s_best[tid] && e2 < s_idx[tid])) { s_best[tid] = b2; s_idx[tid] = e2; }
        }
        __syncthreads();
    }
    if (tid != 0) return;
    const int uo = p.order[u];
    int* ot = p.out_tokens + (size_t)uo * p.max_len;
    int len = 0;
    float score;
    if (nfin > 0) {
        const int e = s_idx[0];
        const int l = e / k, j = e - l * k;
        const size_t o = ((size_t)l * p.B + u) * k + j;
        len = backtrace(p, p.fin_row[o], l, toks);
        score = p.fin_score[o];                    // the un-rescored log-prob (model.py:762)
        atomicAdd(p.out_info + 3, nfin);
    } else {
        // un-finished fallback (model.py:961-972): best active beam with the length bonus
        const float bonus = (float)(p.length_weight * (double)steps);
        int bi = 0;
        float bv = -CUDART_INF_F;
        for (int i = 0; i < k; ++i) {
            const float v = __fadd_rn(p.beam_score[u * k + i], bonus);
            if (v > bv) { bv = v; bi = i; }
        }
        len = backtrace(p, u * k + bi, steps, toks);
        score = bv;
        atomicAdd(p.out_info + 2, 1);
    }
    for (int i = 0; i < p.max_len; ++i) ot[i] = i < len ? toks[i] : kPad;
    p.out_len[uo] = len;
    p.out_score[uo] = score;
    if (u == 0) { p.out_info[0] = steps; p.out_info[1] = p.ctrl[0]; }
}

int launch_beam_finalise(asr_handle* h, int k, int max_len, int second_pass, double lm_weight,
                         double length_weight, cudaStream_t st) {
    Workspace& w = h->ws;
    if (max_len > kMaxDecodeLen) { set_error("max_len %d > %d", max_len, kMaxDecodeLen); return ASR_ERR_ARG; }
    ASR_CUDA(cudaMemsetAsync(w.out_info, 0, 4 * sizeof(int), st));
    FinalParams p{w.fin_score, w.fin_row, w.tok_hist, w.prev_hist, w.beam_score, w.ctrl,
                  h->meta.d_order, w.out_tokens, w.out_len, w.out_score, w.out_info, lm_dev(h),
                  h->meta.B, k, max_len, second_pass, lm_weight, length_weight};
    beam_finalise_kernel<<<h->meta.B, 256, 0, st>>>(p);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

// ---------------------------------------------------------------------------------------------
// greedy
constexpr int kRowElems = 20;   // 256 threads * 20 >= 5004

struct GreedyParams {
    const float* logits;      // [B, V] (export path) or nullptr
    const uint2* part;        // [kVocabTiles, B, 2] tile partials of the vocabulary GEMM (KP = 2)
    const float2* part_ms;    // [kVocabSums, B]
    int* tok_hist;            // [max_len + 1, B]
    int* g_tokens;            // [max_len, B]
    float* g_accum; int* g_finished; int* g_len;
    int* ctrl;
    int B, step;
};

__device__ __forceinline__ float block_reduce_sum(float v, float* s_red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r += s_red[i];
    __syncthreads();
    return r;
}

// argmax (first maximum = lowest token id), log-sum-exp and the score / length bookkeeping of one greedy step
// (model.py:554-578).  FROM_LOGITS: over the materialised row (the driver's logits export); otherwise over the
// vocabulary GEMM's tile partials - slot 0 of a tile is its maximum with the lowest token id.
template <bool FROM_LOGITS>
__global__ void __launch_bounds__(FROM_LOGITS ? 256 : 32)
greedy_pick_kernel(GreedyParams p) {
    if (p.ctrl[0] >= 0) return;
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    float m = -CUDART_INF_F, lse;
    int mi = 0x7fffffff;
    if (FROM_LOGITS) {
        __shared__ float s_red[8];
        __shared__ float s_v[8];
        __shared__ int s_i[8];
        const int warp = tid >> 5;
        const float* row = p.logits + (size_t)u * kVocab;
        float v[kRowElems];
#pragma unroll
        for (int i = 0; i < kRowElems; ++i) {
            const int e = tid + 256 * i;
            v[i] = e < kVocab ? row[e] : -CUDART_INF_F;
            if (v[i] > m) { m = v[i]; mi = e; }          // ascending e per thread: first max kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, mi, o);
            if (m2 > m || (m2 == m && i2 < mi)) { m = m2; mi = i2; }
        }
        if (lane == 0) { s_v[warp] = m; s_i[warp] = mi; }
        __syncthreads();
        m = s_v[0]; mi = s_i[0];
#pragma unroll
        for (int i = 1; i < 8; ++i)
            if (s_v[i] > m || (s_v[i] == m && s_i[i] < mi)) { m = s_v[i]; mi = s_i[i]; }
        __syncthreads();
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kRowElems; ++i) s += expf(v[i] - m);
        s = block_reduce_sum(s, s_red);
        lse = m + logf(s);
    } else {
        const float2 m0 = __ldcg(p.part_ms + (size_t)lane * p.B + u);
        const float2 m1 = lane + 32 < kVocabSums ? __ldcg(p.part_ms + (size_t)(lane + 32) * p.B + u) : make_float2(-CUDART_INF_F, 0.f);
        if (lane < kVocabTiles) {
            const uint2 c = __ldcg(p.part + ((size_t)lane * p.B + u) * 2);
            m = __uint_as_float(c.x);
            mi = (int)c.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, mi, o);
            if (m2 > m || (m2 == m && i2 < mi)) { m = m2; mi = i2; }
        }
        const float s = warp_sum(m0.y * __expf(m0.x - m) + m1.y * __expf(m1.x - m));
        lse = m + logf(s);
    }
    if (tid != 0) return;
    const float logp = __fsub_rn(m, lse);            // model.py:554,563
    const int tok = mi;
    p.g_tokens[(size_t)p.step * p.B + u] = tok;
    p.tok_hist[(size_t)(p.step + 1) * p.B + u] = tok;
    int fin = p.g_finished[u];
    float acc = p.g_accum[u];
    const int now = tok == kEos;
    if (!fin && now) acc = __fadd_rn(acc, logp);     // model.py:569
    fin |= now;
    if (!fin) { p.g_len[u] += 1; acc = __fadd_rn(acc, logp); }   // model.py:572-575
    p.g_finished[u] = fin;
    p.g_accum[u] = acc;
    __threadfence();
    const int t = atomicAdd(p.ctrl + 2, 1);
    if (t == p.B - 1) {
        __threadfence();
        int all = 1;
        for (int b = 0; b < p.B; ++b) all &= (*((volatile int*)p.g_finished + b) != 0);
        p.ctrl[2] = 0;
        p.ctrl[3] = p.step + 1;
        if (all) p.ctrl[0] = p.step;                 // model.py:578
    }
}

int launch_greedy_pick(asr_handle* h, int step, bool from_logits, cudaStream_t st) {
    Workspace& w = h->ws;
    GreedyParams p{from_logits ? w.logits : nullptr, w.topk_part, w.topk_ms, w.tok_hist, w.g_tokens, w.g_accum,
                   w.g_finished, w.g_len, w.ctrl, h->meta.B, step};
    if (from_logits) greedy_pick_kernel<true><<<h->meta.B, 256, 0, st>>>(p);
    else greedy_pick_kernel<false><<<h->meta.B, 32, 0, st>>>(p);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

__global__ void greedy_finalise_kernel(const int* __restrict__ g_tokens, const int* __restrict__ g_len,
                                       const float* __restrict__ g_accum,
                                       const int* __restrict__ g_finished,
                                       const int* __restrict__ order, const int* __restrict__ ctrl,
                                       int B, int max_len, int* out_tokens, int* out_len,
                                       float* out_score, int* out_fin, int* out_info) {
    const int u = blockIdx.x;
    const int uo = order[u];
    const int steps = ctrl[3];
    for (int s = threadIdx.x; s < max_len; s += blockDim.x)
        out_tokens[(size_t)uo * max_len + s] = s < steps ? g_tokens[(size_t)s * B + u] : kPad;
    if (threadIdx.x == 0) {
        out_len[uo] = g_len[u];
        out_score[uo] = g_accum[u];
        out_fin[uo] = g_finished[u];
        if (u == 0) { out_info[0] = steps; out_info[1] = ctrl[0]; out_info[2] = 0; out_info[3] = 0; }
    }
}

int launch_greedy_finalise(asr_handle* h, int max_len, cudaStream_t st) {
    Workspace& w = h->ws;
    greedy_finalise_kernel<<<h->meta.B, 64, 0, st>>>(w.g_tokens, w.g_len, w.g_accum, w.g_finished,
                                                     h->meta.d_order, w.ctrl, h->meta.B, max_len,
                                                     w.out_tokens, w.out_len, w.out_score,
                                                     w.top_done, w.out_info);
    ASR_CHECK_LAUNCH();
    h->launches++;
    return ASR_OK;
}

}  // namespace asr
