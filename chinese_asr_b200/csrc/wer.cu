// Character error rate on the device: get_wer (reference util.py:237-262, python-Levenshtein
// `distance(pred, ref)` over the two STRINGS) as used by the decode drivers when `text` is given
// (model.py:595-598, 982-985).  The hypotheses of a batch are still on the device when a decode
// call returns, so the distance is computed there and only [B] integers come back.
//
// Strings are sequences of Unicode code points: token t expands to vocab_cp[vocab_off[t] ..
// vocab_off[t+1]) ("<unk>" is five characters, like ''.join(int2word[...]) makes it).
// One CTA per utterance; the DP table is swept by anti-diagonals (cells of a diagonal are
// independent), three diagonals live in shared memory.
#include <algorithm>
#include <vector>

#include "asr_internal.cuh"

namespace asr {

constexpr int kWerThreads = 128;
constexpr int kWerMaxHyp = 1024;      // code points of one hypothesis (max_len * longest token)

__global__ void __launch_bounds__(kWerThreads)
edit_distance_kernel(const int* __restrict__ hyp_tok, const int* __restrict__ hyp_len, int hyp_ld,
                     const int* __restrict__ vocab_cp, const int* __restrict__ vocab_off, int V,
                     const int* __restrict__ ref_cp, const long long* __restrict__ ref_off,
                     int* __restrict__ dist, int* __restrict__ hyp_chars) {
    __shared__ int s_hyp[kWerMaxHyp];
    __shared__ int s_diag[3][kWerMaxHyp + 1];
    __shared__ int s_n;
    const int u = blockIdx.x, tid = threadIdx.x;
    const int* ref = ref_cp + ref_off[u];
    const int m = (int)(ref_off[u + 1] - ref_off[u]);

    // expand the hypothesis tokens to code points (serial prefix: at most hyp_ld tokens)
    if (tid == 0) {
        int n = 0;
        const int nt = min(hyp_len[u], hyp_ld);
        for (int i = 0; i < nt; ++i) {
            const int t = hyp_tok[(size_t)u * hyp_ld + i];
            if (t < 0 || t >= V) continue;
            for (int c = vocab_off[t]; c < vocab_off[t + 1] && n < kWerMaxHyp; ++c) s_hyp[n++] = vocab_cp[c];
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    if (hyp_chars && tid == 0) hyp_chars[u] = n;
    if (n == 0 || m == 0) {
        if (tid == 0) dist[u] = max(n, m);
        return;
    }
    // D[i][j], i over the hypothesis (0..n), j over the reference (0..m); diagonal d = i + j holds D[i][d-i] at [i]
    for (int d = 0; d <= n + m; ++d) {
        int* cur = s_diag[d % 3];
        const int* p1 = s_diag[(d + 2) % 3];     // diagonal d - 1
        const int* p2 = s_diag[(d + 1) % 3];     // diagonal d - 2
        const int lo = max(0, d - m), hi = min(n, d);
        for (int i = lo + tid; i <= hi; i += kWerThreads) {
            const int j = d - i;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else v = min(min(p1[i - 1] + 1, p1[i] + 1), p2[i - 1] + (s_hyp[i - 1] != ref[j - 1] ? 1 : 0));
            cur[i] = v;
        }
        __syncthreads();
    }
    if (tid == 0) dist[u] = s_diag[(n + m) % 3][n];
}

}  // namespace asr

using namespace asr;

extern "C" {

int asr_set_vocab(asr_handle* h, const int32_t* h_codepoints, const int32_t* h_tok_off, int V) {
    if (!h || !h_codepoints || !h_tok_off || V <= 0) { set_error("asr_set_vocab: bad argument"); return ASR_ERR_ARG; }
    if (h_tok_off[0] != 0) { set_error("asr_set_vocab: offsets must start at 0"); return ASR_ERR_ARG; }
    int maxc = 0;
    for (int t = 0; t < V; ++t) {
        if (h_tok_off[t + 1] < h_tok_off[t]) { set_error("asr_set_vocab: offsets not monotone at %d", t); return ASR_ERR_ARG; }
        maxc = std::max(maxc, h_tok_off[t + 1] - h_tok_off[t]);
    }
    VocabChars& vc = h->vocab;
    if (vc.d_cp) { cudaFree(vc.d_cp); vc.d_cp = nullptr; }
    if (vc.d_off) { cudaFree(vc.d_off); vc.d_off = nullptr; }
    vc.V = V;
    vc.max_chars = maxc;
    ASR_CUDA(cudaMalloc(&vc.d_cp, sizeof(int) * std::max<size_t>(1, (size_t)h_tok_off[V])));
    ASR_CUDA(cudaMalloc(&vc.d_off, sizeof(int) * (V + 1)));
    ASR_CUDA(cudaMemcpy(vc.d_cp, h_codepoints, sizeof(int) * (size_t)h_tok_off[V], cudaMemcpyHostToDevice));
    ASR_CUDA(cudaMemcpy(vc.d_off, h_tok_off, sizeof(int) * (V + 1), cudaMemcpyHostToDevice));
    return ASR_OK;
}

int asr_wer(asr_handle* h, const int32_t* h_hyp, const int32_t* h_hyp_len, int hyp_ld, const int32_t* h_ref,
            const int64_t* h_ref_off, int B, int32_t* h_dist, int32_t* h_hyp_chars, void* stream) {
    if (!h || !h_ref || !h_ref_off || !h_dist || B <= 0) { set_error("asr_wer: bad argument"); return ASR_ERR_ARG; }
    VocabChars& vc = h->vocab;
    if (!vc.d_cp) { set_error("asr_wer: call asr_set_vocab first"); return ASR_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int* d_hyp = nullptr;
    const int* d_len = nullptr;
    int ld = hyp_ld;
    int* d_tmp = nullptr;
    if (h_hyp) {
        if (!h_hyp_len || hyp_ld <= 0) { set_error("asr_wer: hypotheses need lengths and a row stride"); return ASR_ERR_ARG; }
    } else {
        // hypotheses of the last decode call on this handle, still resident (original utterance order)
        if (!h->encoded || h->last_B != B || h->last_out_ld <= 0) {
            set_error("asr_wer: no resident hypotheses for a batch of %d (last decode: %d)", B, h->last_B);
            return ASR_ERR_STATE;
        }
        d_hyp = h->ws.out_tokens;
        d_len = h->ws.out_len;
        ld = h->last_out_ld;
    }
    if ((int64_t)ld * std::max(1, vc.max_chars) > kWerMaxHyp) {
        set_error("asr_wer: hypotheses of up to %d x %d characters exceed %d", ld, vc.max_chars, kWerMaxHyp);
        return ASR_ERR_CAPACITY;
    }
    // reference strings arrive as code points (the caller owns the id -> string mapping of its transcripts)
    const int64_t r0 = h_ref_off[0];
    std::vector<long long> roff(B + 1, 0);
    for (int u = 0; u < B; ++u) {
        if (h_ref_off[u + 1] < h_ref_off[u]) { set_error("asr_wer: reference offsets not monotone at %d", u); return ASR_ERR_ARG; }
        roff[u + 1] = h_ref_off[u + 1] - r0;
    }
    const size_t n_ref = (size_t)roff[B];
    const size_t hyp_ints = h_hyp ? (size_t)B * ld + B : 0;
    const size_t bytes = sizeof(long long) * (B + 1) + sizeof(int) * (n_ref + 2 * (size_t)B + hyp_ints + 4);
    char* d_blk = nullptr;
    ASR_CUDA(cudaMalloc(&d_blk, bytes));
    long long* d_roff = reinterpret_cast<long long*>(d_blk);
    int* d_rcp = reinterpret_cast<int*>(d_roff + B + 1);
    int* d_dist = d_rcp + n_ref;
    int* d_hc = d_dist + B;
    d_tmp = d_hc + B;
    int rc = ASR_OK;
    cudaError_t e = cudaMemcpyAsync(d_roff, roff.data(), sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_ref > 0)
        e = cudaMemcpyAsync(d_rcp, h_ref + r0, sizeof(int) * n_ref, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && h_hyp) {
        e = cudaMemcpyAsync(d_tmp, h_hyp, sizeof(int) * (size_t)B * ld, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tmp + (size_t)B * ld, h_hyp_len, sizeof(int) * B, cudaMemcpyHostToDevice, st);
        d_hyp = d_tmp;
        d_len = d_tmp + (size_t)B * ld;
    }
    if (e == cudaSuccess) {
        edit_distance_kernel<<<B, kWerThreads, 0, st>>>(d_hyp, d_len, ld, vc.d_cp, vc.d_off, vc.V, d_rcp, d_roff,
                                                         d_dist, d_hc);
        e = cudaGetLastError();
        h->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_dist, d_dist, sizeof(int) * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && h_hyp_chars) e = cudaMemcpyAsync(h_hyp_chars, d_hc, sizeof(int) * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // roff (pageable) must stay alive until here
    if (e != cudaSuccess) { set_error("asr_wer: %s", cudaGetErrorString(e)); rc = ASR_ERR_CUDA; }
    cudaFree(d_blk);
    return rc;
}

}  // extern "C"
