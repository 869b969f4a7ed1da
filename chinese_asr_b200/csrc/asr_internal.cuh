// Internal declarations shared by the translation units of libasr_b200.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/asr_b200.h"

namespace asr {

constexpr int kFeat = ASR_FEAT_DIM;     // 720
constexpr int kMel = ASR_N_MELS;        // 80
constexpr int kEncH = 256;              // per direction
constexpr int kEnc = ASR_ENC_OUT;       // 512
constexpr int kGates = 1024;            // 4 * kEncH, per direction
constexpr int kAtt = ASR_ATT;           // 128
constexpr int kDecH = ASR_DEC_H;        // 512
constexpr int kEmb = ASR_EMB;           // 256
constexpr int kVocab = ASR_VOCAB;       // 5004
constexpr int kDecK = kEmb + kEnc + kDecH;   // 1280: [emb | ctx_prev | h_prev]
constexpr int kProjK = kDecH + kEnc;    // 1024: [h | ctx]
constexpr int kPad = 0, kSos = 1, kEos = 2;
constexpr int kNfft = 512, kHop = 160, kWin = 400, kBins = 257, kWinOff = 56;
constexpr int kSampleRate = 16000;        // gpd['sample_rate']
constexpr int kMaxBeam = ASR_MAX_BEAM;
constexpr int kNumSMs = 148;
constexpr int kStages = 12;       // 8 pipeline stages + kernel-level timers (GEMM kernel, operand split)
// vocabulary projection with fused log-sum-exp / top-k partials (gemm_tc.cu): 224-column tiles, 5004 = 22 x 224 + 76
constexpr int kVocabTileN = 224;
constexpr int kVocabTiles = (kVocab + kVocabTileN - 1) / kVocabTileN;     // 23
constexpr int kVocabSums = 2 * kVocabTiles;      // (max, sum exp) partials per row: two column halves per tile

void set_error(const char* fmt, ...);
// raise a kernel's dynamic shared-memory limit on the current device (once per kernel and device)
int ensure_dynamic_smem(const void* func, size_t bytes);

// ---- split-precision operands of the tcgen05 GEMM engine (gemm_tc.cu) ---------------------------
// x = hi + lo with hi = fp16(x) (11 significant bits, like tf32, but half the bytes and the full-rate
// kind::f16 MMA) and lo = x - hi exact in fp32.  |x| > 65504 saturates hi; the residual carries the rest.
typedef __half hi_t;
#ifdef __CUDACC__
__device__ __forceinline__ float hi_part(float x) {
    return __half2float(__float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)));
}
// two hi parts (already fp16-representable) -> packed half2, first argument in the low half-word
__device__ __forceinline__ uint32_t pack_hi2(float a, float b) {
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
#endif

#ifdef __CUDACC__
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream still runs.  Every kernel of the decoder step calls
// griddep_launch_dependents() first (the next kernel's CTAs may take SMs as soon as this grid's CTAs leave them
// and run their own prologue) and griddep_wait() before it reads or writes ANY global memory (the wait returns
// once the predecessor grid has completed and its writes are visible) - on every path, so that completion stays
// transitive along the chain.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

#define ASR_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            asr::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,              \
                           cudaGetErrorString(e_));                                           \
            return ASR_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ASR_CHECK_LAUNCH()                                                                    \
    do {                                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess) {                                                              \
            asr::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,              \
                           cudaGetErrorString(e_));                                           \
            return ASR_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ASR_TRY(expr)                                                                         \
    do {                                                                                      \
        int rc_ = (expr);                                                                     \
        if (rc_ != ASR_OK) return rc_;                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------
// A operand of the GEMMs: up to three K-segments, each optionally row-gathered.  This is how the
// decoder reads [emb[tok] | ctx[src] | h[src]] without materialising the concatenation or the
// reference's reorder-by-backpointer copies (model.py:913-925).
struct ASeg {
    const float* base;
    const int* rowidx;   // nullptr -> identity
    int ld;              // row stride in floats
    int kend;            // exclusive end of this segment along K
};
struct AOperand {
    ASeg seg[3];
    int nseg;
};
inline AOperand plain_a(const float* a, int ld, int K) {
    AOperand o{};
    o.nseg = 1;
    o.seg[0] = ASeg{a, nullptr, ld, K};
    return o;
}

enum class Epi : int {
    kBias = 0,      // C = acc + bias[n]
    kBiasScale,     // C = (acc + bias[n]) / scale      (logit /= temperature, model.py:834)
    kLstmCell       // N is gate-interleaved (n = 4*u + gate): writes h, c  (nn.LSTMCell)
};

struct GemmEpilogue {
    Epi kind;
    const float* bias;     // [N]
    float* C;              // [M, ldc]        (kBias / kBiasScale)
    int ldc;
    float scale;
    // kLstmCell
    const float* c_prev;   // [*, H] gathered through c_rowidx
    const int* c_rowidx;
    float* h_out;          // [M, H]
    float* c_out;          // [M, H]
    int H;
    const int* stop_flag;  // optional: kernel returns immediately when *stop_flag >= 0
    bool pdl;              // launch with programmatic stream serialization (kernels of the decoder step)
    // tcgen05 engine only:
    int lda, ldw;          // row strides (floats) of the pre-split A / W operands; 0 = K (dense)
    const float* addrow;   // kLstmCell: gate pre-activations += addrow[addrow_idx[row]][n] (the
    const int* addrow_idx; //   pre-multiplied embedding table E' = emb * W_ih[:, :256]^T, row = token)
    int addrow_ld;
    hi_t* split_hi;        // kLstmCell: also write fp16(h) / the cross operand into [M, split_ld] at column u,
    float* split_lo;       //   the A operand of the query / vocabulary GEMMs (no separate split pass)
    int split_ld;
    // vocabulary epilogue (kBias / kBiasScale with topk_slots = KP > 0): nothing is written to C; per 224-column
    // tile tn and row r the KP largest logits (value bits, token id) and the (max, sum of exp) of the tile
    int topk_slots;
    uint2* topk_part;      // [kVocabTiles, M, KP]
    float2* topk_ms;       // [kVocabSums, M]
};

// C[M,N] = A[M,K] * W[N,K]^T (+ epilogue) on the tcgen05 split-precision engine (gemm_tc.cu): operands
// pre-split into an fp16 `hi` part ([rows, K] halves) and a `lo` buffer ([rows, K] 4-byte words) holding the
// bf16 cross-term operand (layout: see split_operand_kernel).
enum { kSplitAct = 1, kSplitWeight = 2 };
int split_operand(const AOperand& A, int M, int K, hi_t* hi, float* lo, const int* stop_flag,
                  cudaStream_t st, int64_t* launches, int fmt = kSplitAct);
int launch_gemm_tc(const hi_t* a_hi, const float* a_lo, const hi_t* w_hi, const float* w_lo, int M,
                   int N, int K, const GemmEpilogue& epi, cudaStream_t st, int64_t* launches);
int vocab_topk_slots(int k);      // KP of the vocabulary epilogue for beam width k (>= 2k): 2, 8, 16 or 32

// ---------------------------------------------------------------------------------------------
struct FeatureConsts {
    float* window = nullptr;     // [400]
    float2* tw512 = nullptr;     // [512] exp(-2 pi i k / 512)
    int* mel_start = nullptr;    // [80]
    int* mel_len = nullptr;      // [80]
    float* mel_w = nullptr;      // [80, mel_maxw]
    int mel_maxw = 0;
    float taps[27];
    float preemph = 0.97f;
};

struct PackedWeights {
    // encoder: per layer, both directions concatenated along N and permuted so that the 128
    // gate pre-activations a recurrence CTA needs are contiguous:
    //   n' = dir*1024 + j*128 + 4*uu + gate   <->  reference row gate*256 + 32*j + uu
    float* enc_w_ih[4] = {};     // [2048, K_l]
    float* enc_bias[4] = {};     // [2048]  b_ih + b_hh, same permutation
    float* enc_w_hh[4] = {};     // [2, 1024, 256] same permutation per direction (split on the fly by encoder_tc3.cu)
    // decoder cell: [2048, 1280] = [W_ih | W_hh], rows interleaved n' = 4*u + gate
    float* dec_w = nullptr;
    float* dec_b = nullptr;      // [2048] b_ih + b_hh interleaved
    float* emb = nullptr;        // [5004, 256]
    float* proj_w = nullptr;     // [5004, 1024]
    float* proj_b = nullptr;     // [5004]
    float* att_w_enc_t = nullptr;    // [128, 512]  (W_enc transposed -> [N, K])
    float* att_b = nullptr;          // [128]
    float* att_w_hidden = nullptr;   // [512, 128]  as stored (k-major rows, coalesced over d)
    float* att_v = nullptr;          // [128]
    // fp16 hi / bf16 cross splits of the GEMM weights (tcgen05 split-precision path), same shapes as the originals
    hi_t* enc_w_ih_hi[4] = {};
    float* enc_w_ih_lo[4] = {};
    hi_t* dec_w_hi = nullptr;
    float* dec_w_lo = nullptr;
    hi_t* proj_w_hi = nullptr;
    float* proj_w_lo = nullptr;
    hi_t* att_w_enc_t_hi = nullptr;
    float* att_w_enc_t_lo = nullptr;
    float* att_w_hidden_t = nullptr;     // [128, 512] (W_hidden transposed -> [N, K]) for the query GEMM
    hi_t* att_w_hidden_t_hi = nullptr;
    float* att_w_hidden_t_lo = nullptr;
    float* zero_bias = nullptr;          // [2048] zeros
    float* emb_proj = nullptr;           // [5004, 2048] E' = embedding * W_ih[:, :256]^T (gate-interleaved columns)
};

// token id -> Unicode code points of int2word[id] (asr_set_vocab), for the device edit distance (wer.cu)
struct VocabChars {
    int* d_cp = nullptr;
    int* d_off = nullptr;
    int V = 0, max_chars = 0;
};

struct LmTables {
    float* uni_logp = nullptr;
    float* uni_bo = nullptr;
    long long* bi_keys = nullptr;
    float* bi_vals = nullptr;
    long long bi_cap = 0;
    long long* tri_keys = nullptr;
    float* tri_vals = nullptr;
    long long tri_cap = 0;
    int vocab = 0;
    int skip_id = -1;
    int* id_map = nullptr;       // [vocab] token -> LM word (tokens without a unigram -> <unk>), nullptr = identity
    bool loaded = false;
};

// Per-batch metadata (host copy + device copy).  Utterances are processed in "sorted order"
// (descending length, like encoder.py:47); `order[r]` is the original index of sorted rank r.
struct BatchMeta {
    int B = 0;
    int Lmax = 0;
    int64_t rows = 0;                 // sum of L
    std::vector<int> order;           // [B] sorted rank -> original index
    std::vector<int> len_sorted;      // [B]
    std::vector<int> toff;            // [Lmax+1] packed time-major offsets (rows with len > t come first)
    std::vector<int> uoff_sorted;     // [B+1] utterance-major offsets in sorted order
    std::vector<int> foff_orig;       // [B+1] utterance-major offsets in original order (d_feats)
    // device mirrors
    int* d_order = nullptr;
    int* d_len_sorted = nullptr;
    int* d_toff = nullptr;
    int* d_uoff_sorted = nullptr;
    int* d_foff_orig = nullptr;
    int* d_pack_src = nullptr;        // [rows] packed row -> row of d_feats (original order)
    int* d_feat2packed = nullptr;     // [rows] row of d_feats -> packed row
};

// Device allocations of a handle.  Every buffer is followed by a guard of kGuardBytes filled with a pattern that
// asr_check_guards() verifies: compute-sanitizer is closed on the B200 pool, so writes past the end of a buffer are
// caught by the library's own canaries (checked by the GPU tests and by smoke()).
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardByte = 0xA5;
struct DevicePool {
    std::vector<void*> ptrs;
    std::vector<size_t> bytes;       // payload bytes of ptrs[i] (the guard follows); 0 = no guard (registered buffer)
    void push_back(void* p) { ptrs.push_back(p); bytes.push_back(0); }
    void clear() { ptrs.clear(); bytes.clear(); }
    std::vector<void*>::const_iterator begin() const { return ptrs.begin(); }
    std::vector<void*>::const_iterator end() const { return ptrs.end(); }
};

struct Workspace {
    int max_utts = 0, max_beam = 0, max_len = 0;
    int64_t max_rows = 0, max_samples = 0, max_frames = 0;
    // features
    float* pcm = nullptr;        // [max_samples] staging for asr_transcribe
    float* pcm_pre[2] = {};      // [max_samples] x 2, allocated on first asr_prefetch_pcm
    float* mel = nullptr;        // [max_frames, 80]
    float* xpack = nullptr;      // [max_rows, 720] packed time-major encoder input
    double* feat_partial = nullptr;  // [max_utts, 4, 720] (sum, sum of squares) partials of the CMVN statistics
    long long* d_pcm_off = nullptr;  // [max_utts + 1]
    int* d_frame_off = nullptr;      // [max_utts + 1]  STFT frames prefix
    int* d_featrow_off = nullptr;    // [max_utts + 1]  feature rows prefix (original order)
    // encoder
    float* xg = nullptr;         // [max_rows, 2048]
    float* act[2] = {};          // [max_rows, 512] packed time-major residual stream
    float* enc = nullptr;        // [max_rows, 512] utterance-major, sorted order
    float* keys = nullptr;       // [max_rows, 128]
    float* keys_exp = nullptr;   // [max_rows, 128] e^(2 key), the per-frame factor of the attention scores
    int* keys_big = nullptr;     // [max_utts] utterance has a key outside the product form's range
    float* h0 = nullptr;         // [max_utts, 512] last-layer final h (fwd|bwd), sorted order
    float* c0 = nullptr;
    // decoder
    float* dh[2] = {};           // [R, 512]
    float* dc[2] = {};
    float* dctx[2] = {};
    float* logits = nullptr;     // [max_utts, 5004]: only the greedy driver's logits export (tests) materialises them
    uint2* topk_part = nullptr;  // [23, R, KP] per vocabulary tile and row: top-KP (logit bits, token id)
    float2* topk_ms = nullptr;   // [46, R] per half vocabulary tile and row: (tile max logit, sum of exp(logit - max))
    float* att_q = nullptr;      // [R, 128] query projection of the current step
    hi_t* dec_split_hi = nullptr;    // [R, 1024] fp16 hi of [h_new | ctx_new], written by the producing kernels
    float* dec_split_lo = nullptr;
    float* att_part = nullptr;   // [B, S, k, 2 + 512] partial (max, sum, ctx)
    float* att_score = nullptr;  // [R, Lmax_cap] raw scores (alignment export)
    int* att_ticket = nullptr;   // [B]
    int* tok_hist = nullptr;     // [max_len + 1, R]
    int* prev_hist = nullptr;    // [max_len + 1, R]
    int* src_row = nullptr;      // [R] source row of the current step's state
    float* beam_score = nullptr; // [R]
    float* fin_score = nullptr;  // [max_len, B, k]  (NaN = none)
    int* fin_row = nullptr;      // [max_len, B, k]
    float* tr_cand_s = nullptr;  // [max_len, B, 2k]
    int* tr_cand_b = nullptr;
    int* tr_cand_t = nullptr;
    int* tr_bp = nullptr;        // [max_len, B, k]
    int* tr_tok = nullptr;       // [max_len, B, k]
    int* top_done = nullptr;     // [B]
    int* ctrl = nullptr;         // [8]: 0 stop_step(-1), 1 done_count, 2 ticket, 3 steps_run
    float* rec_stage = nullptr;  // recurrence: per CTA the exchange image of its new h slice (encoder_tc3.cu)
    size_t rec_stage_ctas = 0;   // capacity in 8 KB units
    // greedy
    float* g_accum = nullptr;    // [B]
    int* g_finished = nullptr;   // [B]
    int* g_len = nullptr;        // [B]
    int* g_tokens = nullptr;     // [max_len, B]
    // results
    int* out_tokens = nullptr;   // [B, max_len]
    int* out_len = nullptr;      // [B]
    float* out_score = nullptr;  // [B]
    int* out_info = nullptr;     // [4]
    int64_t att_score_ld = 0;
    // pinned host staging
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    hi_t* a_hi = nullptr;        // split A operand of the tcgen05 GEMMs: max(rows*720, R*1280) elements
    float* a_lo = nullptr;
    DevicePool allocs;
};

}  // namespace asr

struct asr_handle {
    int device = 0;
    asr::FeatureConsts fc;
    asr::PackedWeights w;
    asr::LmTables lm;
    asr::VocabChars vocab;
    asr::Workspace ws;
    asr::BatchMeta meta;
    bool encoded = false;
    int last_k = 0, last_steps = 0, last_B = 0;
    int last_info[4] = {};       // {steps, stop step or -1, fallbacks, finished hypotheses} of the last decode
    int last_out_ld = 0;         // row stride of ws.out_tokens after the last decode (its max_len)
    int64_t launches = 0;
    int rec_chunks = 7;          // recurrence chunks per direction (clusters = 2 x chunks), asr_set_recurrence_chunks
    bool timing = false;
    bool rec_timeline = false;   // print the recurrence kernel's clock64 step timeline (asr_stage_timing(h, 2))
    cudaGraphExec_t graph_exec = nullptr;   // captured beam-decode loop for the shape in graph_key
    long long graph_key[8] = {};
    long long graph_seen[8] = {};
    int64_t graph_launches = 0;             // kernels inside the captured graph
    cudaStream_t graph_stream = nullptr;    // blocking stream standing in for the legacy stream (not capturable)
    cudaStream_t copy_stream = nullptr;     // asr_prefetch_pcm: H2D copies overlapping the previous batch
    cudaEvent_t pre_ev[2] = {};
    const void* pre_src[2] = {};            // host address a staged copy came from (nullptr = free / consumed)
    int64_t pre_n[2] = {};
    int pre_fmt[2] = {};                    // asr_pcm_format of the staged samples
    bool pre_issued[2] = {};                // the copy has been enqueued (it is issued behind the next batch's uploads)
    uint64_t pre_count = 0;
    bool feat_split_ready = false;  // ws.a_hi / a_lo hold the split of the packed features (written by feat_write_kernel)
    bool enc_split_ready = false;  // ws.a_hi / a_lo hold the split of `enc` (written by the last recurrence)
    double gemm_flops = 0.0;     // algorithmic 2*M*N*K of every GEMM-engine launch since the last reset
    cudaEvent_t ev[2 * 1024] = {};
    int n_ev = 0;
    int ev_stage[1024] = {};
    float stage_ms[asr::kStages] = {};
    asr::DevicePool weight_allocs;
};

namespace asr {

// launch a plain (non-cluster) kernel, optionally with programmatic stream serialization
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- features.cu -----------------------------------------------------------------------------
int build_feature_consts(asr_handle* h, const asr_feature_consts* fc);
int launch_logmel(asr_handle* h, const void* d_pcm, int format, const long long* d_pcm_off,
                  const int* d_frame_off, int B, int max_frames_per_utt, float* d_mel, cudaStream_t st);
// out_rowmap: feature row (utterance-major, original order) -> output row; nullptr = identity
int launch_delta_cmvn(asr_handle* h, const float* d_mel, const int* d_frame_off,
                      const int* d_featrow_off, int B, int max_rows_per_utt, int normalise, float eps,
                      const int* out_rowmap, float* d_out, cudaStream_t st, hi_t* split_hi = nullptr,
                      float* split_x = nullptr);   // optional: also (or only, d_out = nullptr) the split GEMM operand
// AudioLoader.batch_audio (data.py:513-518) on features that already exist
int launch_cmvn(asr_handle* h, const float* d_in, const int* d_featrow_off, int B, int max_rows_per_utt,
                float eps, float* d_out, cudaStream_t st);

// ---- encoder.cu ------------------------------------------------------------------------------
int launch_pack_rows(asr_handle* h, const float* src, const int* rowmap, int64_t rows, int width,
                     float* dst, cudaStream_t st);
// One bidirectional layer recurrence on the tensor cores, weights TMEM-resident (encoder_tc3.cu).
// xg [rows, 2048]; x_in residual input (nullptr for layer 0); y_packed: packed time-major output (or nullptr);
// y_utt: utterance-major output (or nullptr).  split_hi / split_lo (optional, [rows, 512]): the layer output is
// also written as the pre-split A operand of the GEMM that consumes it (kSplitAct layout, same row order as
// y_packed - or as y_utt when y_packed is nullptr)
int launch_lstm_recurrence_tc3(asr_handle* h, int layer, const float* xg, const float* x_in,
                               float* y_packed, float* y_utt, float* h_fin, float* c_fin,
                               cudaStream_t st, hi_t* split_hi = nullptr, float* split_lo = nullptr);
size_t rec3_stage_bytes_per_cta();
int launch_export_padded(asr_handle* h, const float* src_utt, int width, float* dst, int Lmax, int B,
                         const float* pad_row, cudaStream_t st);
int launch_export_packed_padded(asr_handle* h, const float* src_packed, int width, float* dst,
                                cudaStream_t st);
int launch_unsort_rows(asr_handle* h, const float* src, int width, float* dst, cudaStream_t st);

// ---- decoder.cu ------------------------------------------------------------------------------
int decode_init(asr_handle* h, int k, int max_len, bool greedy, cudaStream_t st);
int launch_keys_exp(asr_handle* h, cudaStream_t st);
int launch_attention(asr_handle* h, int k, int step, int cur, float* d_align_step, cudaStream_t st);
// per utterance: log-sum-exp and top-2k over the k rows' vocabulary-tile partials, then the beam bookkeeping
int launch_beam_merge(asr_handle* h, int k, int step, cudaStream_t st);
int launch_beam_finalise(asr_handle* h, int k, int max_len, int second_pass, double lm_weight,
                         double length_weight, cudaStream_t st);
// from_logits: argmax / log-sum-exp over materialised logits (export path); else over the tile partials
int launch_greedy_pick(asr_handle* h, int step, bool from_logits, cudaStream_t st);
int launch_greedy_finalise(asr_handle* h, int max_len, cudaStream_t st);
int launch_lm_score(asr_handle* h, const int* d_ids, const int* d_n, int n, int max_n,
                    float* d_scores, cudaStream_t st);

}  // namespace asr
