/*
 * asr_b200.h - C ABI of the B200-native (sm_100a) inference hot path of shawnthu/chinese-asr.
 *
 * The reference has no FFI/plugin interface (it is pure Python); the drop-in boundary is a set
 * of Python signatures.  Every entry point below names the reference
 * call it replaces (file:line into the reference tree).  The Python mirror of the reference
 * interface (chinese_asr_b200/model.py, data.py, main.py) binds exactly these symbols with
 * ctypes; torch tensors are used only as buffer carriers (their data_ptr() is passed here).
 *
 * Conventions
 *   - every function returns 0 on success, a negative asr_status otherwise; nothing throws
 *     across the ABI.  asr_last_error() gives a human-readable message for the last failure
 *     on the calling thread.
 *   - "d_" pointers are device pointers owned by the caller, "h_" pointers are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued on it; functions that return host-side results synchronise that stream.
 *   - one handle per GPU (created on the current device), not shared between threads.
 *   - utterances are identified by their position in the caller's batch ("original order");
 *     the library sorts by length internally (encoder.py:47) and un-sorts every output.
 */
#ifndef ASR_B200_H_
#define ASR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct asr_handle asr_handle;

enum asr_status {
    ASR_OK = 0,
    ASR_ERR_ARG = -1,      /* bad argument (shape, NULL, k not in 1..16, ...)               */
    ASR_ERR_CUDA = -2,     /* a CUDA call failed                                            */
    ASR_ERR_STATE = -3,    /* call order violated (decode before encode, LM not loaded ...) */
    ASR_ERR_CAPACITY = -4  /* batch exceeds what asr_reserve() provisioned                   */
};

/* architecture constants of the path (gpd.py:4-133; frozen at import time in the reference) */
#define ASR_FEAT_DIM 720
#define ASR_N_MELS 80
#define ASR_ENC_OUT 512
#define ASR_ATT 128
#define ASR_DEC_H 512
#define ASR_EMB 256
#define ASR_VOCAB 5004
#define ASR_MAX_BEAM 16

/* Weights in the reference checkpoint layout (Model.load, model.py:357-369; tensor names in
 * SURVEY.md section 8b).  All host pointers, float32, row-major, copied during asr_create.
 * enc_*[layer*2 + dir], dir 0 = forward, 1 = reverse. */
typedef struct asr_weights {
    const float* enc_w_ih[8];   /* [1024, 720] for layer 0, [1024, 512] otherwise */
    const float* enc_w_hh[8];   /* [1024, 256]                                     */
    const float* enc_b_ih[8];   /* [1024]                                          */
    const float* enc_b_hh[8];   /* [1024]                                          */
    const float* embedding;     /* [5004, 256]   decoder.embedding.weight          */
    const float* dec_w_ih;      /* [2048, 768]   cell.cell.0.weight_ih             */
    const float* dec_w_hh;      /* [2048, 512]   cell.cell.0.weight_hh             */
    const float* dec_b_ih;      /* [2048]                                          */
    const float* dec_b_hh;      /* [2048]                                          */
    const float* proj_w;        /* [5004, 1024]  proj_linear.weight                */
    const float* proj_b;        /* [5004]                                          */
    const float* att_w_enc;     /* [512, 128]    attn_mechanism.W_enc  ([in,out])  */
    const float* att_b;         /* [128]                                           */
    const float* att_w_hidden;  /* [512, 128]    attn_mechanism.W_hidden ([in,out])*/
    const float* att_v;         /* [128]                                           */
} asr_weights;

/* Feature constants computed by the host mirror of AudioBase (data.py:371-382). */
typedef struct asr_feature_consts {
    const float* mel_fb;   /* [257, 80] triangular filterbank (data.py:21-57)   */
    const float* window;   /* [400] periodic Hann (data.py:381-382)             */
    const float* taps;     /* [3, 9] identity / delta / delta-delta taps (data.py:138-149) */
    float preemphasis;     /* gpd['preemphasis'] = 0.97                         */
} asr_feature_consts;

const char* asr_last_error(void);
int asr_version(void);

/* Model() + Model.load()  (model.py:18-82, 357-369) and AudioBase() (data.py:371-382). */
int asr_create(asr_handle** out, const asr_weights* w, const asr_feature_consts* fc);
int asr_destroy(asr_handle* h);

/* Provision device workspaces: at most max_utts utterances per batch, max_rows encoder frames
 * summed over the batch, beam width <= max_beam, max_samples PCM samples per batch (0 if the
 * feature kernels are not used), decode length <= max_len. */
int asr_reserve(asr_handle* h, int max_utts, int64_t max_rows, int max_beam, int64_t max_samples,
                int max_len);

/* get_log_mel() + CMVN  (data.py:167-280, main.py:37).
 * d_pcm: concatenated float32 waveforms; h_pcm_off[B+1]: sample offsets of each utterance.
 * d_feats: [sum L_u, 720], utterance-major in original order (what the reference returns per
 * utterance, concatenated).  h_L[B] receives L_u = T_u / 3.  normalise=0 skips the CMVN. */
int asr_features(asr_handle* h, const float* d_pcm, const int64_t* h_pcm_off, int B,
                 float* d_feats, int32_t* h_L, int normalise, void* stream);
/* frame count helper: L for an utterance of n samples */
int asr_num_frames(int64_t n_samples);

/* Batched front end.  The reference reads every WAV with
 * soundfile.read(path, dtype='float32') (fast_read, data.py:109-121), i.e. 16-bit files become
 * x / 32768 on the host; here the 16-bit samples travel to the device as they are (half the
 * host->device bytes) and the log-mel kernel converts them while loading - bit-identical frames. */
enum asr_pcm_format {
    ASR_PCM_F32 = 0,   /* float32 in [-1, 1) */
    ASR_PCM_S16 = 1    /* little-endian int16, value / 32768 */
};
/* asr_features with a sample format and the CMVN epsilon spelled out: 1e-6 is main.py:37 (parse),
 * 1e-7 is AudioLoader.batch_audio (data.py:517).  Offsets count samples, not bytes. */
int asr_features_pcm(asr_handle* h, const void* d_pcm, int format, const int64_t* h_pcm_off, int B,
                     float* d_feats, int32_t* h_L, int normalise, float cmvn_eps, void* stream);
/* AudioLoader.batch_audio (data.py:509-518) for features that already exist on the device:
 * d_out[r, c] = (d_feats[r, c] - mean_u[c]) / (std_u[c] + eps), per utterance u and column c,
 * unbiased std.  d_feats / d_out: [sum h_L, 720] utterance-major; in-place is allowed. */
int asr_cmvn(asr_handle* h, const float* d_feats, const int32_t* h_L, int B, float eps, float* d_out,
             void* stream);

/* RNNEncoder.forward + get_mask_for_softmax + get_initial_state + compute_key_value
 * (encoder.py:36-81, util.py:131-142, decoder.py:56-59, attention.py:67-78).
 * d_feats as produced by asr_features (utterance-major, original order).  The encoder memory,
 * attention keys and the decoder initial state stay inside the handle for the decode calls. */
int asr_encode(asr_handle* h, const float* d_feats, const int32_t* h_L, int B, void* stream);

/* Export the encoder results in the reference's layouts (tests / debugging):
 * d_enc_out [Lmax, B, 512] zero padded, d_keys [Lmax, B, 128] (padded rows = b_attn, as
 * attention.py:77 produces on zero rows), d_h / d_c [B, 512].  Any pointer may be NULL. */
int asr_export_encoder(asr_handle* h, float* d_enc_out, float* d_keys, float* d_h, float* d_c,
                       void* stream);
/* Per-layer residual-stream output of layer `layer` (0..3): [Lmax, B, 512] zero padded. */
int asr_encode_layers(asr_handle* h, const float* d_feats, const int32_t* h_L, int B,
                      int upto_layer, float* d_layer_out, void* stream);

/* Model.eval_one_batch_with_greedy (model.py:503-602) after asr_encode.
 * h_tokens [B, max_len] (argmax token of every executed step), h_len[B] = text_len,
 * h_score_sum[B] = accumulated log-prob, h_finished[B]; *h_steps = number of executed steps.
 * d_align (optional, may be NULL): [max_len, Lmax, B] alignments (EvalOutput.alignment).
 * d_logits (optional): [max_len, B, 5004] raw logits of each step (parity tests). */
int asr_decode_greedy(asr_handle* h, int max_len, int32_t* h_tokens, int32_t* h_len,
                      float* h_score_sum, int32_t* h_finished, int32_t* h_steps, float* d_align,
                      float* d_logits, void* stream);

/* Second-pass LM tables (builder-defined back-off trigram with KenLM scoring semantics; the
 * reference calls kenlm.LanguageModel.score at model.py:755).  Host pointers, copied. */
typedef struct asr_lm_tables {
    const float* uni_logp;    /* [vocab]                       */
    const float* uni_bo;      /* [vocab]                       */
    const int64_t* bi_keys;   /* [bi_cap]  a*V+b or -1         */
    const float* bi_vals;     /* [bi_cap, 2] logp, backoff     */
    int64_t bi_cap;           /* power of two                  */
    const int64_t* tri_keys;  /* [tri_cap] (a*V+b)*V+c or -1   */
    const float* tri_vals;    /* [tri_cap] logp                */
    int64_t tri_cap;          /* power of two                  */
    int32_t vocab;
    int32_t skip_id;          /* token that vanishes under str.split(): id 781 ' ' */
    const int32_t* id_map;    /* [vocab] token id -> LM word id, NULL = identity: tokens the LM has
                                 no unigram for score as <unk> in every n-gram position, like kenlm */
} asr_lm_tables;
int asr_set_lm(asr_handle* h, const asr_lm_tables* t);
/* kenlm-style score of token-id sequences on the device (tests): h_ids [n, max_n], h_n[n] */
int asr_lm_score(asr_handle* h, const int32_t* h_ids, const int32_t* h_n, int n, int max_n,
                 float* h_scores, void* stream);

/* Model.eval_one_batch_with_beam (model.py:604-987) after asr_encode.
 * k = bmsz (1..16).  second_pass != 0 requires asr_set_lm.  Results on the host:
 * h_tokens [B, max_len], h_len[B], h_score[B] (the reference's EvalOutput.score),
 * h_info[4] = {steps executed, stop step or -1, #utterances that took the un-finished
 * fallback, #finished hypotheses}. */
int asr_decode_beam(asr_handle* h, int k, int max_len, float temperature, int second_pass,
                    double lm_weight, double length_weight, int32_t* h_tokens, int32_t* h_len,
                    float* h_score, int32_t* h_info, void* stream);

/* Every device buffer of a handle (weights, workspaces) ends in a 256-byte canary.  Returns the number of buffers
 * whose canary was overwritten (0 = intact; asr_last_error() names the first one), negative on a CUDA error.
 * Synchronises the device.  The GPU tests and smoke() call it: compute-sanitizer is not available on the B200 pool. */
int asr_check_guards(asr_handle* h);

/* h_info[4] of the last decode on this handle (asr_decode_beam, asr_decode_greedy or asr_transcribe*):
 * {steps executed, stop step or -1 (model.py:578, 897-901), #utterances that took the un-finished
 * fallback, #finished hypotheses}. */
int asr_decode_info(asr_handle* h, int32_t* h_info);

/* Per-step internals of the last asr_decode_beam call (parity tests; original utterance order):
 * h_cand_score/h_cand_beam/h_cand_tok [steps, B, 2k] (model.py:863-867),
 * h_backptr/h_active_tok [steps, B, k] (model.py:907-909),
 * h_fin_score [steps, B, k] (score of the EOS candidates among the top k, NaN where none).
 * Any pointer may be NULL. */
int asr_beam_trace(asr_handle* h, float* h_cand_score, int32_t* h_cand_beam, int32_t* h_cand_tok,
                   int32_t* h_backptr, int32_t* h_active_tok, float* h_fin_score);

/* The finished hypotheses of the last asr_decode_beam / asr_transcribe* (k >= 1) call, per utterance in
 * the (step, rank) order in which parse_finished_tensors (model.py:708-747) lists them - what a caller
 * needs to rescore with its OWN lm_model.score (model.py:749-763; main.py:79-85 loads a KenLM binary):
 * h_count[B] = number of finished hypotheses of every utterance (0 = it took the un-finished fallback);
 * the first min(count, cap) of them in h_tokens [B, cap, tok_ld] (padded with <pad>), h_len [B, cap],
 * h_score [B, cap] (accumulated log-prob including the </s> step, no length normalisation).
 * cap = 0 only counts.  Original utterance order. */
int asr_beam_nbest(asr_handle* h, int cap, int tok_ld, int32_t* h_count, int32_t* h_tokens,
                   int32_t* h_len, float* h_score);

/* get_wer (util.py:237-262) as the decode drivers use it when `text` is given (model.py:595-598,
 * 982-985): Levenshtein distance between the predicted and the reference STRING.  Strings are code
 * point sequences: token t stands for h_codepoints[h_tok_off[t] .. h_tok_off[t+1]) (int2word[t];
 * "<unk>" is five characters).  asr_set_vocab copies the table (dict.pkl, data.py:373). */
int asr_set_vocab(asr_handle* h, const int32_t* h_codepoints, const int32_t* h_tok_off, int V);
/* h_dist[B] = edit distance of every utterance; WER of the batch = mean over u of
 * h_dist[u] / (h_ref_off[u+1] - h_ref_off[u]).  h_ref: the reference strings as code points,
 * concatenated; h_ref_off[B+1].  h_hyp [B, hyp_ld] + h_hyp_len[B]: hypotheses as token ids, or
 * h_hyp = NULL to score the hypotheses of the last decode / transcribe call on this handle where
 * they lie on the device (no host round trip of the tokens).  h_hyp_chars (optional): length of
 * every hypothesis string in characters. */
int asr_wer(asr_handle* h, const int32_t* h_hyp, const int32_t* h_hyp_len, int hyp_ld,
            const int32_t* h_ref, const int64_t* h_ref_off, int B, int32_t* h_dist,
            int32_t* h_hyp_chars, void* stream);

/* convert_audio (main.py:19-24: `ffmpeg -sample_fmt s16 -ar 16000 -ac 1` + `sox --norm=-1`): interleaved PCM of any
 * rate and channel count (h_pcm [n_frames, channels], enum asr_pcm_format) -> 16 kHz mono int16 whose peak sits at
 * norm_db dBFS (-1 for the reference's sox call).  Down-mix = channel mean, resampler = Hann-windowed sinc with 16
 * zero crossings and cutoff 0.97 x the lower Nyquist (16 kHz input passes through), normalisation without dither.
 * BUILDER-DEFINED: the two external programs are not part of the reference tree, so there is no reference output
 * to be identical to (parity unpinned; checked against oracle/asr_oracle.py:convert_audio to 1 LSB).  Needs no
 * handle; runs on the current device.  *n_out = asr_convert_audio_length(n_frames, sample_rate) <= out_cap.
 * Decoding compressed containers (the reference accepts whatever ffmpeg reads) is not part of it. */
int64_t asr_convert_audio_length(int64_t n_frames, int sample_rate);
int asr_convert_audio(const void* h_pcm, int format, int64_t n_frames, int channels, int sample_rate,
                      float norm_db, int16_t* h_out, int64_t out_cap, int64_t* n_out, void* stream);

/* Whole path for one batch, host buffers in / host buffers out (parse() in main.py:27-65 for a
 * batch): H2D copy of the PCM, features, encoder, greedy (k = 0) or beam decode, D2H of the
 * hypotheses.  h_pcm should be pinned for the copy to be asynchronous. */
int asr_transcribe(asr_handle* h, const float* h_pcm, const int64_t* h_pcm_off, int B, int k,
                   int max_len, float temperature, int second_pass, double lm_weight,
                   double length_weight, int32_t* h_tokens, int32_t* h_len, float* h_score,
                   void* stream);
/* Pipelining aid for servers: starts the host->device copy of a batch's PCM (pinned host memory) on
 * the handle's own copy stream, so it overlaps the decode of the previous batch.  The next
 * asr_transcribe call given the same h_pcm / offsets consumes the prefetched copy (it waits on the
 * copy's event instead of copying).  At most two prefetches may be outstanding. */
int asr_prefetch_pcm(asr_handle* h, const float* h_pcm, const int64_t* h_pcm_off, int B);

/* asr_transcribe / asr_prefetch_pcm / asr_transcribe_device with a sample format (enum asr_pcm_format)
 * and the CMVN epsilon as arguments; the float32 entry points are these with ASR_PCM_F32, 1e-6. */
int asr_transcribe_pcm(asr_handle* h, const void* h_pcm, int format, float cmvn_eps,
                       const int64_t* h_pcm_off, int B, int k, int max_len, float temperature,
                       int second_pass, double lm_weight, double length_weight, int32_t* h_tokens,
                       int32_t* h_len, float* h_score, void* stream);
int asr_prefetch_pcm_fmt(asr_handle* h, const void* h_pcm, int format, const int64_t* h_pcm_off, int B);
int asr_transcribe_device_pcm(asr_handle* h, const void* d_pcm, int format, float cmvn_eps,
                              const int64_t* h_pcm_off, int B, int k, int max_len, float temperature,
                              int second_pass, double lm_weight, double length_weight,
                              int32_t* h_tokens, int32_t* h_len, float* h_score, void* stream);

/* Same, PCM already resident on the device (d_pcm); used for the HBM-resident throughput. */
int asr_transcribe_device(asr_handle* h, const float* d_pcm, const int64_t* h_pcm_off, int B,
                          int k, int max_len, float temperature, int second_pass, double lm_weight,
                          double length_weight, int32_t* h_tokens, int32_t* h_len, float* h_score,
                          void* stream);

/* Scheduling knob of the encoder recurrence (the nn.LSTM time loop, util.py:1259): a batch is split into
 * `chunks_per_direction` (1..7, default 7) chunks of sequences per direction, each run by one cluster of 8 SMs.
 * 7 = shortest latency for one batch (14 clusters = 112 SMs); fewer, wider chunks leave more SMs to other handles'
 * work on the same GPU at the price of a longer step (a step's MMA phase and gate phase both grow with the rows per
 * cluster; measured at 512 x 10 s: 4.7 ms per batch with 7 chunks).  Results do not depend on it. */
int asr_set_recurrence_chunks(asr_handle* h, int chunks_per_direction);

/* The GEMM engine of the GEMM-shaped stages (nn.LSTM input projections util.py:1259, attention keys
 * attention.py:77, query attention.py:92, nn.LSTMCell util.py:1650-1661, vocabulary projection decoder.py:133):
 * tcgen05 / TMEM / TMA tensor cores in split precision (fp16 hi + bf16 cross terms, fp32-faithful: ~2^-20
 * relative per product; operands beyond +-65504 fall back to bf16 accuracy).  Standalone entry for tests:
 * d_C[M,N] = d_A[M,K] * d_W[N,K]^T + d_bias[N]. */
int asr_test_gemm(asr_handle* h, const float* d_A, const float* d_W, const float* d_bias, float* d_C,
                  int M, int N, int K, void* stream);

/* Tuning aid: average ms per launch of the tensor-core GEMM engine on an [M,K] x [N,K]^T problem
 * (operands already split, CUDA events on `stream`, `iters` timed launches after 2 warm-ups).
 * topk_slots = 0: bias epilogue storing C; 2 / 8 / 16 / 32 (N must be 5004): the vocabulary epilogue
 * (log-sum-exp partials + top candidates per tile, no logits stored). */
int asr_bench_gemm(asr_handle* h, int M, int N, int K, int iters, int topk_slots, float* ms_out,
                   void* stream);

/* Number of kernels this library launched since the handle was created / last reset. */
int64_t asr_launch_count(asr_handle* h, int reset);

/* Stage timing with CUDA events on `stream` for the next transcribe/decode call:
 * h_ms[8] = {features, encoder input GEMMs, encoder recurrence, keys+init, decoder cell,
 * attention, vocab projection, top-k + bookkeeping + finalise}; h_ms[8..11] = {all GEMM-engine
 * launches, operand splits, algorithmic GFLOP of those GEMMs, the attention kernel alone} (nested in
 * the stages).  enable bit 0 records events around every stage (adds event overhead; the decode loop then runs
 * eagerly instead of from its CUDA graph); asr_stage_times reads them after a sync.  enable bit 1 prints the
 * recurrence kernel's per-step clock64 timeline (phases of cluster 0) to stderr. */
int asr_stage_timing(asr_handle* h, int enable);
int asr_stage_times(asr_handle* h, float* h_ms, int n);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H_ */
