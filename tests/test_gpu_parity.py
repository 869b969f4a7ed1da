"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
(libasr_b200.so via chinese_asr_b200.model.Model); the oracle / committed reference goldens are
only the checker.  Tolerances (BASELINE.json north_star): tokens, back-pointers and n-best order
bit-exact; features and logits <= 1e-3 absolute (fp32); final scores <= 1e-3 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FEAT_TOL = 1e-3       # absolute, fp32 features (north_star)
LOGIT_TOL = 1e-3      # absolute, fp32 logits
SCORE_RTOL = 1e-3     # relative, final beam scores
ENC_TOL = 1e-4        # absolute, encoder memory / keys / states (tighter than required)


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_checks
    return gpu_checks


def test_features(G):
    r = G.check_features()
    for k, v in r.items():
        if k.endswith("_L_ok"):
            assert v == 1, k
        else:
            assert v <= FEAT_TOL, (k, v)
    assert r["zero_raw_vs_ref"] == 0.0          # exact-zero -> eps branch (data.py:223)


def test_encoder_known_answer(G):
    r = G.check_encoder_kat()                   # encoder.py:636-652
    assert abs(r["out_sum"] - r["want_out"]) < 0.5
    assert abs(r["h_sum"] - r["want_h"]) < 1e-2 and abs(r["c_sum"] - r["want_c"]) < 1e-2


@pytest.mark.parametrize("cname", ["greedy3", "beam4", "beam16"])
def test_encoder(G, cname):
    r = G.check_encoder(cname)
    assert r["pad_exact_zero"] == 1
    for k, v in r.items():
        if k != "pad_exact_zero":
            assert v <= ENC_TOL, (k, v)


def test_gemm_engine_fp32_faithful(G):
    # max |C - C_fp64| / max |C_fp64| of the tcgen05 split-precision engine (fp16 hi + bf16 cross terms).  The
    # products are fp32-faithful; the tensor core accumulates with truncation, which the cross-first / hi-second
    # pass order keeps to K / 16 full-magnitude steps (a CUDA-core fp32 FMA loop measures ~1e-6 here).  Also:
    # row-wise accuracy over fp16's normal range (1e-4 ... 1e4 row magnitudes), the absolute error bound below it,
    # and finite results (bf16-accurate residual) when |a| > 65504.
    r = G.check_gemm()
    print("[gemm accuracy]", r)
    for k, v in r.items():
        assert v <= 5e-6, (k, v)


def test_lm_scores_bit_exact(G):
    r = G.check_lm()
    assert r["mismatch"] == 0 and r["string_api"] == 0.0


@pytest.mark.parametrize("cname", ["greedy1", "greedy3"])
def test_greedy(G, cname):
    r = G.check_greedy(cname)
    assert r["text_vs_ref"] == 1 and r["text_vs_oracle"] == 1 and r["len_vs_ref"] == 1
    assert r["steps"] == r["oracle_steps"]
    assert r["score_rel_vs_ref"] <= SCORE_RTOL
    assert r["logit_all_vs_oracle"] <= LOGIT_TOL
    assert r["align0_vs_ref"] <= 1e-5


def test_greedy_plain_init(G):
    """Plain init: logit gaps ~1e-6, so only the first step's logits, lengths and scores are pinned."""
    r = G.check_greedy("greedy3_plain")
    assert r["logit0_vs_oracle"] <= LOGIT_TOL
    assert r["len_vs_ref"] == 1 and r["score_rel_vs_ref"] <= SCORE_RTOL


@pytest.mark.parametrize("cname", ["beam4", "beam4es", "beam16", "beam8lm"])
def test_beam(G, cname):
    r = G.check_beam(cname)
    assert r["text_vs_ref"] == 1 and r["tokens_vs_oracle"] == 1, r
    assert r["steps"] == r["oracle_steps"] and r["stopped_at"] == r["oracle_stopped_at"]
    assert r["fallback"] == r["oracle_fallback"] and r["finished"] == r["oracle_finished"]
    assert r["score_rel_vs_ref"] <= SCORE_RTOL
    assert r.get("backptr_mismatch", 0) == 0 and r.get("active_tok_mismatch", 0) == 0
    assert r.get("backptr_mismatch_vs_ref", 0) == 0
    assert r["cand_scores_rel_vs_ref"] <= 1e-4        # relative: accumulated scores reach |s| ~ 300
    # the reference's recorded torch.topk indices: every CONSUMED candidate (top k + everything up to the k-th
    # non-</s> one) is identical; an unconsumed tail rank may differ only as a proven near-tie
    assert r["cand_index_consumed_mismatch_vs_ref"] == 0, r
    assert all(mg < 1e-4 for mg in r["cand_index_tail_mismatch_margins"]), r      # unconsumed tail ranks: gap printed


def test_beam_plain_init_fallback(G):
    r = G.check_beam("beam4_plain")
    assert r["fallback"] == r["oracle_fallback"] == 2 and r["steps"] == 40
    assert r["score_rel_vs_ref"] <= SCORE_RTOL


def test_fused_transcribe_equals_staged(G):
    r = G.check_fused()
    assert all(v == 1 for v in r.values()), r


def test_batch_invariance(G):
    r = G.check_batch_invariance()
    assert r["single_eq_batch"] == r["of"] and r["perm_invariant"] == 1


@pytest.mark.parametrize("which", ["keys", "queries"])
def test_attention_out_of_range_falls_back_to_exact_scores(G, which):
    r = G.check_attention_range(which)
    if which == "keys":
        assert r["max_abs_key"] > 21.0, r          # the case really is outside the product form's range
    assert r["beam_tokens_vs_oracle"] == 1 and r["greedy_text_vs_oracle"] == 1, r
    assert r["stream_greedy_text_vs_oracle"] == 1 and r["beam_score_rel"] <= SCORE_RTOL, r


def test_graph_replay_equals_eager_loop(G):
    r = G.check_graph_replay()
    assert r["replay_eq_eager"] == r["of"] and r["batches_differ"] == 1, r


def test_full_size_properties(G):
    """BASELINE.json configs[1] shape (bw=4, 32 x 10 s) through the fused path: every utterance
    decodes, lengths are within [0, max_len], scores are finite and the result is reproducible
    (bit-identical when run twice)."""
    from oracle import asr_oracle as O
    m = G.get_model((1234, "sharp", 8.0), O.make_weights(1234, "sharp", eos_bias=8.0))
    B, n = 32, 160000
    pcm = np.stack([O.synth_pcm(7000 + i, n) for i in range(B)]).reshape(-1)
    off = np.arange(B + 1, dtype=np.int64) * n
    t1, l1, s1 = m.transcribe(pcm, off, bw=4)
    t2, l2, s2 = m.transcribe(pcm, off, bw=4)
    assert np.array_equal(t1, t2) and np.array_equal(l1, l2) and np.array_equal(s1, s2)
    assert l1.min() >= 0 and l1.max() <= 40 and np.isfinite(s1).all()
    # greedy == beam with k = 1 on the argmax path (same first token wherever both emit one)
    tg, lg, sg = m.transcribe(pcm, off, bw=None)
    tb, lb, sb = m.transcribe(pcm, off, bw=1)
    both = (lg > 0) & (lb > 0)
    assert (tg[both, 0] == tb[both, 0]).all()


def test_errors_are_loud(G):
    from chinese_asr_b200._cabi import AsrError
    from oracle import asr_oracle as O
    m = G.get_model((1234, "plain", None), O.make_weights(1234, "plain"))
    with pytest.raises(AsrError):
        m.features([np.zeros(100, dtype=np.float32)])          # too short for one STFT frame
    with pytest.raises((AsrError, ValueError)):
        m.eval_one_batch_with_beam(m.device, 17, [torch.zeros(5, 720)], torch.tensor([5]), None, {}, second_pass=False)


def test_config1_greedy_5s(G):
    """BASELINE.json configs[0]: greedy decode of one 5 s utterance."""
    r = G.check_config_shape(1, None, [5.0])
    assert r["exact"] == 1 and r["lens_max"] == 165


def _strict(r, n):
    """Every compared utterance identical to the oracle at every step; a difference only as a proven near-tie
    (gpu_checks.near_tie: the oracle's gap at the first differing rank <= twice the measured score noise of that
    step, which itself must be <= 1e-4 relative; or gap < 2e-6), listed in r['flips'] and printed."""
    assert r["bad"] == [], r
    assert r["exact"] + len(r["flips"]) == n and r["picks"] == n, r
    assert len(r["flips"]) <= max(2, n // 4), r              # proven near-ties only (each printed with its gap / noise)
    assert r["score_rel_max"] <= SCORE_RTOL and r["cand_score_rel_max"] <= 1e-4, r


def test_config2_beam4_batch32_10s(G):
    """configs[1]: bw=4, 32 x 10 s, every utterance against the oracle step by step."""
    r = G.check_config_shape(32, 4, [10.0] * 32)
    _strict(r, 32)
    assert r["finished_hyps"] > 32 and r["min_margin"] > 0.0


def test_config3_beam16_mixed_2_to_20s(G):
    """configs[2] in shape (mixed 2-20 s, padding / masking / EOS handling), 12 utterances."""
    secs = [2.0, 20.0, 3.7, 11.3, 7.9, 16.4, 2.6, 13.0, 5.5, 19.2, 9.1, 4.4]
    r = G.check_config_shape(12, 16, secs, seed0=3300, eos_bias=9.0, wseed=77)
    _strict(r, 12)


def test_config4_beam8_lm_batch32_10s(G):
    """configs[3]: bw=8 + second-pass LM rescoring, 32 x 10 s."""
    r = G.check_config_shape(32, 8, [10.0] * 32, lm_seed=7, seed0=3600)
    _strict(r, 32)


def test_bench_shape_picks_vs_oracle(G):
    """The bench workload's shape (configs[4] per GPU: 512 x 10 s, bw=8; 80 sequences per recurrence cluster,
    attention_stream_kernel<8>, 224-wide vocabulary tiles): 32 picked utterances, including the first and last
    of every recurrence chunk, against the oracle."""
    r = G.check_bench_shape()
    _strict(r, 32)
    assert r["B"] == 512 and r["finished_hyps"] > 32


def test_odd_beam_widths_vs_oracle(G):
    """bw 1...16 (reference main.py knob): widths between the instantiation sizes (3, 6, 11) against the oracle, on
    the two-phase attention path (small batches) and the streaming path (300 utterances)."""
    r = G.check_odd_beam_widths()
    for name, res in r.items():
        _strict(res, res["picks"])
        assert res["finished_hyps"] > 0, name


def test_config3_full_size_picks_vs_oracle(G):
    """configs[2] at full size: bw=16, 256 mixed 2-20 s utterances; 24 picks incl. shortest / longest."""
    r = G.check_config3_full()
    _strict(r, 24)
    assert r["reproducible"] == 1 and r["secs_min"] == 2 and r["secs_max"] == 20


# ---- rows either side of the hot path -----------------------------
def test_frontend_int16_ingest_and_batch_audio(G):
    r = G.check_frontend()
    assert r["s16_eq_f32_bitwise"] == 1 and r["lens_ok"] == 1 and r["collate_ok"] == 1 and r["loader_ok"] == 1, r
    for k, v in r.items():
        if "_vs_" in k:
            assert v <= FEAT_TOL, (k, v)


def test_device_wer_bit_exact(G):
    r = G.check_wer()
    assert r["dist_mismatch"] == 0 and r["per_utt_max_abs"] <= 1e-12 and r["string_refs_same"] == 1, r


@pytest.mark.parametrize("cname", ["greedy3", "beam4"])
def test_driver_wer_matches_reference(G, cname):
    r = G.check_driver_wer(cname)
    assert abs(r["wer"] - r["ref_wer"]) < 1e-9 and r["text_ok"] == 1, r
    assert abs(r["test_model_wer"] - r["ref_wer"]) < 1e-9 and r["test_model_n"] == 2 * len(G.CASES[cname]["seeds"])
    assert r["test_model_error_rate"] == r["oracle_error_rate"]


def test_transcribe_int16_equals_float32(G):
    """asr_transcribe_pcm with 16-bit samples (half the host->device bytes) decodes exactly what the
    float32 path decodes from the same samples / 32768, resident or through the prefetch pipeline."""
    from oracle import asr_oracle as O
    m = G.get_model((1234, "sharp", 8.0), O.make_weights(1234, "sharp", eos_bias=8.0))
    B, n = 6, 48000
    x16 = np.stack([O.synth_pcm_int16(8100 + i, n) for i in range(B)]).reshape(-1)
    x32 = O.pcm_from_int16(x16)
    off = np.arange(B + 1, dtype=np.int64) * n
    a = m.transcribe(x32, off, bw=4)
    b = m.transcribe(x16, off, bw=4)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    h16 = torch.from_numpy(x16).pin_memory()
    m.prefetch(h16, off, bw=4)
    c = m.transcribe(h16, off, bw=4)
    d = m.transcribe(torch.from_numpy(x16).cuda(), off, bw=4, resident=True)
    assert all(np.array_equal(p, q) for p, q in zip(a, c)) and all(np.array_equal(p, q) for p, q in zip(a, d))
    with pytest.raises(TypeError):
        m.transcribe(x16.astype(np.int32), off, bw=4)


@pytest.mark.parametrize("B", [600, 700])
def test_wide_recurrence_batches(G, B):
    """96 (B = 600) and 128 (B = 700) sequences per recurrence cluster: 16 picks (first / last of every chunk)
    against the oracle; and the same hypotheses as in batches of 40 (16 per cluster) wherever the stop step agrees -
    an utterance that differs must be a proven near-tie against the oracle in both batchings."""
    r = G.check_wide_recurrence(B)
    _strict(r, 16)
    assert r["of"] >= B // 2 and r["same"] + len(r["differ"]) == r["of"] and r["len_spread"] > 0, r
    assert r["differ_bad"] == [] and len(r["differ"]) <= B // 100 and r["self_score_rel"] <= 1e-4, r


def test_frontend_and_wer_errors_are_loud(G):
    import ctypes as C
    from chinese_asr_b200 import _cabi
    from chinese_asr_b200._cabi import AsrError
    from chinese_asr_b200.model import Model
    from oracle import asr_oracle as O
    m = Model()
    m.load_state(O.make_weights(1234, "plain"))
    i2w = G.vocab()[1]
    with pytest.raises(AsrError):                                   # no resident hypotheses yet, no vocabulary
        _cabi.check(_cabi.lib.asr_wer(m._h, None, None, 0, (C.c_int32 * 1)(65), (C.c_int64 * 2)(0, 1), 1,
                                      (C.c_int32 * 1)(), None, None), "asr_wer")
    with pytest.raises(AsrError):                                   # decode first: nothing resident for B = 3
        m.wer([[10], [11], [12]], i2w)
    with pytest.raises(AsrError):                                   # unknown sample format
        m.reserve(1, 100, 4, 32000)
        x = torch.zeros(32000, device="cuda")
        _cabi.check(_cabi.lib.asr_features_pcm(m._h, C.c_void_p(x.data_ptr()), 7, (C.c_int64 * 2)(0, 32000), 1,
                                               C.c_void_p(x.data_ptr()), (C.c_int32 * 1)(), 1, 1e-6, None), "asr_features_pcm")
    with pytest.raises(ValueError):
        m.cmvn([torch.zeros(5, 100)])
    one = m.cmvn([torch.ones(1, 720, device="cuda")])[0]           # a single row: unbiased std is NaN, like torch.std
    assert bool(torch.isnan(one).all())
    m.close()


def test_batch_pipeline_two_engines_equal_one(G):
    """Two handles on one GPU with batches in flight on two host threads / streams (BatchPipeline) decode
    exactly what one handle decodes, batch by batch, in submission order."""
    from chinese_asr_b200.model import Model
    from chinese_asr_b200.parallel import BatchPipeline
    from oracle import asr_oracle as O
    w = O.make_weights(1234, "sharp", eos_bias=8.0)
    one = G.get_model((1234, "sharp", 8.0), w)
    B, n = 24, 40000
    batches = []
    for j in range(5):
        x = np.stack([O.synth_pcm_int16(9100 + 100 * j + i, n) for i in range(B)]).reshape(-1)
        batches.append((x, np.arange(B + 1, dtype=np.int64) * n))
    want = [one.transcribe(x, off, bw=4) for x, off in batches]
    engines = []
    for _ in range(2):
        m = Model()
        m.load_state(w)
        engines.append(m)
    pipe = BatchPipeline(engines)
    for rounds in range(2):                     # second round: both engines replay their captured graphs
        got = pipe.map(batches, bw=4)
        for a, b in zip(want, got):
            assert all(np.array_equal(p, q) for p, q in zip(a, b))
    # every engine stages its next batch on its copy stream while it decodes the current one
    pinned = [(torch.from_numpy(x).pin_memory(), off) for x, off in batches]
    got = pipe.map(pinned, prefetch=True, bw=4)
    for a, b in zip(want, got):
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
    pipe.close()
    for m in engines:
        m.close()


# ---- boundary: lm_model duck typing and ARPA vocabularies ---------------------------------------------
def test_host_scorable_lm_model_equals_device_tables(G):
    """model.py:749-763 / main.py:79-85: lm_model is any object with .score(); here it is rescored on the host
    from asr_beam_nbest and picks what the device tables (and the reference) pick."""
    r = G.check_host_lm()
    assert r["host_eq_device"] == 1 and r["host_eq_ref"] == 1 and r["fused_eq_device"] == 1, r
    assert r["lm_changes_a_pick"] == 1 and r["lm_calls"] == r["ref_lm_calls"] > 0, r


def test_arpa_vocabulary_mismatch_scores_like_kenlm(G):
    r = G.check_arpa_oov()
    assert r["max_abs"] <= 2e-5 and r["unk_kept"] == 1 and r["missing_maps_to_unk"] == 1, r


def test_reload_resets_lm_and_vocab_and_mixed_pcm_is_scaled(G):
    """a second load_state() must not keep the old handle's LM / vocabulary caches; a mixed
    int16 / float32 batch is scaled like fast_read (x / 32768)."""
    from chinese_asr_b200.lm import NGramLM
    from chinese_asr_b200.model import Model
    from oracle import asr_oracle as O
    w2i, i2w = G.vocab()
    w = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = Model()
    m.load_state(w)
    lm = NGramLM(O.NGramLM(seed=7, word2int=w2i).tables(), w2i)
    pcm = [O.synth_pcm_int16(77, 30000), O.synth_pcm_int16(78, 26000)]
    off = np.array([0, 30000, 56000], dtype=np.int64)
    a = m.transcribe(np.concatenate(pcm), off, bw=4, second_pass=True, lm_model=lm, lm_weight=0.3, length_weight=2.0)
    m.wer([[10], [11]], i2w)
    m.load_state(w)                                            # new handle: tables and vocabulary must be re-uploaded
    b = m.transcribe(np.concatenate(pcm), off, bw=4, second_pass=True, lm_model=lm, lm_weight=0.3, length_weight=2.0)
    m.wer([[10], [11]], i2w)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    mixed = m.features([pcm[0], O.pcm_from_int16(pcm[1])], normalise=False)
    both16 = m.features(pcm, normalise=False)
    assert all(torch.equal(p, q) for p, q in zip(mixed, both16))
    with pytest.raises(TypeError):
        m.features([pcm[0].astype(np.int32)])
    m.close()


def test_zz_device_buffer_guards_intact(G):
    """Last test of the file: every engine the suite used (all batch shapes, beam widths, LM, WER, front end) still has
    the canary behind each of its device buffers (asr_check_guards) - the library's own out-of-bounds check, since
    compute-sanitizer is closed on the B200 pool."""
    assert len(G._models) >= 4
    for key, m in G._models.items():
        m.check_guards()


def test_convert_audio_vs_oracle(G, tmp_path):
    """convert_audio (main.py:19-24) on the device against the numpy restatement of the same builder-defined filter:
    16-bit output within 1 LSB (fp32 taps on the device, float64 in the oracle), lengths equal, 16 kHz pass-through
    and the WAV round trip through data.convert_audio."""
    r = G.check_convert_audio(str(tmp_path))
    for name, (n_dev, n_ora, max_diff, frac_diff) in r["cases"].items():
        assert n_dev == n_ora, name
        assert max_diff <= 1, (name, max_diff)
        assert frac_diff <= 0.02, (name, frac_diff)      # share of samples that differ at all
    assert r["wav_rate"] == 16000 and r["wav_channels"] == 1 and r["wav_equal"]
