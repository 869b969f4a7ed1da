"""Parity checks of the CUDA path (through the C ABI / Python mirror) against the oracle and the
committed reference goldens.  Each check returns a dict of metrics; tests/test_gpu_parity.py asserts
on them and tools/gpu_diag.py prints them all (first-run diagnostics)."""
import os
import pickle

import numpy as np
import torch

from oracle import asr_oracle as O
from tests.cases import CASES, case_inputs, case_weights, load_golden

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_vocab = None
_models = {}


def vocab():
    global _vocab
    if _vocab is None:
        with open(os.path.join(GOLD_DIR, "dict.pkl"), "rb") as f:
            _vocab = pickle.load(f)
    return _vocab


def get_model(weights_key, weights):
    """One engine per distinct weight set (handle creation uploads ~64 MB)."""
    from chinese_asr_b200.model import Model
    from chinese_asr_b200.gpd import gpd
    gpd['verbose'] = False
    if weights_key not in _models:
        m = Model()
        m.load_state(weights)
        _models[weights_key] = m
    return _models[weights_key]


def wkey(cs):
    return (cs["wseed"], cs["variant"], cs.get("eos_bias"))


def maxabs(a, b):
    return float((torch.as_tensor(a).float().cpu() - torch.as_tensor(b).float().cpu()).abs().max())


# ---------------------------------------------------------------------------------------------
def check_features():
    g = load_golden()
    w = O.make_weights(1234, "plain")
    m = get_model((1234, "plain", None), w)
    res = {}
    pcms = {"feat_2s": O.synth_pcm(11, 32000), "feat_5s": O.synth_pcm(12, 80000),
            "feat_odd": O.synth_pcm(13, 20011), "feat_10s": O.synth_pcm(14, 160000),
            "feat_20s": O.synth_pcm(15, 320000)}
    names = list(pcms)
    raw = m.features([pcms[n] for n in names], normalise=False)
    nrm = m.features([pcms[n] for n in names], normalise=True)
    for n, fr, fn in zip(names, raw, nrm):
        rows = g[n + "_rows"]
        res[n + "_L_ok"] = int(fr.size(0) == int(g[n + "_L"]))
        res[n + "_raw_vs_ref"] = maxabs(fr[rows], g[n + "_raw"])
        res[n + "_cmvn_vs_ref"] = maxabs(fn[rows], g[n + "_cmvn"])
        res[n + "_raw_vs_oracle"] = maxabs(fr, O.features(pcms[n], normalise=False))
        res[n + "_cmvn_vs_oracle"] = maxabs(fn, O.features(pcms[n], normalise=True))
    z = m.features([np.zeros(8000, dtype=np.float32)], normalise=False)[0]
    res["zero_raw_vs_ref"] = maxabs(z, g["feat_zero_raw"])
    return res


def check_encoder(cname):
    g = load_golden()
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    pcms, feats, lens = case_inputs(cs)
    res = {}
    o_out, _, (o_h, o_c), o_layers = O.encoder_forward(weights, feats, lens, return_layers=True)
    for layer in range(4):
        y = m.encode_layer(feats, lens, layer)
        res[f"layer{layer}_vs_oracle"] = maxabs(y, o_layers[layer])
    enc, keys, h, c = m.encode_export(feats, lens)
    res["enc_vs_oracle"] = maxabs(enc, o_out)
    res["enc_vs_ref"] = maxabs(enc, g[cname + "_enc_out"])
    res["h_vs_ref"] = maxabs(h, g[cname + "_enc_h"])
    res["c_vs_ref"] = maxabs(c, g[cname + "_enc_c"])
    res["keys_vs_ref"] = maxabs(keys, g[cname + "_keys"])
    # padded rows must be exactly zero (pad_packed_sequence, encoder.py:63-64)
    pad_ok = 1
    for b, n in enumerate(lens.tolist()):
        if n < enc.size(0) and float(enc[n:, b].abs().max()) != 0.0:
            pad_ok = 0
    res["pad_exact_zero"] = pad_ok
    return res


def check_encoder_kat():
    """The reference's only reproducible known answer (encoder.py:636-652): all-ones weights and
    inputs, lens [10, 8, 23, 14] -> out.sum()=110345.43, h.sum()=2048, c.sum()=28160."""
    g = load_golden()
    w = O.make_weights(1, "plain")
    for d in (w["encoder_state_dict"],):
        for k in d:
            d[k] = torch.ones_like(d[k])
    m = get_model(("ones",), w)
    lens = torch.tensor([10, 8, 23, 14])
    feats = [torch.ones(n, 720) for n in lens.tolist()]
    enc, _, h, c = m.encode_export(feats, lens)
    want = g["kat_rnn_sums"]
    return {"out_sum": float(enc.double().sum()), "h_sum": float(h.double().sum()),
            "c_sum": float(c.double().sum()), "want_out": float(want[0]), "want_h": float(want[1]),
            "want_c": float(want[2])}


def check_greedy(cname):
    g = load_golden()
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    _, i2w = vocab()
    pcms, feats, lens = case_inputs(cs)
    tr = {}
    o = O.greedy_decode(weights, feats, lens, i2w, trace=tr)
    out, logits, toks = m.eval_one_batch_with_greedy(m.device, feats, lens, i2w, None, return_logits=True)
    res = {"steps": int(logits.size(0)), "oracle_steps": o["steps"]}
    res["text_vs_ref"] = int(list(out.pred_text) == list(g[cname + "_text"]))
    res["text_vs_oracle"] = int(list(out.pred_text) == o["pred_text"])
    res["len_vs_ref"] = int(out.text_len.tolist() == g[cname + "_text_len"].tolist())
    ref_s = g[cname + "_score"]
    res["score_rel_vs_ref"] = float(max(abs(a - b) / max(1e-6, abs(b)) for a, b in zip(out.score, ref_s)))
    n = min(logits.size(0), len(tr["logit"]))
    res["logit0_vs_oracle"] = maxabs(logits[0], tr["logit"][0])
    # later steps only comparable while the fed-back tokens agree
    res["logit_all_vs_oracle"] = max(maxabs(logits[s], tr["logit"][s]) for s in range(n))
    res["align0_vs_ref"] = maxabs(out.alignment[0], g[cname + "_align0"])
    top2 = torch.stack(tr["logit"]).topk(2, dim=2)[0]
    res["oracle_min_margin"] = float((top2[..., 0] - top2[..., 1]).min())
    return res


def check_beam(cname):
    from chinese_asr_b200.gpd import gpd
    from chinese_asr_b200.lm import NGramLM
    g = load_golden()
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    w2i, i2w = vocab()
    pcms, feats, lens = case_inputs(cs)
    k = cs["bw"]
    lm_o = O.NGramLM(seed=cs["lm"], word2int=w2i) if cs.get("lm") else None
    lm_d = NGramLM(lm_o.tables(), w2i) if lm_o is not None else None
    tr = {}
    o = O.beam_decode(weights, k, feats, lens, i2w, second_pass=lm_o is not None, lm_model=lm_o,
                      lm_weight=cs.get("lm_weight", 0.0), length_weight=cs.get("length_weight", 0.0),
                      temperature=cs.get("temperature", 1), trace=tr)
    gpd['temperature'] = cs.get("temperature", 1)
    try:
        out = m.eval_one_batch_with_beam(m.device, k, feats, lens, None, i2w,
                                         second_pass=lm_d is not None, lm_model=lm_d,
                                         lm_weight=cs.get("lm_weight", 0.0),
                                         length_weight=cs.get("length_weight", 0.0))
    finally:
        gpd['temperature'] = 1.
    info = m.last_beam_info
    t = m.beam_trace(len(feats), k)
    res = {"steps": info["steps"], "oracle_steps": o["steps"], "stopped_at": info["stopped_at"],
           "oracle_stopped_at": -1 if o["stopped_at"] is None else o["stopped_at"],
           "fallback": info["fallback"], "oracle_fallback": len(o["fallback"]),
           "finished": info["finished"], "oracle_finished": sum(len(v) for v in o["nbest"].values()),
           "oracle_min_margin": min(tr["min_margin"])}
    res["text_vs_ref"] = int(list(out.pred_text) == list(g[cname + "_text"]))
    res["tokens_vs_oracle"] = int(info["tokens"] == o["tokens"])
    ref_s = g[cname + "_score"]
    res["score_rel_vs_ref"] = float(max(abs(a - b) / max(1e-6, abs(b)) for a, b in zip(out.score, ref_s)))
    # per-step internals against the reference's recorded torch.topk outputs
    ns = min(info["steps"], g[cname + "_cand_scores"].shape[0])
    ref_cs = g[cname + "_cand_scores"][:ns]
    res["cand_scores_vs_ref"] = float(np.abs(t["cand_scores"][:ns] - ref_cs).max())
    res["cand_scores_rel_vs_ref"] = float((np.abs(t["cand_scores"][:ns] - ref_cs) / np.maximum(1.0, np.abs(ref_cs))).max())
    flat = t["cand_beams"][:ns].astype(np.int64) * O.VOCAB + t["cand_tokens"][:ns]
    ref_idx = g[cname + "_cand_index"][:ns]
    res["cand_index_mismatch_vs_ref"] = int((flat != ref_idx).sum())
    # consumed candidates: ranks < k (finished set) and, while an active set is formed, everything up to the
    # k-th non-</s> candidate; mismatches beyond are reported with the oracle's gap at that rank
    consumed_bad, tail_margins = 0, []
    n_act = g[cname + "_active"].shape[0]
    for s_, u, r_ in np.argwhere(flat != ref_idx):
        ref_tok = ref_idx[s_, u] % O.VOCAB
        used = k if s_ >= n_act else max(k, int(np.searchsorted(np.cumsum(ref_tok != O.EOS), k)) + 1)
        if r_ < used:
            consumed_bad += 1
        else:
            sc_ref = g[cname + "_cand_scores"][s_, u]
            lo, hi = max(0, r_ - 1), min(2 * k - 1, r_ + 1)
            tail_margins.append(float(min(abs(sc_ref[lo] - sc_ref[r_]) if lo != r_ else np.inf,
                                          abs(sc_ref[hi] - sc_ref[r_]) if hi != r_ else np.inf)))
    res["cand_index_consumed_mismatch_vs_ref"] = consumed_bad
    res["cand_index_tail_mismatch_margins"] = tail_margins
    nb = len(tr.get("backptr", []))
    if nb:
        bp = torch.stack(tr["backptr"]).numpy()
        at = torch.stack(tr["active_tokens"]).numpy()
        res["backptr_mismatch"] = int((t["backptr"][:nb] != bp).sum())
        res["active_tok_mismatch"] = int((t["active_tokens"][:nb] != at).sum())
        act_ref = g[cname + "_active"]
        if act_ref.shape[0]:
            # reference active_hypos index into the 2k candidates -> beams via its cand_index
            na = min(nb, act_ref.shape[0])
            ref_beam = np.take_along_axis(g[cname + "_cand_index"][:na] // O.VOCAB, act_ref[:na], axis=2)
            res["backptr_mismatch_vs_ref"] = int((t["backptr"][:na] != ref_beam).sum())
    return res


def check_gemm():
    """Both GEMM engines against a float64 product: fp32-faithful means ~1e-6 relative error."""
    m = get_model((1234, "plain", None), O.make_weights(1234, "plain"))
    g = torch.Generator().manual_seed(3)
    res = {}
    for (M, N, K) in ((300, 2048, 720), (256, 5004, 1024), (129, 128, 512), (2500, 2048, 512)):
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) * 0.05
        b = torch.randn(N, generator=g)
        ref = (A.double() @ W.double().t() + b.double())
        out = m.test_gemm(A.cuda(), W.cuda(), b.cuda()).cpu().double()
        scale = float(ref.abs().max())
        res[f"{M}x{N}x{K}_relerr"] = float((out - ref).abs().max()) / scale
    # dynamic range of the fp16 hi part, per ROW of A so that every output row is dominated by its own magnitude.
    # fp16's normal range (6.1e-5 ... 65504): full relative accuracy, measured per row against the row's largest |C|
    M, N, K = 384, 256, 512
    W = torch.randn(N, K, generator=g) * 0.05
    b = torch.zeros(N)
    mag = 10.0 ** torch.linspace(-4.0, 4.0, M)            # 6 sigma of the largest row stays below 65504
    A = torch.randn(M, K, generator=g) * mag[:, None]
    ref = A.double() @ W.double().t()
    out = m.test_gemm(A.cuda(), W.cuda(), b.cuda()).cpu().double()
    res["normal_range_rowwise_relerr"] = float(((out - ref).abs().max(dim=1).values / ref.abs().max(dim=1).values).max())
    # below it (down to values whose hi part is 0): the cross operand carries what hi dropped, so the ABSOLUTE error
    # stays ~2^-34 |w| per term (outputs here are ~1e-6; the bound asserted is 1e-8 for the tensor-core engine)
    mag = 10.0 ** torch.linspace(-8.0, -5.0, M)
    A = torch.randn(M, K, generator=g) * mag[:, None]
    ref = A.double() @ W.double().t()
    out = m.test_gemm(A.cuda(), W.cuda(), b.cuda()).cpu().double()
    res["tiny_abs_err_x1e3"] = float((out - ref).abs().max()) * 1e3
    # beyond fp16's range the hi part saturates and the residual is carried at bf16 accuracy only: still finite
    A2 = torch.randn(128, K, generator=g) * 3e5
    out2 = m.test_gemm(A2.cuda(), W.cuda(), b.cuda()).cpu().double()
    ref2 = A2.double() @ W.double().t()
    res["saturated_finite"] = 0.0 if bool(torch.isfinite(out2).all()) else 1.0
    res["saturated_relerr_x1e-3"] = float((out2 - ref2).abs().max() / ref2.abs().max()) * 1e-3
    return res


def check_lm():
    from chinese_asr_b200.lm import NGramLM
    w2i, i2w = vocab()
    lm_o = O.NGramLM(seed=7, word2int=w2i)
    m = get_model((1234, "plain", None), O.make_weights(1234, "plain"))
    lm_d = NGramLM(lm_o.tables(), w2i)
    m.set_lm(lm_d)
    rng = np.random.default_rng(5)
    seqs = [[]]
    for _ in range(300):
        n = int(rng.integers(0, 40))
        hot = rng.random() < 0.6
        s = rng.integers(0, 48 if hot else O.VOCAB, n).tolist()
        if n and rng.random() < 0.3:
            s[int(rng.integers(0, n))] = O.SPACE_ID
        seqs.append(s)
    dev = m.lm_score(seqs)
    ref = np.array([lm_o.score_ids(s) for s in seqs], dtype=np.float32)
    sent = " ".join(i2w[t] for t in seqs[5])
    return {"n": len(seqs), "mismatch": int((dev != ref).sum()), "maxabs": float(np.abs(dev - ref).max()),
            "string_api": float(abs(lm_d.score(sent) - lm_o.score(sent)))}


def check_fused(cname="beam4"):
    """asr_transcribe (PCM in, hypotheses out, one call) == staged features/encode/decode."""
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    _, i2w = vocab()
    pcms, feats, lens = case_inputs(cs)
    dfeats = m.features(pcms, normalise=True)
    staged = m.eval_one_batch_with_beam(m.device, cs["bw"], dfeats, lens, None, i2w, second_pass=False)
    off = np.zeros(len(pcms) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in pcms])
    pin = torch.from_numpy(np.concatenate(pcms)).pin_memory()
    tok, ln, sc, texts = m.transcribe(pin, off, bw=cs["bw"], int2word=i2w)
    dev = torch.from_numpy(np.concatenate(pcms)).cuda()
    tok2, ln2, sc2, texts2 = m.transcribe(dev, off, bw=cs["bw"], int2word=i2w, resident=True)
    tokg, lng, scg, textsg = m.transcribe(pin, off, bw=None, int2word=i2w)
    greedy = m.eval_one_batch_with_greedy(m.device, dfeats, lens, i2w, None)
    # asr_prefetch_pcm pipeline: batch 2 is registered before batch 1 runs and staged during it
    pin2 = torch.from_numpy(np.concatenate(pcms[::-1])).pin_memory()
    off2 = np.zeros(len(pcms) + 1, dtype=np.int64)
    off2[1:] = np.cumsum([len(p) for p in pcms[::-1]])
    want2 = m.transcribe(pin2, off2, bw=cs["bw"], int2word=i2w)
    m.prefetch(pin, off, bw=cs["bw"])
    m.prefetch(pin2, off2, bw=cs["bw"])
    got1 = m.transcribe(pin, off, bw=cs["bw"], int2word=i2w)
    got2 = m.transcribe(pin2, off2, bw=cs["bw"], int2word=i2w)
    pre_ok = int(got1[3] == texts and np.array_equal(got1[2], sc) and got2[3] == want2[3]
                 and np.array_equal(got2[2], want2[2]) and got2[3] == texts[::-1])
    return {"fused_eq_staged": int(texts == list(staged.pred_text)),
            "prefetched_eq_direct": pre_ok,
            "resident_eq_host": int(texts2 == texts and np.array_equal(sc, sc2)),
            "score_eq": int(np.allclose(sc, np.array(staged.score, dtype=np.float32), rtol=0, atol=0)),
            "greedy_fused_eq_staged": int(textsg == list(greedy.pred_text))}


def check_batch_invariance(B=12, k=4, seconds=3.0, seed=900):
    """Size-independent property: an utterance's hypothesis does not depend on what else is in
    the batch (no cross-utterance reduction anywhere on the path)."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = get_model((1234, "sharp", 8.0), weights)
    _, i2w = vocab()
    rng = np.random.default_rng(seed)
    ns = [int(16000 * seconds * (0.5 + rng.random())) for _ in range(B)]
    pcms = [O.synth_pcm(seed + i, n) for i, n in enumerate(ns)]
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum(ns)
    tok, ln, sc, texts = m.transcribe(np.concatenate(pcms), off, bw=k, int2word=i2w)
    same = 0
    for i in (0, B // 2, B - 1):
        o1 = np.array([0, ns[i]], dtype=np.int64)
        t1, l1, s1, x1 = m.transcribe(pcms[i], o1, bw=k, int2word=i2w)
        same += int(x1[0] == texts[i] and abs(float(s1[0]) - float(sc[i])) <= 1e-4 * max(1.0, abs(float(sc[i]))))
    perm = rng.permutation(B)
    offp = np.zeros(B + 1, dtype=np.int64)
    offp[1:] = np.cumsum([ns[j] for j in perm])
    tp, lp, sp, xp = m.transcribe(np.concatenate([pcms[j] for j in perm]), offp, bw=k, int2word=i2w)
    perm_ok = int(all(xp[i] == texts[j] for i, j in enumerate(perm)))
    return {"single_eq_batch": same, "of": 3, "perm_invariant": perm_ok}


def check_attention_range(which, scale=400.0, k=4):
    """Attention pre-activations outside the product form's range (|x| > 20.8): scaling W_enc makes
    the keys large (flagged per utterance by keys_exp_kernel), scaling W_hidden the queries (flagged
    per step); both must fall back to the exact sum-then-exp scores and still match the oracle."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    name = {"keys": "attn_mechanism.W_enc", "queries": "attn_mechanism.W_hidden"}[which]
    weights["decoder_state_dict"][name] = weights["decoder_state_dict"][name] * scale
    m = get_model((1234, "sharp", 8.0, which, scale), weights)
    _, i2w = vocab()
    pcms = [O.synth_pcm(5100 + i, n) for i, n in enumerate((40000, 33000, 26000))]
    feats = [O.features(p) for p in pcms]
    lens = torch.tensor([f.size(0) for f in feats])
    enc = O.encoder_forward(weights, feats, lens)
    tr = {}
    o = O.beam_decode(weights, k, feats, lens, i2w, trace=tr)
    out = m.eval_one_batch_with_beam(m.device, k, feats, lens, None, i2w, second_pass=False)
    og = O.greedy_decode(weights, feats, lens, i2w)
    gr = m.eval_one_batch_with_greedy(m.device, feats, lens, i2w, None)
    off = np.zeros(len(pcms) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in pcms])
    tok, ln, sc, texts = m.transcribe(np.concatenate(pcms), off, bw=None, int2word=i2w)
    keys = O.attention_keys(weights, enc[0] if isinstance(enc, tuple) else enc)
    return {"beam_tokens_vs_oracle": int(m.last_beam_info["tokens"] == o["tokens"]),
            "beam_score_rel": float(max(abs(a - b) / max(1e-6, abs(b)) for a, b in zip(out.score, o["score"]))),
            "greedy_text_vs_oracle": int(list(gr.pred_text) == og["pred_text"]),
            "stream_greedy_text_vs_oracle": int(texts == og["pred_text"]),
            "max_abs_key": float(torch.as_tensor(keys).abs().max()),
            "oracle_min_margin": float(min(tr["min_margin"]))}


def check_graph_replay(B=6, k=4, n=40000, seed=700):
    """The beam loop is replayed from a CUDA graph once a batch shape repeats (eager, capture,
    replay on calls 1-3).  Batches with the same shape but different audio must give exactly what
    the eager loop gives (stage timing forces the eager loop), also after another shape intervened."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = get_model((1234, "sharp", 8.0), weights)
    _, i2w = vocab()
    off = np.arange(B + 1, dtype=np.int64) * n
    batches = [np.concatenate([O.synth_pcm(seed + 10 * j + i, n) for i in range(B)]) for j in range(4)]
    other = np.concatenate([O.synth_pcm(seed + 99 + i, n // 2) for i in range(B)])
    off2 = np.arange(B + 1, dtype=np.int64) * (n // 2)
    got = [m.transcribe(b, off, bw=k, int2word=i2w) for b in batches[:3]]
    m.transcribe(other, off2, bw=k, int2word=i2w)
    got.append(m.transcribe(batches[3], off, bw=k, int2word=i2w))
    m.stage_timing(True)
    try:
        want = [m.transcribe(b, off, bw=k, int2word=i2w) for b in batches]
    finally:
        m.stage_timing(False)
    same = sum(int(np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and np.array_equal(g[2], w[2]))
               for g, w in zip(got, want))
    distinct = int(len({tuple(g[3]) for g in got}) > 1)
    return {"replay_eq_eager": same, "of": len(batches), "batches_differ": distinct}


NEAR_TIE = 2e-6     # a candidate order may differ from the reference's only where the reference's own gap is below this ...
NOISE_RTOL = 1e-4   # ... or below twice the measured distance between the two implementations' scores at that step


def near_tie(gap, score, noise):
    """Is a differing decision a proven near-tie?  Beam scores are un-normalised fp32 sums of log-probs
    (model.py:836) that reach |s| ~ 50-300 with the test weights, and the logits behind them are only specified to
    1e-3 absolute (BASELINE.json north_star).  Two candidates whose reference scores are `gap` apart can legitimately
    swap when each implementation's scores carry an error of gap / 2.  `noise` is the largest |score_cuda -
    score_oracle| over the candidates (source beam, token) that appear in BOTH top-2k lists of that step, the
    swapped ones included: the swap is excused only if gap <= 2 * noise AND that noise itself is small (<= 1e-4
    relative to the score, ten times tighter than the 1e-3 relative tolerance on final scores) - or if the gap is
    below 2e-6 outright.  In words: the two implementations agree on every candidate's score to 1e-4 relative, and
    only candidates the oracle itself holds closer together than twice that disagreement may change places."""
    return gap < NEAR_TIE or (gap <= 2.0 * noise and noise <= NOISE_RTOL * max(1.0, abs(score)))


def _rel(a, b):
    return abs(float(a) - float(b)) / max(1.0, abs(float(b)))


def compare_with_oracle(m, weights, pcms, k, picks=None, lm_seed=None, lm_weight=0.0, length_weight=0.0,
                        temperature=1.0, group=32, label=""):
    """Decode ALL of `pcms` as one batch through the fused C-ABI path (PCM in, hypotheses out), then decode
    the PICKED utterances with the CPU oracle and compare, per utterance and per step: the consumed
    candidates of the top-2k list (token, source beam, score), back-pointers, active tokens, the whole list
    of finished hypotheses in (step, rank) order (asr_beam_nbest) and the final pick.

    Utterances of a batch interact only through the early stop (model.py:897-901), so the oracle decodes the
    picks alone with the batch's stop step passed in (`batch_stop_step`) and the check verifies that this
    stop step is consistent with every pick (its rank-0 </s> came no later).

    A difference is tolerated ONLY as a proven near-tie (near_tie(): the oracle's own gap at the first differing
    rank is within twice the measured score noise of that step, or < 2e-6); it is reported in `flips` with the
    step, rank, gap, score and noise.  Anything else lands in `bad`."""
    from chinese_asr_b200.gpd import gpd
    from chinese_asr_b200.lm import NGramLM
    w2i, i2w = vocab()
    B = len(pcms)
    picks = list(range(B)) if picks is None else sorted(set(int(i) for i in picks))
    s16 = pcms[0].dtype == np.int16
    lm_o = O.NGramLM(seed=lm_seed, word2int=w2i) if lm_seed else None
    lm_d = NGramLM(lm_o.tables(), w2i) if lm_o else None
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in pcms])
    gpd['temperature'] = temperature
    try:
        tok, ln, sc = m.transcribe(np.concatenate(pcms), off, bw=k, second_pass=lm_d is not None, lm_model=lm_d,
                                   lm_weight=lm_weight, length_weight=length_weight)
    finally:
        gpd['temperature'] = 1.
    info = m.decode_info()
    t = m.beam_trace(B, k)
    nbest = m.beam_nbest(B)
    S = info["stopped_at"]
    res = {"B": B, "picks": len(picks), "steps": info["steps"], "stopped_at": S, "exact": 0, "flips": [], "bad": [],
           "nonempty": 0, "finished_hyps": 0, "score_rel_max": 0.0, "cand_score_rel_max": 0.0,
           "min_margin": float("inf"), "lens": sorted({int(x) for x in ln[picks]})}
    K = 2 * k
    for g0 in range(0, len(picks), group):
        ids = picks[g0:g0 + group]
        feats = [O.features(O.pcm_from_int16(pcms[i]) if s16 else pcms[i]) for i in ids]
        lens = torch.tensor([f.size(0) for f in feats])
        tr = {}
        o = O.beam_decode(weights, k, feats, lens, i2w, second_pass=lm_o is not None, lm_model=lm_o,
                          lm_weight=lm_weight, length_weight=length_weight, temperature=temperature, trace=tr,
                          batch_stop_step=S)
        if o["steps"] != info["steps"]:
            res["bad"].append(("steps", ids[0], o["steps"], info["steps"]))
            continue
        n_act = len(tr.get("backptr", []))
        for j, i in enumerate(ids):
            tds = int(tr["top_done_step"][j])
            if (S >= 0 and not 0 <= tds <= S):
                res["bad"].append(("stop step inconsistent", i, tds, S))
                continue
            div = None
            for s in range(info["steps"]):
                o_tok, o_beam = tr["cand_tokens"][s][j].numpy(), tr["cand_beams"][s][j].numpy()
                o_sc = tr["cand_scores"][s][j].numpy()
                res["min_margin"] = min(res["min_margin"], float(tr["margin_utt"][s][j]))
                # consumed ranks: the top k (finished set, model.py:876-889) and, when the step forms an active
                # set, everything up to the k-th non-</s> candidate (model.py:904-909)
                used = k
                if s < n_act:
                    used = max(k, int(np.searchsorted(np.cumsum(o_tok != O.EOS), k)) + 1)
                same = (np.array_equal(t["cand_tokens"][s, i, :used], o_tok[:used])
                        and np.array_equal(t["cand_beams"][s, i, :used], o_beam[:used]))
                if same and s < n_act:
                    same = (np.array_equal(t["backptr"][s, i], tr["backptr"][s][j].numpy())
                            and np.array_equal(t["active_tokens"][s, i], tr["active_tokens"][s][j].numpy()))
                if not same:
                    bad_r = np.nonzero((t["cand_tokens"][s, i, :used] != o_tok[:used])
                                       | (t["cand_beams"][s, i, :used] != o_beam[:used]))[0]
                    r0 = int(bad_r[0]) if len(bad_r) else -1          # -1: candidates agree, back-pointers differ
                    # gap between the oracle's scores at the first differing rank and its neighbours, the score
                    # there, and how far the two implementations' scores are apart on the ranks that still agree
                    o_ext = np.append(o_sc, float(tr["next_score"][s][j]))      # + the best candidate left out of the 2k
                    nb_ = [abs(o_ext[r0] - o_ext[x]) for x in (r0 - 1, r0 + 1) if 0 <= x <= K]
                    gap = float(min(nb_)) if r0 >= 0 else 0.0
                    # noise: candidates are matched by IDENTITY (source beam, token) between the two top-2k lists of
                    # the step - the ranks before the first difference alone are too few samples (one, when the
                    # lists part at rank 1), and the swapped candidates themselves are the ones whose scores matter
                    o_id = {(int(b_), int(t_)): float(v_) for b_, t_, v_ in zip(o_beam, o_tok, o_sc)}
                    d_ = [abs(float(v_) - o_id[(int(b_), int(t_))])
                          for b_, t_, v_ in zip(t["cand_beams"][s, i], t["cand_tokens"][s, i], t["cand_scores"][s, i])
                          if (int(b_), int(t_)) in o_id]
                    noise = float(max(d_)) if d_ else 0.0
                    div = (s, float(tr["margin_utt"][s][j]), r0, gap, float(o_sc[max(r0, 0)]), noise)
                    break
                res["cand_score_rel_max"] = max(res["cand_score_rel_max"],
                                                float(np.max(np.abs(t["cand_scores"][s, i, :used] - o_sc[:used])
                                                             / np.maximum(1.0, np.abs(o_sc[:used])))))
            if div is not None:
                (res["flips"] if near_tie(div[3], div[4], div[5]) else res["bad"]).append(
                    dict(utt=i, step=div[0], oracle_margin=div[1], rank=div[2], gap=div[3], score=div[4], noise=div[5]))
                continue
            o_nb = o["nbest"].get(j, [])
            g_nb = nbest[i]
            nb_ok = len(o_nb) == len(g_nb) and all(a[0] == b[0] and _rel(a[1], b[1]) <= 1e-3 for a, b in zip(g_nb, o_nb))
            final_ok = tok[i, :ln[i]].tolist() == o["tokens"][j] and _rel(sc[i], o["score"][j]) <= 1e-3
            if not (nb_ok and final_ok):
                res["bad"].append(("nbest" if not nb_ok else "final", i, len(g_nb), len(o_nb)))
                continue
            res["exact"] += 1
            res["nonempty"] += int(ln[i] > 0)
            res["finished_hyps"] += len(o_nb)
            res["score_rel_max"] = max(res["score_rel_max"], _rel(sc[i], o["score"][j]))
    print(f"[parity {label}] B={B} k={k} picks={len(picks)} exact={res['exact']} flips={res['flips']} "
          f"bad={res['bad']} steps={info['steps']} stop={S} finished_hyps={res['finished_hyps']} "
          f"min_margin={res['min_margin']:.2e} score_rel={res['score_rel_max']:.1e} "
          f"cand_score_rel={res['cand_score_rel_max']:.1e}")
    return res


def synth_batch(secs, seed0, int16=True):
    mk = O.synth_pcm_int16 if int16 else O.synth_pcm
    return [mk(seed0 + i, int(16000 * s)) for i, s in enumerate(secs)]


def check_config_shape(B, k, seconds_list, lm_seed=None, wseed=1234, eos_bias=8.0, seed0=3000, lm_weight=0.3,
                       length_weight=2.0):
    """BASELINE.json configs at their real utterance lengths, fused path (PCM in) vs the oracle on the same
    seeded inputs, every utterance compared step by step (compare_with_oracle); greedy: texts."""
    weights = O.make_weights(wseed, "sharp", eos_bias=eos_bias)
    m = get_model((wseed, "sharp", eos_bias), weights)
    w2i, i2w = vocab()
    if k is not None:
        pcms = synth_batch(seconds_list, seed0, int16=False)
        return compare_with_oracle(m, weights, pcms, k, lm_seed=lm_seed, lm_weight=lm_weight if lm_seed else 0.0,
                                   length_weight=length_weight if lm_seed else 0.0, label=f"config B={B} k={k}")
    ns = [int(16000 * s) for s in seconds_list]
    pcms = [O.synth_pcm(seed0 + i, n) for i, n in enumerate(ns)]
    feats = [O.features(p) for p in pcms]
    lens = torch.tensor([f.size(0) for f in feats])
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum(ns)
    o = O.greedy_decode(weights, feats, lens, i2w)
    tok, ln, sc, texts = m.transcribe(np.concatenate(pcms), off, bw=None, int2word=i2w)
    same = [a == b for a, b in zip(texts, o["pred_text"])]
    return {"B": B, "exact": int(sum(same)), "lens_min": int(lens.min()), "lens_max": int(lens.max())}


def rec_chunk_edges(B):
    """First / last sorted rank of every recurrence chunk (encoder_tc3.cu: NB = 16..128 sequences per
    cluster, at least 7 chunks per direction) - equal-length batches keep their order when sorted."""
    nb = 16
    while 7 * nb < B and nb < 128:
        nb += 16
    edges = set()
    for c0 in range(0, B, nb):
        edges.update((c0, min(B, c0 + nb) - 1))
    return sorted(edges)


def check_bench_shape(B=512, k=8, seconds=10.0, n_picks=32, seed0=41000):
    """The bench workload (BASELINE.json configs[4] per GPU: 512 x 10 s, bw=8) with sharp weights so that
    decisions are well separated: 32 picked utterances (first / last of every recurrence chunk + random)
    against the oracle."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = get_model((1234, "sharp", 8.0), weights)
    pcms = synth_batch([seconds] * B, seed0)
    picks = rec_chunk_edges(B)
    rng = np.random.default_rng(seed0)
    while len(set(picks)) < n_picks:
        picks.append(int(rng.integers(0, B)))
    return compare_with_oracle(m, weights, pcms, k, picks=picks, label="bench shape")


def check_odd_beam_widths():
    """Beam widths that are not an instantiation size (the kernels are built for K = 1, 4, 8, 16 beam slots; bw = 3,
    6, 11 run them with dead slots): small batches (two-phase attention kernel) with every utterance compared, and
    one batch of 300 utterances (streaming attention kernel, partial-beam instantiation) with 10 picks."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = get_model((1234, "sharp", 8.0), weights)
    out = {}
    for k in (3, 6, 11):
        pcms = synth_batch([2.0, 3.1, 2.4, 4.0, 2.2, 3.6], 52000 + k)
        out[f"small_k{k}"] = compare_with_oracle(m, weights, pcms, k, label=f"odd beam k={k}")
    B = 300
    pcms = synth_batch([2.0] * B, 53000)
    rng = np.random.default_rng(53)
    picks = [0, B - 1] + [int(x) for x in rng.integers(0, B, size=8)]
    out["stream_k6"] = compare_with_oracle(m, weights, pcms, 6, picks=picks, label="odd beam k=6 B=300")
    return out


def check_config3_full(B=256, k=16, n_picks=24, seed0=26000):
    """BASELINE.json configs[2] at full size: bw=16, 256 utterances of mixed 2-20 s; 24 picks including the
    shortest and the longest against the oracle, plus bit-reproducibility of the whole batch."""
    weights = O.make_weights(77, "sharp", eos_bias=9.0)
    m = get_model((77, "sharp", 9.0), weights)
    rng = np.random.default_rng(2600)
    secs = rng.integers(2, 21, size=B).tolist()
    pcms = synth_batch(secs, seed0)
    picks = [int(np.argmin(secs)), int(np.argmax(secs)), 0, B - 1]
    while len(set(picks)) < n_picks:
        picks.append(int(rng.integers(0, B)))
    r = compare_with_oracle(m, weights, pcms, k, picks=picks, group=12, label="config 3 full")
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in pcms])
    x = np.concatenate(pcms)
    a, b = m.transcribe(x, off, bw=k), m.transcribe(x, off, bw=k)
    r["reproducible"] = int(all(np.array_equal(p, q) for p, q in zip(a, b)))
    r["secs_min"], r["secs_max"] = int(min(secs)), int(max(secs))
    return r


# ---------------------------------------------------------------------------------------------
# rows either side of the hot path
def check_frontend():
    """16-bit ingest, loader CMVN (eps 1e-7), batch_audio / collate_fn, the AudioLoader pipeline."""
    import wave
    import tempfile
    from tests.cases import FRONTEND, load_frontend_golden
    from chinese_asr_b200 import data
    g = load_frontend_golden()
    w = O.make_weights(1234, "plain")
    m = get_model((1234, "plain", None), w)
    pcm16 = [O.synth_pcm_int16(s, n) for s, n in zip(FRONTEND["seeds"], FRONTEND["nsamp"])]
    res = {}
    raw16 = m.features(pcm16, normalise=False)                                  # int16 converted on the device
    raw32 = m.features([O.pcm_from_int16(x) for x in pcm16], normalise=False)   # host conversion (fast_read)
    res["s16_eq_f32_bitwise"] = int(all(torch.equal(a, b) for a, b in zip(raw16, raw32)))
    fused = m.features(pcm16, normalise=True, eps=1e-7)                         # loader path: features + CMVN fused
    normed, lens = data.AudioLoader.batch_audio([r.clone() for r in raw16], m)  # batch_audio on existing features
    res["lens_ok"] = int(lens.dtype == torch.int32 and lens.tolist() == g["fe_lens"].tolist())
    for i in range(len(pcm16)):
        rows = g[f"fe{i}_rows"]
        res[f"fe{i}_raw_vs_ref"] = maxabs(raw16[i][rows], g[f"fe{i}_raw"])
        res[f"fe{i}_fused_vs_ref"] = maxabs(fused[i][rows], g[f"fe{i}_norm"])
        res[f"fe{i}_batch_audio_vs_ref"] = maxabs(normed[i][rows], g[f"fe{i}_norm"])
        res[f"fe{i}_batch_audio_vs_oracle"] = maxabs(normed[i], O.batch_audio([raw16[i].cpu()])[0][0])
    t, l2, text = data.AudioLoader.collate_fn([(r, [1, 2]) for r in raw16], m)
    res["collate_ok"] = int(text == [[1, 2]] * len(raw16) and all(torch.equal(a, b) for a, b in zip(t, normed)))
    # eps really is an argument: 1e-6 (main.py:37) differs from 1e-7 and matches the oracle's cmvn
    e6 = m.features(pcm16[:1], normalise=True, eps=1e-6)[0]
    res["eps6_vs_oracle"] = maxabs(e6, O.cmvn(O.features(O.pcm_from_int16(pcm16[0]), normalise=False), 1e-6))
    # the loader: WAV files on disk -> batches of (t, lens, text)
    _, i2w = vocab()
    w2i = vocab()[0]

    class AB:
        word2int, int2word = w2i, i2w
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, x in enumerate(pcm16):
            p = os.path.join(d, f"u{i}.wav")
            with wave.open(p, "wb") as wf:
                wf.setnchannels(1)
                wf.setsampwidth(2)
                wf.setframerate(16000)
                wf.writeframes(x.tobytes())
            paths.append(p)
        texts = [[i2w[10 + i], i2w[20 + i]] for i in range(len(paths))]
        ld = data.AudioLoader(data.AudioDst(AB, "eval", "dev", path_list=paths, text_list=texts), m, batch_size=3)
        got = list(ld.loader)
    ok = len(got) == 2 and [len(b[0]) for b in got] == [3, 1] and got[0][2] == [[10, 20], [11, 21], [12, 22]]
    flat = [t for b in got for t in b[0]]
    res["loader_ok"] = int(ok and all(torch.equal(a, b) for a, b in zip(flat, fused)))
    return res


def check_wer():
    from tests.cases import load_frontend_golden, wer_token_pairs
    g = load_frontend_golden()
    _, i2w = vocab()
    m = get_model((1234, "plain", None), O.make_weights(1234, "plain"))
    pairs = wer_token_pairs()
    hyp = [h for h, _ in pairs]
    ref = [r for _, r in pairs]
    per = m.wer(ref, i2w, hyp=hyp)
    want = g["wer_norm"]
    res = {"per_utt_max_abs": float(np.abs(np.asarray(per) - want).max())}
    dist = [round(p * len("".join(i2w[t] for t in r))) for p, r in zip(per, ref)]
    res["dist_mismatch"] = int(sum(int(a != b) for a, b in zip(dist, g["wer_dist"].tolist())))
    # references given as strings
    per_s = m.wer(["".join(i2w[t] for t in r) for r in ref], i2w, hyp=hyp)
    res["string_refs_same"] = int(per_s == per)
    return res


def check_driver_wer(cname):
    """WER reported by the decode drivers when `text` is given (model.py:595-598, 982-985): the
    hypotheses are scored on the device where the decode left them."""
    from tests.cases import load_frontend_golden
    g = load_frontend_golden()
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    _, i2w = vocab()
    _, feats, lens = case_inputs(cs)
    if cs["bw"] is None:
        hyp = O.greedy_decode(weights, feats, lens, i2w)["tokens"]
    else:
        hyp = O.beam_decode(weights, cs["bw"], feats, lens, i2w)["tokens"]
    refs = [O.synth_reference_text(7000 + i, h) for i, h in enumerate(hyp)]
    if cs["bw"] is None:
        out = m.eval_one_batch_with_greedy(m.device, feats, lens, i2w, refs)
    else:
        out = m.eval_one_batch_with_beam(m.device, cs["bw"], feats, lens, refs, i2w, second_pass=False)
    res = {"wer": float(out.wer), "ref_wer": float(g[cname + "_wer"]),
           "text_ok": int(list(out.text) == g[cname + "_wer_text"].tolist())}

    class OneBatch:
        loader = [(feats, torch.IntTensor(lens.tolist()), refs)] * 2
    tm = m.test_model(OneBatch, i2w, bw=cs["bw"])
    res["test_model_wer"] = tm["wer"]
    res["test_model_n"] = tm["n"]
    res["test_model_error_rate"] = tm["error_rate"]
    res["oracle_error_rate"] = float(np.mean([h != r for h, r in zip(hyp, refs)]))
    return res


def check_wide_recurrence(B, k=4, group=40, seed=5200, n_picks=16):
    """Batches of 561..896 utterances run the recurrence with 96 / 128 sequences per cluster (one round of 14
    clusters).  Picked utterances (first / last sorted rank of every chunk + random) against the oracle, and the
    size-independent property: every hypothesis is bit-identical to the one decoded in a batch of `group`."""
    weights = O.make_weights(1234, "sharp", eos_bias=8.0)
    m = get_model((1234, "sharp", 8.0), weights)
    rng = np.random.default_rng(seed)
    ns = [int(16000 * (1.2 + 1.3 * rng.random())) for _ in range(B)]
    pcms = [O.synth_pcm_int16(seed + i, n) for i, n in enumerate(ns)]
    order = np.argsort(-np.array([int(O.num_frames(n)) // 3 for n in ns]), kind="stable")     # the library's sort
    picks = [int(order[r]) for r in rec_chunk_edges(B)][:n_picks]
    while len(set(picks)) < n_picks:
        picks.append(int(rng.integers(0, B)))
    res = compare_with_oracle(m, weights, pcms, k, picks=picks, label=f"wide recurrence B={B}")
    pcm = np.concatenate(pcms)
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum(ns)
    tok, ln, sc = m.transcribe(pcm, off, bw=k)
    stop = m.decode_info()["stopped_at"]
    same = compared = 0
    score_rel, differ = 0.0, []
    for g0 in range(0, B, group):
        g1 = min(B, g0 + group)
        t1, l1, s1 = m.transcribe(pcm[off[g0]:off[g1]], off[g0:g1 + 1] - off[g0], bw=k)
        if m.decode_info()["stopped_at"] != stop:
            continue            # utterances interact through the early stop: only equal stop steps are comparable
        compared += g1 - g0
        for i in range(g1 - g0):
            if l1[i] == ln[g0 + i] and (t1[i] == tok[g0 + i]).all():
                same += 1
                score_rel = max(score_rel, _rel(s1[i], sc[g0 + i]))
            else:
                differ.append(g0 + i)
    # Small batches take the frame-split attention kernel and other GEMM tile shapes, so scores differ in the last
    # bits and a near-tie may resolve differently: every utterance whose tokens differ is checked against the oracle
    # in BOTH batchings (exact, or a proven near-tie) - never waved through.
    bad = []
    for i in differ:
        g0 = (i // group) * group
        g1 = min(B, g0 + group)
        big = compare_with_oracle(m, weights, pcms, k, picks=[i], label=f"wide recurrence B={B}, differing utt {i}")
        small = compare_with_oracle(m, weights, pcms[g0:g1], k, picks=[i - g0], label=f"batch of {g1 - g0}, utt {i}")
        bad += big["bad"] + small["bad"]
    res.update({"same": same, "of": compared, "differ": differ, "differ_bad": bad, "self_score_rel": score_rel,
                "len_spread": int(ln.max() - ln.min())})
    return res


# ---------------------------------------------------------------------------------------------
# boundary: lm_model duck typing (model.py:749-763, main.py:79-85) and ARPA vocabularies
class ScoreOnlyLM:
    """What the reference is handed: any object with .score(sentence, bos=True) (a kenlm.LanguageModel)."""

    def __init__(self, lm):
        self._lm = lm
        self.calls = 0

    def score(self, sentence, bos=True):
        self.calls += 1
        return self._lm.score(sentence, bos=bos)


def check_host_lm(cname="beam8lm"):
    """An lm_model offering only .score() is rescored on the host from asr_beam_nbest with the rule of
    model.py:749-763 and must pick exactly what the device tables pick (and what the reference picked)."""
    from chinese_asr_b200.lm import NGramLM
    g = load_golden()
    cs = CASES[cname]
    weights = case_weights(cs)
    m = get_model(wkey(cs), weights)
    w2i, i2w = vocab()
    _, feats, lens = case_inputs(cs)
    lm_o = O.NGramLM(seed=cs["lm"], word2int=w2i)
    kw = dict(second_pass=True, lm_weight=cs["lm_weight"], length_weight=cs["length_weight"])
    dev = m.eval_one_batch_with_beam(m.device, cs["bw"], feats, lens, None, i2w, lm_model=NGramLM(lm_o.tables(), w2i), **kw)
    plain = ScoreOnlyLM(lm_o)
    host = m.eval_one_batch_with_beam(m.device, cs["bw"], feats, lens, None, i2w, lm_model=plain, **kw)
    none = m.eval_one_batch_with_beam(m.device, cs["bw"], feats, lens, None, i2w, second_pass=False)
    off = np.zeros(len(feats) + 1, dtype=np.int64)
    pcms = [O.synth_pcm(s_, n) for s_, n in zip(cs["seeds"], cs["nsamp"])]
    off[1:] = np.cumsum([len(p) for p in pcms])
    fused = m.transcribe(np.concatenate(pcms), off, bw=cs["bw"], lm_model=ScoreOnlyLM(lm_o), int2word=i2w, **kw)
    nb = m.beam_nbest(len(feats))
    return {"host_eq_device": int(list(host.pred_text) == list(dev.pred_text) and list(host.score) == list(dev.score)),
            "host_eq_ref": int(list(host.pred_text) == list(g[cname + "_text"])),
            "fused_eq_device": int(fused[3] == list(dev.pred_text)),
            "lm_changes_a_pick": int(list(none.pred_text) != list(dev.pred_text)),
            "lm_calls": plain.calls, "ref_lm_calls": int(len(g[cname + "_lm_seen"])),
            "nbest_counts": [len(x) for x in nb]}


ARPA = """\\data\\
ngram 1={n1}
ngram 2={n2}
ngram 3={n3}

\\1-grams:
{uni}

\\2-grams:
{bi}

\\3-grams:
{tri}

\\end\\
"""


def arpa_backoff_score(grams, sentence):
    """kenlm's .score(sentence, bos=True, eos=True) restated on the ARPA entries themselves (dicts keyed by word
    tuples -> (log10 p, backoff)); words without a unigram score as <unk>."""
    def p(ctx, w):
        if not ctx:
            return grams[1][(w,)][0]
        key = tuple(ctx) + (w,)
        if key in grams[len(key)]:
            return grams[len(key)][key][0]
        return grams[len(ctx)].get(tuple(ctx), (0.0, 0.0))[1] + p(ctx[1:], w)
    ctx, total = ["<s>"], 0.0
    for w in sentence.split() + ["</s>"]:
        w = w if (w,) in grams[1] else "<unk>"
        total += p(ctx[-2:], w)
        ctx.append(w)
    return total


def check_arpa_oov():
    """ARPA file whose vocabulary differs from dict.pkl in both directions: a word dict.pkl lacks
    must not overwrite <unk>; a dict.pkl token the ARPA lacks scores as <unk> in every n-gram position."""
    import tempfile
    from chinese_asr_b200.lm import NGramLM
    w2i, i2w = vocab()
    a, b, c, d, miss = (i2w[i] for i in (10, 11, 12, 13, 14))        # `miss` has no unigram in the ARPA
    grams = {1: {("<unk>",): (-2.5, -0.3), ("<s>",): (-99.0, -0.4), ("</s>",): (-1.1, 0.0), (a,): (-1.3, -0.2),
                 (b,): (-1.6, -0.25), (c,): (-1.9, -0.1), (d,): (-2.1, 0.0), ("ZZZ",): (-0.7, -0.9)},
             2: {("<s>", a): (-0.5, -0.15), (a, b): (-0.6, -0.05), (b, "<unk>"): (-0.9, -0.07), ("<unk>", c): (-0.8, -0.02),
                 (a, "ZZZ"): (-0.1, -0.6), ("ZZZ", b): (-0.2, 0.0), (c, "</s>"): (-0.3, 0.0)},
             3: {("<s>", a, b): (-0.25, 0.0), (a, b, "<unk>"): (-0.35, 0.0), (b, "<unk>", c): (-0.45, 0.0),
                 (a, "ZZZ", b): (-0.05, 0.0)}}
    fmt = lambda n: "\n".join(f"{v[0]}\t{' '.join(k)}" + (f"\t{v[1]}" if n < 3 else "") for k, v in grams[n].items())
    text = ARPA.format(n1=len(grams[1]), n2=len(grams[2]), n3=len(grams[3]), uni=fmt(1), bi=fmt(2), tri=fmt(3))
    with tempfile.NamedTemporaryFile("w", suffix=".arpa", delete=False, encoding="utf-8") as f:
        f.write(text)
    lm = NGramLM.from_arpa(f.name, w2i)
    os.unlink(f.name)
    m = get_model((1234, "plain", None), O.make_weights(1234, "plain"))
    m._lm = None
    m.set_lm(lm)
    sents = [f"{a} {b} {miss} {c}", f"{miss}", f"{a} {b}", f"{d} {miss} {miss} {a}", f"{c}", "", f"{a} {b} {i2w[3]} {c}",
             f"{b} {miss} {c} {a} {b} {miss}"]
    dev = [lm.score(s_) for s_ in sents]
    ref = [arpa_backoff_score({n: {k: v for k, v in gs.items() if "ZZZ" not in k} for n, gs in grams.items()}, s_)
           for s_ in sents]
    t = lm.tables()
    m._lm = None                       # the next caller uploads its own tables
    return {"max_abs": float(np.max(np.abs(np.array(dev) - np.array(ref)))),
            "unk_kept": int(t["uni_logp"][3] == np.float32(-2.5) and t["uni_bo"][3] == np.float32(-0.3)),
            "missing_maps_to_unk": int(t["id_map"][14] == 3 and t["id_map"][10] == 10)}


def check_convert_audio(tmp_dir):
    """Device resampler (asr_convert_audio through data.convert_pcm / data.convert_audio) vs O.convert_audio."""
    import wave
    from chinese_asr_b200 import data as D
    rng = np.random.default_rng(77)
    cases = {}

    def speechlike(n, sr):
        t = np.arange(n) / sr
        x = sum(a * np.sin(2 * np.pi * f * t + p) for a, f, p in
                zip((0.3, 0.2, 0.1, 0.05), (180.0, 950.0, 2900.0, 6100.0), (0.1, 1.3, 2.2, 0.7)))
        return (x + 0.02 * rng.standard_normal(n)).astype(np.float32)

    inputs = {
        "stereo_44k1_f32": (np.stack([speechlike(44100, 44100), 0.5 * speechlike(44100, 44100)], 1), 44100),
        "mono_48k_s16": (np.clip(np.rint(speechlike(30011, 48000) * 20000), -32768, 32767).astype(np.int16), 48000),
        "mono_8k_s16": (np.clip(np.rint(speechlike(9001, 8000) * 12000), -32768, 32767).astype(np.int16), 8000),
        "mono_22k05_f32": (speechlike(22050, 22050), 22050),
        "stereo_16k_s16": (np.clip(np.rint(np.stack([speechlike(16000, 16000)] * 2, 1) * 9000), -32768, 32767).astype(np.int16), 16000),
    }
    for name, (x, sr) in inputs.items():
        dev = D.convert_pcm(x, sr)
        ora = O.convert_audio(x, sr)
        n = min(len(dev), len(ora))
        d = np.abs(dev[:n].astype(np.int32) - ora[:n].astype(np.int32))
        cases[name] = (len(dev), len(ora), int(d.max()) if n else 0, float((d > 0).mean()) if n else 0.0)
    # WAV round trip like main.py:30
    x, sr = inputs["stereo_44k1_f32"]
    src = os.path.join(tmp_dir, "in.wav")
    with wave.open(src, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(sr)
        xi = np.clip(np.rint(x * 32767), -32768, 32767).astype("<i2")
        w.writeframes(xi.tobytes())
    out = D.convert_audio(src, os.path.join(tmp_dir, "tmp.wav"))
    with wave.open(out, "rb") as w:
        rate, ch, raw = w.getframerate(), w.getnchannels(), w.readframes(w.getnframes())
    got = np.frombuffer(raw, dtype="<i2")
    want = D.convert_pcm(xi, sr)
    return {"cases": cases, "wav_rate": rate, "wav_channels": ch, "wav_equal": bool(np.array_equal(got, want))}
