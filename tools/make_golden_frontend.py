#!/usr/bin/env python
"""Generate tests/golden/ref_frontend.npz: outputs of the UNMODIFIED reference (imported from
/root/reference behind tools/ref_shim.py) for the rows either side of the hot path
(SURVEY.md section 8f rows 1 and 3):

  * fast_read -> get_log_mel on 16-bit PCM (data.py:109-121, 167-280; the soundfile stub serves
    int16 / 32768 as float32, libsndfile's documented conversion),
  * AudioLoader.batch_audio / collate_fn (data.py:496-518; eps 1e-7),
  * get_wer_python (util.py:186-234, the reference's own DP form of get_wer) on seeded string pairs,
  * the WER the decode drivers report when `text` is given (model.py:595-598, 982-985), with
    Levenshtein.distance served by get_wer_python(normalize=False).

Runs only in the build container; the fixture it writes is committed.

    python tools/make_golden_frontend.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402
from oracle import asr_oracle as O  # noqa: E402
from make_golden import ref_features, ref_model  # noqa: E402
from tests.cases import CASES, FRONTEND, case_inputs, case_weights, wer_pairs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_shim.load_reference()
    i2w = ref.audio_base.int2word
    out = {}

    # Levenshtein.distance: the reference's own DP (util.py:186-234); its assert forbids empty refs
    # (util.py:9 binds `distance` at import time, so the name is replaced inside the reference's util module)
    ref.util.distance = lambda a, b: (len(a) if len(b) == 0 else int(ref.util.get_wer_python(a, b, normalize=False)))

    # ---- 16-bit ingest + batch_audio / collate_fn -------------------------------------------------
    pcm16 = [O.synth_pcm_int16(s, n) for s, n in zip(FRONTEND["seeds"], FRONTEND["nsamp"])]
    raw = [ref_features(ref, O.pcm_from_int16(x), normalise=False) for x in pcm16]
    texts = [[int(t) for t in np.random.default_rng(900 + i).integers(4, 5004, size=5 + i)] for i in range(len(raw))]
    assert ref.gpd["normalize"] and ref.gpd["encoder_type"] == "LSTM"
    t, lens, text = ref.data.AudioLoader.collate_fn([(f, tx) for f, tx in zip(raw, texts)])
    t2, lens2 = ref.data.AudioLoader.batch_audio(list(raw))
    assert all(torch.equal(a, b) for a, b in zip(t, t2)) and torch.equal(lens, lens2) and text == texts
    out["fe_lens"] = lens.numpy()
    for i, (r, n) in enumerate(zip(raw, t)):
        rows = np.linspace(0, r.size(0) - 1, min(r.size(0), 16)).astype(int)
        out[f"fe{i}_rows"] = rows
        out[f"fe{i}_raw"] = r.numpy()[rows]
        out[f"fe{i}_norm"] = n.numpy()[rows]
        d_raw = (O.delta_stack(O.log_mel_frames(O.pcm_from_int16(pcm16[i]))) - r).abs().max().item()
        d_nrm = (O.batch_audio([r])[0][0] - n).abs().max().item()
        print(f"fe{i}: L={r.size(0)} oracle-vs-ref raw {d_raw:.2e} batch_audio {d_nrm:.2e}")

    # ---- get_wer_python on seeded string pairs ------------------------------------------------------
    pairs = wer_pairs(i2w)
    out["wer_dist"] = np.array([int(ref.util.get_wer_python(p, r, normalize=False)) for p, r in pairs])
    out["wer_norm"] = np.array([float(ref.util.get_wer_python(p, r)) for p, r in pairs])
    out["wer_get_wer"] = np.array([float(ref.util.get_wer(p, r)) for p, r in pairs])
    mine = [O.edit_distance(p, r) for p, r in pairs]
    print("wer pairs:", out["wer_dist"].tolist(), "oracle identical:", mine == out["wer_dist"].tolist())

    # ---- WER reported by the decode drivers -------------------------------------------------------------
    for cname in ("greedy3", "beam4"):
        cs = CASES[cname]
        weights = case_weights(cs)
        m = ref_model(ref, weights)
        pcms, feats, lens_t = case_inputs(cs)
        rfeats = [ref_features(ref, p) for p in pcms]
        if cs["bw"] is None:
            g = O.greedy_decode(weights, rfeats, lens_t, i2w)
            hyp = g["tokens"]
        else:
            hyp = O.beam_decode(weights, cs["bw"], rfeats, lens_t, i2w)["tokens"]
        refs = [O.synth_reference_text(7000 + i, h) for i, h in enumerate(hyp)]
        with ref_shim.legacy_torch():
            if cs["bw"] is None:
                r = m.eval_one_batch_with_greedy(m.device, rfeats, lens_t, i2w, [list(x) for x in refs])
            else:
                r = m.eval_one_batch_with_beam(m.device, cs["bw"], rfeats, lens_t, [list(x) for x in refs], i2w,
                                               second_pass=False)
        out[cname + "_wer"] = np.array(float(r.wer))
        out[cname + "_wer_text"] = np.array(list(r.text))
        mean, per = O.batch_wer(hyp, refs, i2w)
        print(f"{cname}: ref wer={float(r.wer):.6f} oracle={mean:.6f} per-utt={['%.3f' % x for x in per]}")
    np.savez_compressed(os.path.join(GOLD, "ref_frontend.npz"), **out)
    print("wrote ref_frontend.npz:", os.path.getsize(os.path.join(GOLD, "ref_frontend.npz")), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
