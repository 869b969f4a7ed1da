#!/usr/bin/env python
"""Generate tests/golden/ref_golden.npz by running the UNMODIFIED reference (imported from
/root/reference behind tools/ref_shim.py) on the seeded inputs / weights defined in
oracle/asr_oracle.py.  Runs only in the build container (the reference tree is absent on the
GPU box); the fixture it writes is committed.

    python tools/make_golden.py            # writes tests/golden/ref_golden.npz (+ dict.pkl copy)
"""
import io
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402
from oracle import asr_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# the seeded cases; tests/cases.py re-creates the inputs from the same table
from tests.cases import CASES, case_inputs, case_weights  # noqa: E402


def ref_model(ref, weights):
    """Model() + Model.load through a real checkpoint file (weight hand-off, model.py:357-369)."""
    m = ref.model.Model()
    path = "/tmp/_golden_ckpt.pt"
    torch.save(weights, path)
    m.load(path)
    m.model.eval()
    return m


def ref_features(ref, pcm, normalise=True):
    ref_shim.register_wav("mem.wav", pcm)
    with ref_shim.legacy_torch():
        f = ref.data.get_log_mel(False, "mem.wav", ref.audio_base.ms, ref.audio_base.window, False)
    if normalise:
        f = (f - f.mean(dim=0)) / (f.std(dim=0) + 1e-6)         # main.py:37
    return f


class TopkRecorder:
    def __init__(self):
        self.calls = []
        self._orig = torch.topk

    def __enter__(self):
        def rec(*a, **kw):
            r = self._orig(*a, **kw)
            self.calls.append((r[0].clone(), r[1].clone(), kw.get("largest", True)))
            return r
        torch.topk = rec
        return self

    def __exit__(self, *e):
        torch.topk = self._orig


class RecordingLM:
    def __init__(self, lm):
        self.lm = lm
        self.seen = []

    def score(self, s, bos=True):
        v = self.lm.score(s, bos=bos)
        self.seen.append((s, v))
        return v


def main():
    ref = ref_shim.load_reference()
    out = {}
    i2w = ref.audio_base.int2word
    w2i = ref.audio_base.word2int
    shutil.copyfile(os.path.join(ref_shim.REF, "dict.pkl"), os.path.join(GOLD, "dict.pkl"))

    # ---- known-answer of the reference itself: encoder.py:636-652 ---------------------------
    enc = ref.encoder.RNNEncoder()
    for p in enc.parameters():
        torch.nn.init.ones_(p)
    lens = [10, 8, 23, 14]
    with torch.no_grad():
        y = enc([torch.ones(n, 720) for n in lens], torch.tensor(lens))
    out["kat_rnn_sums"] = np.array([y[0].sum().item(), y[2][0].sum().item(), y[2][1].sum().item()])
    print("KAT encoder.py:652 ->", out["kat_rnn_sums"], "(comment says 110345.5, 2048, 28160)")

    # ---- constants --------------------------------------------------------------------------
    out["fb"] = ref.audio_base.ms.fb.numpy()
    out["window"] = ref.audio_base.window.numpy()

    # ---- features ---------------------------------------------------------------------------
    # 10 s and 20 s are the utterance lengths of BASELINE.json's configs (332 / 665 encoder frames)
    for name, seed, n in (("feat_2s", 11, 32000), ("feat_5s", 12, 80000), ("feat_odd", 13, 20011),
                          ("feat_10s", 14, 160000), ("feat_20s", 15, 320000)):
        pcm = O.synth_pcm(seed, n)
        f_raw = ref_features(ref, pcm, normalise=False)
        f = ref_features(ref, pcm, normalise=True)
        rows = np.arange(f.size(0)) if f.size(0) <= 70 else np.linspace(0, f.size(0) - 1, 24).astype(int)
        out[name + "_rows"] = rows
        out[name + "_raw"] = f_raw.numpy()[rows]
        out[name + "_cmvn"] = f.numpy()[rows]
        out[name + "_L"] = np.array(f.size(0))
        d = (O.features(pcm) - f).abs().max().item()
        print(f"{name}: L={f.size(0)} oracle-vs-ref max|d|={d:.3e}")
    z = np.zeros(8000, dtype=np.float32)                         # exact-zero -> eps branch (data.py:223)
    out["feat_zero_raw"] = ref_features(ref, z, normalise=False).numpy()

    # ---- model cases ------------------------------------------------------------------------
    for cname, cs in CASES.items():
        weights = case_weights(cs)
        m = ref_model(ref, weights)
        pcms, feats, lens = case_inputs(cs)
        rfeats = [ref_features(ref, p) for p in pcms]
        lens_t = torch.tensor([f.size(0) for f in rfeats])
        assert lens_t.tolist() == lens.tolist()
        with torch.no_grad():
            eo = m.encoder(rfeats, lens_t)
            keys, _ = m.attn_mechanism.compute_key_value(eo.out)
        out[cname + "_enc_out"] = eo.out.numpy()[:: cs.get("enc_stride", 1)]
        out[cname + "_enc_h"] = eo.state[0].numpy()
        out[cname + "_enc_c"] = eo.state[1].numpy()
        out[cname + "_keys"] = keys.numpy()[:: cs.get("enc_stride", 1)]
        print(f"{cname}: lens={lens.tolist()} variant={cs['variant']}")
        if cs["bw"] is None:
            with ref_shim.legacy_torch():
                r = m.eval_one_batch_with_greedy(m.device, rfeats, lens_t, i2w, None)
            out[cname + "_text"] = np.array(r.pred_text)
            out[cname + "_score"] = np.array(r.score, dtype=np.float64)
            out[cname + "_text_len"] = r.text_len.numpy()
            out[cname + "_align0"] = r.alignment[0].numpy()
            g = O.greedy_decode(weights, rfeats, lens_t, i2w)
            print("   ref  :", r.pred_text, r.score)
            print("   orcl :", g["pred_text"], g["score"])
        else:
            lm = None
            if cs.get("lm"):
                lm = RecordingLM(O.NGramLM(seed=cs["lm"], word2int=w2i))
            ref.gpd["temperature"] = cs.get("temperature", 1)
            with ref_shim.legacy_torch(), TopkRecorder() as tk:
                r = m.eval_one_batch_with_beam(m.device, cs["bw"], rfeats, lens_t, None, i2w,
                                               second_pass=lm is not None, lm_model=lm,
                                               lm_weight=cs.get("lm_weight", 0.0),
                                               length_weight=cs.get("length_weight", 0.0))
            ref.gpd["temperature"] = 1
            out[cname + "_text"] = np.array(r.pred_text)
            out[cname + "_score"] = np.array(r.score, dtype=np.float64)
            cand = [c for c in tk.calls if c[2] and c[0].size(1) == 2 * cs["bw"]]
            act = [c for c in tk.calls if not c[2]]
            out[cname + "_cand_scores"] = torch.stack([c[0] for c in cand]).numpy()
            out[cname + "_cand_index"] = torch.stack([c[1] for c in cand]).numpy()
            out[cname + "_active"] = (torch.stack([c[1] for c in act]).numpy() if act
                                      else np.zeros((0, len(pcms), cs["bw"]), dtype=np.int64))
            if lm is not None:
                out[cname + "_lm_seen"] = np.array([s for s, _ in lm.seen])
                out[cname + "_lm_vals"] = np.array([v for _, v in lm.seen], dtype=np.float64)
            tr = {}
            lm2 = O.NGramLM(seed=cs["lm"], word2int=w2i) if cs.get("lm") else None
            g = O.beam_decode(weights, cs["bw"], rfeats, lens_t, i2w, second_pass=lm2 is not None,
                              lm_model=lm2, lm_weight=cs.get("lm_weight", 0.0),
                              length_weight=cs.get("length_weight", 0.0),
                              temperature=cs.get("temperature", 1), trace=tr)
            same = g["pred_text"] == list(r.pred_text)
            ds = max(abs(a - b) / max(1e-9, abs(b)) for a, b in zip(g["score"], r.score))
            print(f"   steps={len(cand)} oracle tokens identical={same} score rel diff={ds:.2e} "
                  f"min margin={min(tr['min_margin']):.2e} fallback={g['fallback']} "
                  f"nfinished={[len(v) for v in g['nbest'].values()]}")
            print("   ref  :", [t[:12] for t in r.pred_text])
    np.savez_compressed(os.path.join(GOLD, "ref_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(GOLD, "ref_golden.npz"))
    print(f"wrote ref_golden.npz: {sz / 1e6:.2f} MB, {len(out)} arrays")


if __name__ == "__main__":
    main()
