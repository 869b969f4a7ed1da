"""Seeded parity cases shared by tools/make_golden.py (reference side, build container only) and
the tests (oracle / CUDA side).  Inputs are regenerated from seeds; only the reference OUTPUTS are
stored in tests/golden/ref_golden.npz."""
import numpy as np
import torch

from oracle import asr_oracle as O

# durations in samples; kept short so the reference/oracle finish in seconds on CPU
CASES = {
    # BASELINE.json configs[0] shape in miniature: greedy, one utterance (all 40 steps run)
    "greedy1": dict(bw=None, wseed=1234, variant="sharp", eos_bias=9.0, seeds=[101], nsamp=[32000]),
    "greedy3_plain": dict(bw=None, wseed=1234, variant="plain", seeds=[111, 112, 113],
                          nsamp=[24000, 16000, 30011]),
    # greedy with EOS at steps 2 / 22 / 0 (empty hypothesis -> score 0.0, model.py:588-590)
    "greedy3": dict(bw=None, wseed=1234, variant="sharp", eos_bias=9.0, seeds=[121, 122, 123],
                    nsamp=[24000, 16000, 30011]),
    # beam bw=4 (configs[1] in miniature), mixed lengths; two utterances finish, two fall back
    "beam4": dict(bw=4, wseed=1234, variant="sharp", eos_bias=8.0, seeds=[201, 202, 203, 204],
                  nsamp=[32000, 18000, 26000, 40000]),
    # early stop: every utterance's rank-0 candidate has been </s> by step 16 (model.py:897-901)
    "beam4es": dict(bw=4, wseed=5, variant="sharp", eos_bias=9.5, seeds=[601, 602],
                    nsamp=[24000, 30000]),
    # bw=16 mixed lengths (configs[2] in miniature) with temperature != 1 and a length bonus
    "beam16": dict(bw=16, wseed=77, variant="sharp", eos_bias=9.0, seeds=[301, 302, 303],
                   nsamp=[20000, 36000, 28000], length_weight=1.5, temperature=2),
    # bw=8 + second-pass LM rescoring (configs[3] in miniature); the LM changes utterance 4's pick
    "beam8lm": dict(bw=8, wseed=1234, variant="sharp", eos_bias=8.0,
                    seeds=[411, 412, 413, 414, 415],
                    nsamp=[32000, 24000, 30000, 22000, 34000], lm=7, lm_weight=0.3,
                    length_weight=2.0),
    # plain init never emits </s>: every utterance takes the un-finished fallback
    # (model.py:961-972).  Logit gaps are ~1e-6 here (SURVEY.md section 7 hard part 1), so this
    # case pins scores / lengths only, not tokens.
    "beam4_plain": dict(bw=4, wseed=1234, variant="plain", seeds=[501, 502],
                        nsamp=[20000, 26000], length_weight=1.5),
}


def case_weights(cs):
    kw = {}
    if "eos_bias" in cs:
        kw["eos_bias"] = cs["eos_bias"]
    return O.make_weights(cs["wseed"], cs["variant"], **kw)


def case_inputs(cs):
    pcms = [O.synth_pcm(s, n) for s, n in zip(cs["seeds"], cs["nsamp"])]
    feats = [O.features(p) for p in pcms]
    lens = torch.tensor([f.size(0) for f in feats])
    return pcms, feats, lens


def load_golden():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")
    return np.load(path, allow_pickle=False)


# ---- rows either side of the hot path: 16-bit ingest, batch_audio, WER ----
FRONTEND = dict(seeds=[801, 802, 803, 804], nsamp=[24000, 16001, 31000, 8000])


def load_frontend_golden():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_frontend.npz")
    return np.load(path, allow_pickle=False)


def wer_token_pairs():
    """Seeded (hypothesis tokens, reference tokens) pairs: identical, edited, empty hypothesis,
    specials ('<unk>' = five characters), the ' ' token, and a long reference."""
    rng = np.random.default_rng(4321)
    pairs = []
    for i in range(10):
        hyp = [int(t) for t in rng.integers(4, 5004, size=int(rng.integers(1, 41)))]
        pairs.append((hyp, O.synth_reference_text(100 + i, hyp)))
    pairs.append(([10, 11, 12], [10, 11, 12]))
    pairs.append(([], [20, 21, 22, 23]))
    pairs.append(([3, 30, 3, 31], [30, 3, 31, 781, 32]))
    pairs.append(([40, 781, 41], [40, 41]))
    pairs.append(([50, 51], [int(t) for t in rng.integers(4, 5004, size=150)]))
    pairs.append(([int(t) for t in rng.integers(4, 5004, size=40)], [60]))
    return pairs


def wer_pairs(int2word):
    return [("".join(int2word[t] for t in h), "".join(int2word[t] for t in r)) for h, r in wer_token_pairs()]
