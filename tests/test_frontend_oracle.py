"""Oracle restatements of the rows either side of the hot path against
tests/golden/ref_frontend.npz - outputs of the UNMODIFIED reference (tools/make_golden_frontend.py) -
plus the host logic of the batched front end.  CPU only."""
import os
import pickle
import wave

import numpy as np
import torch

from oracle import asr_oracle as O
from tests.cases import CASES, FRONTEND, case_inputs, case_weights, load_frontend_golden, wer_pairs, wer_token_pairs

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _vocab():
    with open(os.path.join(GOLD_DIR, "dict.pkl"), "rb") as f:
        return pickle.load(f)


def test_int16_ingest_and_batch_audio_match_reference():
    g = load_frontend_golden()
    pcm16 = [O.synth_pcm_int16(s, n) for s, n in zip(FRONTEND["seeds"], FRONTEND["nsamp"])]
    raw = [O.features(O.pcm_from_int16(x), normalise=False) for x in pcm16]
    norm, lens = O.batch_audio(raw)                               # data.py:509-518, eps 1e-7
    assert lens.dtype == torch.int32 and lens.tolist() == g["fe_lens"].tolist()
    for i, (r, n) in enumerate(zip(raw, norm)):
        rows = g[f"fe{i}_rows"]
        assert float(np.abs(r.numpy()[rows] - g[f"fe{i}_raw"]).max()) <= 1e-6
        assert float(np.abs(n.numpy()[rows] - g[f"fe{i}_norm"]).max()) <= 1e-5
    t, l2, text = O.collate_fn([(f, [1, 2]) for f in raw])        # data.py:496-507
    assert text == [[1, 2]] * len(raw) and all(torch.equal(a, b) for a, b in zip(t, norm))
    assert O.collate_fn([(f,) for f in raw])[2] is None
    # eps 1e-7 (loader) and 1e-6 (main.py:37) are different results
    assert not torch.equal(O.cmvn(raw[0], 1e-7), O.cmvn(raw[0], 1e-6))


def test_edit_distance_matches_reference_dp():
    g = load_frontend_golden()
    _, i2w = _vocab()
    pairs = wer_pairs(i2w)
    assert [O.edit_distance(p, r) for p, r in pairs] == g["wer_dist"].tolist()
    assert np.allclose([O.get_wer(p, r) for p, r in pairs], g["wer_norm"], rtol=0, atol=1e-12)
    assert np.allclose([O.get_wer(p, r) for p, r in pairs], g["wer_get_wer"], rtol=0, atol=1e-12)
    # '<unk>' counts five characters, the ' ' token one
    assert O.get_wer(i2w[3], "x", normalize=False) == 5 and len(i2w[781]) == 1


def test_driver_wer_matches_reference():
    g = load_frontend_golden()
    _, i2w = _vocab()
    for cname in ("greedy3", "beam4"):
        cs = CASES[cname]
        w = case_weights(cs)
        _, feats, lens = case_inputs(cs)
        if cs["bw"] is None:
            hyp = O.greedy_decode(w, feats, lens, i2w)["tokens"]
        else:
            hyp = O.beam_decode(w, cs["bw"], feats, lens, i2w)["tokens"]
        refs = [O.synth_reference_text(7000 + i, h) for i, h in enumerate(hyp)]
        mean, _ = O.batch_wer(hyp, refs, i2w)
        assert abs(mean - float(g[cname + "_wer"])) < 1e-9
        assert ["".join(i2w[t] for t in r) for r in refs] == g[cname + "_wer_text"].tolist()


def test_read_pcm_and_dataset_items(tmp_path):
    from chinese_asr_b200 import data
    w2i, i2w = _vocab()
    x = O.synth_pcm_int16(5, 4000)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(x.tobytes())
    y = data.read_pcm(p)
    assert y.dtype == np.int16 and np.array_equal(y, x)
    assert np.array_equal(data.fast_read(p), O.pcm_from_int16(x))        # data.py:109-121

    class AB:
        word2int, int2word = w2i, i2w
    chars = [i2w[10], i2w[11], "not-in-vocab"]
    dst = data.AudioDst(AB, "eval", "dev", path_list=[p, x], text_list=[chars, chars[:1]])
    assert len(dst) == 2
    pcm, text = dst[0]
    assert pcm.dtype == np.int16 and text == [10, 11, w2i["<unk>"]]      # data.py:455
    assert dst[1][0].dtype == np.int16
    infer = data.AudioDst(AB, "infer", path_list=[p])
    assert len(infer[0]) == 1
    ld = data.AudioLoader(dst, batch_size=1)
    assert len(ld) == 2 and ld.loader is ld
    import pytest
    with pytest.raises(NotImplementedError):                 # the reference's default mode is 'train' (data.py:393)
        data.AudioDst(AB)
    with pytest.raises(ValueError):
        data.AudioDst(AB, "infer")                           # no manifests here: paths must be given


def test_wer_pairs_cover_edge_cases():
    pairs = wer_token_pairs()
    assert any(len(h) == 0 for h, _ in pairs) and any(h == r for h, r in pairs)
    assert any(3 in h for h, _ in pairs) and any(781 in r for _, r in pairs)
    assert max(len(r) for _, r in pairs) >= 150


def test_edit_distance_properties():
    """Metric properties of the oracle's edit distance (the checker of asr_wer): identity, symmetry, the
    length bounds and the triangle inequality, on seeded random strings over a small alphabet."""
    rng = np.random.default_rng(99)

    def rand_s():
        return "".join(chr(97 + int(c)) for c in rng.integers(0, 4, size=int(rng.integers(0, 12))))

    for _ in range(200):
        a, b, c = rand_s(), rand_s(), rand_s()
        dab, dba = O.edit_distance(a, b), O.edit_distance(b, a)
        assert dab == dba and (dab == 0) == (a == b)
        assert abs(len(a) - len(b)) <= dab <= max(len(a), len(b))
        assert O.edit_distance(a, c) <= dab + O.edit_distance(b, c)
        assert O.edit_distance(a + "x", b + "x") == dab            # a common suffix changes nothing


def test_batch_audio_statistics():
    """batch_audio's instance normalisation: zero mean, unbiased unit std per column (up to eps), per utterance."""
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(n, 720, generator=g) * 3 + 1 for n in (7, 40)]
    out, lens = O.batch_audio(feats)
    assert lens.tolist() == [7, 40]
    for y in out:
        assert float(y.mean(dim=0).abs().max()) < 1e-5
        assert float((y.std(dim=0) - 1).abs().max()) < 1e-5


def test_convert_audio_oracle_properties():
    """The builder-defined resampler (main.py:19-24 runs ffmpeg + sox, which are not in the reference tree): a tone
    keeps its frequency and lands at -1 dBFS, content above the new Nyquist is rejected, 16 kHz input passes
    through, the output length is floor(n * 16000 / rate)."""
    sr = 48000
    t = np.arange(sr // 2) / sr
    tone = (0.3 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    y = O.convert_audio(np.stack([tone, tone], 1), sr)
    assert y.dtype == np.int16 and len(y) == len(tone) * 16000 // sr
    assert int(np.abs(y).max()) == int(np.rint(10 ** (-1 / 20) * 32768))
    sp = np.abs(np.fft.rfft(y.astype(np.float64)))
    assert abs(np.argmax(sp) * 16000 / len(y) - 1000) < 4
    mix = (0.3 * np.sin(2 * np.pi * 10000 * t) + 0.3 * np.sin(2 * np.pi * 500 * t)).astype(np.float32)
    y2 = O.convert_audio(mix, sr)
    sp2 = np.abs(np.fft.rfft(y2.astype(np.float64)))
    f = np.arange(len(sp2)) * 16000 / len(y2)
    assert sp2[(f > 5500) & (f < 6500)].max() < 1e-3 * sp2.max()          # the 10 kHz tone would alias to 6 kHz
    x16 = (tone * 32767).astype(np.int16)
    y3 = O.convert_audio(x16, 16000)
    assert len(y3) == len(x16)
    scale = 10 ** (-1 / 20) * 32768 / np.abs(x16).max()
    assert np.array_equal(y3, np.clip(np.rint(x16 * scale), -32768, 32767).astype(np.int16))
    assert np.all(O.convert_audio(np.zeros(4000, np.float32), 8000) == 0)  # silence stays silence
