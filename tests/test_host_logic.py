"""Host-side logic of the product (CPU only): constant tables, WAV ingest, ARPA loader, sharding
and the N>1 gather over gloo."""
import os
import wave

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.cases import load_golden
from oracle import asr_oracle as O


def test_feature_constants_match_reference():
    from chinese_asr_b200 import data
    g = load_golden()
    fc = data.feature_consts()
    assert np.array_equal(fc["mel_fb"], g["fb"])           # data.py:21-57, bit-exact
    assert np.array_equal(fc["window"], g["window"])       # data.py:381-382
    assert np.array_equal(fc["taps"], O.delta_taps())
    assert int((fc["mel_fb"] != 0).sum()) == 504


def test_fast_read_wav(tmp_path):
    from chinese_asr_b200 import data
    x = (np.random.default_rng(0).standard_normal(4000) * 3000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(x.tobytes())
    y = data.fast_read(p)
    assert y.dtype == np.float32 and np.array_equal(y, x.astype(np.float32) / 32768.0)


def test_arpa_loader_matches_tables(tmp_path):
    from chinese_asr_b200.lm import NGramLM
    import pickle
    w2i, i2w = pickle.load(open(os.path.join(os.path.dirname(__file__), "golden", "dict.pkl"), "rb"))
    words = [i2w[i] for i in (4, 5, 6, 7, 8)]
    arpa = tmp_path / "t.arpa"
    lines = ["\\data\\", "ngram 1=7", "ngram 2=2", "ngram 3=1", "", "\\1-grams:"]
    for i, wd in enumerate(["<s>", "</s>", "<unk>"] + words[:4]):
        lines.append(f"{-1.0 - 0.1 * i}\t{wd}\t{-0.2 - 0.01 * i}")
    lines += ["", "\\2-grams:", f"-0.5\t{words[0]} {words[1]}\t-0.3", f"-0.7\t<s> {words[0]}\t-0.1", "",
              "\\3-grams:", f"-0.25\t<s> {words[0]} {words[1]}", "", "\\end\\"]
    arpa.write_text("\n".join(lines), encoding="utf-8")
    lm = NGramLM.from_arpa(str(arpa), w2i)
    t = lm.tables()
    assert t["uni_logp"][w2i[words[0]]] == np.float32(-1.3)
    V = len(w2i)
    key = w2i[words[0]] * V + w2i[words[1]]
    assert key in t["bi_keys"].tolist()
    assert ((w2i["<s>"] * V + w2i[words[0]]) * V + w2i[words[1]]) in t["tri_keys"].tolist()


def test_shard_utterances_balanced():
    from chinese_asr_b200.parallel import shard_utterances
    rng = np.random.default_rng(1)
    n = rng.integers(32000, 320000, 257)
    for world in (1, 2, 4, 8):
        sh = shard_utterances(n, world)
        assert sorted(i for s in sh for i in s) == list(range(257))
        counts = [len(s) for s in sh]
        assert max(counts) - min(counts) <= 1
        tot = [int(n[s].sum()) for s in sh]
        assert max(tot) / min(tot) < 1.05


def _gather_worker(rank, world, port, total, max_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chinese_asr_b200.parallel import shard_utterances, pack_records, gather_hypotheses
    n = np.arange(total) * 7 % 13 + 5
    mine = shard_utterances(n, world)[rank]
    toks = np.stack([np.full(max_len, i, dtype=np.int32) for i in mine]) if mine else np.zeros((0, max_len), np.int32)
    rec = pack_records(mine, toks, [i % max_len for i in mine], [-(i + 0.5) for i in mine], max_len)
    t, l, s = gather_hypotheses(rec, total, max_len)
    ok = all((t[i] == i).all() and l[i] == i % max_len and s[i] == np.float32(-(i + 0.5)) for i in range(total))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gather_hypotheses_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, 11, 40, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res) and len(res) == world


def test_batch_pipeline_round_robin_and_order():
    """BatchPipeline hands batches to its engines in round-robin order, one worker thread per engine, and the
    futures keep submission order (host-only fake engines: no device involved)."""
    import threading
    import time
    from chinese_asr_b200.parallel import BatchPipeline

    class Fake:
        device = None

        def __init__(self, name, delay):
            self.name, self.delay, self.threads, self.seen = name, delay, set(), []

        def transcribe(self, pcm, offsets, **kw):
            self.threads.add(threading.get_ident())
            self.seen.append(pcm)
            time.sleep(self.delay)
            return (self.name, pcm, kw.get("bw"))

    a, b = Fake("a", 0.03), Fake("b", 0.0)
    pipe = BatchPipeline([a, b])
    res = pipe.map([(i, None) for i in range(6)], bw=8)
    assert res == [("a", 0, 8), ("b", 1, 8), ("a", 2, 8), ("b", 3, 8), ("a", 4, 8), ("b", 5, 8)]
    assert a.seen == [0, 2, 4] and b.seen == [1, 3, 5]
    assert len(a.threads) == 1 and len(b.threads) == 1 and a.threads != b.threads
    assert pipe.each(lambda m: m.name) == ["a", "b"]
    pipe.close()
    with pytest.raises(ValueError):
        BatchPipeline([])


def test_arpa_vocabulary_mismatch_tables(tmp_path):
    """an ARPA word missing from dict.pkl must not be folded onto <unk> (it would overwrite <unk>'s
    own unigram and duplicate n-gram keys); a dict.pkl token missing from the ARPA maps to <unk> (id_map)."""
    from chinese_asr_b200.lm import NGramLM, _hash_slot
    import pickle
    w2i, i2w = pickle.load(open(os.path.join(os.path.dirname(__file__), "golden", "dict.pkl"), "rb"))
    a, b = i2w[10], i2w[11]
    arpa = tmp_path / "oov.arpa"
    arpa.write_text("\\data\\\nngram 1=5\nngram 2=3\n\n\\1-grams:\n"
                    f"-2.5\t<unk>\t-0.3\n-99\t<s>\t-0.4\n-1.1\t</s>\n-1.3\t{a}\t-0.2\n-0.7\tZZZ\t-0.9\n\n"
                    f"\\2-grams:\n-0.5\t<s> {a}\t-0.1\n-0.1\t{a} ZZZ\t-0.6\n-0.6\t{a} <unk>\t-0.05\n\n\\end\\\n",
                    encoding="utf-8")
    lm = NGramLM.from_arpa(str(arpa), w2i)
    t = lm.tables()
    assert t["uni_logp"][3] == np.float32(-2.5) and t["uni_bo"][3] == np.float32(-0.3)      # <unk> kept
    assert t["id_map"][10] == 10 and t["id_map"][11] == 3 and t["id_map"][3] == 3 and t["id_map"][1] == 1
    V = len(w2i)
    assert int((t["bi_keys"] >= 0).sum()) == 2                                               # "a ZZZ" dropped
    s = _hash_slot(10 * V + 3, len(t["bi_keys"]))
    while t["bi_keys"][s] != 10 * V + 3:
        s = (s + 1) & (len(t["bi_keys"]) - 1)
    assert t["bi_vals"][s][0] == np.float32(-0.6)                                            # the literal "a <unk>" entry


def test_convert_audio_has_no_cpu_fallback_and_checks_arguments():
    """asr_convert_audio (main.py:19-24 on the device) rejects bad arguments with a message, and on a box without a
    B200 it fails loudly instead of resampling on the host: the product has no CPU path."""
    import ctypes as C
    import torch
    from chinese_asr_b200 import _cabi, data as D
    n_out = C.c_int64(0)
    out = np.zeros(16, dtype=np.int16)
    x = np.zeros(32, dtype=np.int16)
    rc = _cabi.lib.asr_convert_audio(x.ctypes.data_as(C.c_void_p), _cabi.PCM_S16, 32, 0, 16000, -1.0,
                                     out.ctypes.data_as(C.POINTER(C.c_int16)), 16, C.byref(n_out), None)
    assert rc == -1 and b"asr_convert_audio" in _cabi.lib.asr_last_error()          # channels = 0
    rc = _cabi.lib.asr_convert_audio(x.ctypes.data_as(C.c_void_p), _cabi.PCM_S16, 32, 1, 32000, -1.0,
                                     out.ctypes.data_as(C.POINTER(C.c_int16)), 8, C.byref(n_out), None)
    assert rc == -4 and n_out.value == 16                                            # output buffer too small
    assert _cabi.lib.asr_convert_audio_length(48000, 48000) == 16000
    assert _cabi.lib.asr_convert_audio_length(44100, 44100) == 16000
    assert _cabi.lib.asr_convert_audio_length(8000, 8000) == 16000
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.AsrError):
            D.convert_pcm(np.zeros(16000, dtype=np.int16), 48000)
    with pytest.raises(TypeError):
        D.convert_pcm(np.zeros(100, dtype=np.int32), 16000)
