#!/usr/bin/env python
"""Experiment: does the DECODER of a batch run faster as two concurrent half-batches (two handles, two streams)?
Encoders are run beforehand and not timed.  Isolates the decoder-on-decoder overlap from tools/substream_probe.py.

    python tools/subdecode_probe.py [B] [steps]
"""
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import asr_oracle as O  # noqa: E402  (weights only)
from chinese_asr_b200.gpd import gpd  # noqa: E402
from chinese_asr_b200.model import Model  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    k, n = 8, 160000
    gpd["verbose"] = False
    torch.cuda.set_device(0)
    w = O.make_weights(1234, "plain")
    rng = np.random.default_rng(1000)
    pcms = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(B)]

    def make(lo, hi):
        m = Model()
        m.load_state(w)
        feats = m.features(pcms[lo:hi], normalise=True)
        lens = torch.tensor([f.size(0) for f in feats])
        m._encode(feats, lens, k)
        return m

    def decode(m, nb):
        return m._beam(nb, k, False, 0.0, 0.0)

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    one = make(0, B)
    ms1 = timeit(lambda: decode(one, B), steps)
    print(f"decode 1 x {B}: {ms1:.2f} ms")
    del one
    for S in (2, 4):
        bs = B // S
        subs = [make(i * bs, (i + 1) * bs) for i in range(S)]
        streams = [torch.cuda.Stream() for _ in range(S)]

        def worker(i):
            torch.cuda.set_device(0)
            with torch.cuda.stream(streams[i]):
                decode(subs[i], bs)

        for i in range(S):
            worker(i)
            worker(i)
        seq = timeit(lambda: [worker(i) for i in range(S)], steps)

        def par():
            ts = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()

        con = timeit(par, steps)
        print(f"decode {S} x {bs}: sequential {seq:.2f} ms, concurrent {con:.2f} ms")
        del subs


if __name__ == "__main__":
    main()
