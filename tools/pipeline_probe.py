#!/usr/bin/env python
"""Experiment: S handles on one GPU, each decoding full batches of B utterances from its own host thread and
stream, so that one batch's encoder (the latency-bound recurrence occupies 112 SMs at low utilisation) overlaps
another batch's decoder.  Reports batches/s against a single handle.

    python tools/pipeline_probe.py [B] [S] [batches per handle]
"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import asr_oracle as O  # noqa: E402  (weights only)
from chinese_asr_b200.gpd import gpd  # noqa: E402
from chinese_asr_b200.model import Model  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    k, n, max_len = 8, 160000, 40
    gpd["verbose"] = False
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    w = O.make_weights(1234, "plain")
    rng = np.random.default_rng(1000)
    pcm = np.clip(np.round(0.1 * rng.standard_normal((B, n)) * 32768.0), -32768, 32767).astype(np.int16)
    resident = torch.from_numpy(pcm.reshape(-1)).to(dev)
    off = np.arange(B + 1, dtype=np.int64) * n
    L = (1 + (n - 1 - 512) // 160) // 3
    models = []
    for _ in range(S):
        m = Model()
        m.load_state(w)
        m.reserve(B, B * L, k, B * n, max_len)
        models.append(m)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    out = [None] * S

    def worker(i, reps, delay=0.0):
        torch.cuda.set_device(0)
        if delay:
            time.sleep(delay)
        with torch.cuda.stream(streams[i]):
            for _ in range(reps):
                out[i] = models[i].transcribe(resident, off, bw=k, resident=True)

    for i in range(S):
        worker(i, 3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    worker(0, nb)
    torch.cuda.synchronize()
    one = (time.perf_counter() - t0) / nb * 1e3
    print(f"1 handle : {one:.2f} ms per batch of {B}  {B / one * 1e3:.0f} utt/s")
    for delay in (0.0, 0.012):
        ts = [threading.Thread(target=worker, args=(i, nb, delay * i)) for i in range(S)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0 - delay * (S - 1) * 0.5) / (S * nb) * 1e3
        same = all((out[i][0] == out[0][0]).all() for i in range(S))
        print(f"{S} handles (start offset {delay * 1e3:.0f} ms): {dt:.2f} ms per batch  {B / dt * 1e3:.0f} utt/s  identical {same}")


if __name__ == "__main__":
    main()
