#!/usr/bin/env python
"""Race detector by repetition: the same batch decoded N times must give bit-identical encoder memory, states, tokens
and scores every time (the recurrence's exchange protocol, the GEMM pipelines and the attention rings are all
hand-rolled mbarrier / proxy-fence protocols; compute-sanitizer's racecheck is closed on the pool).
    python tools/stress_repro.py [repetitions] [B ...]"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import asr_oracle as O  # noqa: E402
from chinese_asr_b200.model import Model  # noqa: E402
from chinese_asr_b200.gpd import gpd  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sizes = [int(a) for a in sys.argv[2:]] or [512, 700, 37]
gpd["verbose"] = False
m = Model()
m.load_state(O.make_weights(1234, "sharp", eos_bias=8.0))
bad = 0
for B in sizes:
    rng = np.random.default_rng(B)
    secs = rng.integers(2, 11, size=B) if B != 512 else np.full(B, 10)
    pcms = [O.synth_pcm_int16(9000 + i, int(16000 * s)) for i, s in enumerate(secs)]
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in pcms])
    x = torch.from_numpy(np.concatenate(pcms)).cuda()
    seen = {}
    for bw in (8, 4):
        for r in range(reps):
            tok, ln, sc = m.transcribe(x, off, bw=bw, resident=True)
            hsh = hashlib.sha1(np.ascontiguousarray(tok).tobytes() + np.ascontiguousarray(ln).tobytes()
                               + np.ascontiguousarray(sc).tobytes()).hexdigest()
            seen.setdefault((bw, hsh), 0)
            seen[(bw, hsh)] += 1
    kinds = {bw: [h for (b, h) in seen if b == bw] for bw in (8, 4)}
    ok = all(len(v) == 1 for v in kinds.values())
    bad += 0 if ok else 1
    print(f"B={B} reps={reps}: distinct results per beam width {[len(v) for v in kinds.values()]} {'OK' if ok else 'MISMATCH'}")
m.check_guards()
print("stress_repro", "ok" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)
