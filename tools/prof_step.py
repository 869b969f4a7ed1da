#!/usr/bin/env python
"""One warm-up + N profiled passes of the hot path at the bench workload (for ncu): 16-bit PCM resident in HBM,
512 x 10 s utterances, bw=8 - the kernels bench.py times.
    python tools/prof_step.py [B] [bw] [passes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import asr_oracle as O  # noqa: E402
from chinese_asr_b200.model import Model  # noqa: E402
from chinese_asr_b200.gpd import gpd  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gpd["verbose"] = False
n = 160000
m = Model()
m.load_state(O.make_weights(1234, "plain"))
rng = np.random.default_rng(1)
x = np.clip(np.round(0.1 * rng.standard_normal(B * n) * 32768.0), -32768, 32767).astype(np.int16)
pcm = torch.from_numpy(x).cuda()
off = np.arange(B + 1, dtype=np.int64) * n
for _ in range(1 + passes):
    tok, ln, sc = m.transcribe(pcm, off, bw=bw, resident=True)
torch.cuda.synchronize()
print("ok", int(ln.sum()), m.launch_count())
