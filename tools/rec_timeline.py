#!/usr/bin/env python
"""Per-step phase timeline (clock64, cluster 0) of the recurrence kernel at the bench workload:
    python tools/rec_timeline.py [B] [chunks per direction] 2> profiles/rNN_rec_timeline.txt"""
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
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 7
gpd["verbose"] = False
n = 160000
m = Model()
m.load_state(O.make_weights(1234, "plain"))
m.set_recurrence_chunks(chunks)
rng = np.random.default_rng(1)
pcm = torch.from_numpy((0.1 * rng.standard_normal(B * n)).astype(np.float32)).cuda()
off = np.arange(B + 1, dtype=np.int64) * n
m.transcribe(pcm, off, bw=8, resident=True)
m.stage_timing(True, recurrence_timeline=True)
m.transcribe(pcm, off, bw=8, resident=True)
print({k: round(v, 3) for k, v in m.stage_times().items()})
