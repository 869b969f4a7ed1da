#!/usr/bin/env python
"""ms / TFLOP/s of the tensor-core GEMM engine on the hot path's shapes (tuning aid, GPU only).

    python tools/gemm_bench.py [iters]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chinese_asr_b200 import _cabi  # noqa: E402
from chinese_asr_b200._cabi import check, lib  # noqa: E402
from chinese_asr_b200.model import Model  # noqa: E402
from oracle import asr_oracle as O  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
from chinese_asr_b200.gpd import gpd  # noqa: E402
gpd["verbose"] = False
m = Model()
m.load_state(O.make_weights(1234, "sharp", eos_bias=8.0))
SHAPES = [("enc layer0", 169984, 2048, 720, 0), ("enc layer1-3", 169984, 2048, 512, 0), ("keys", 169984, 128, 512, 0),
          ("dec cell", 4096, 2048, 1024, 0), ("query", 4096, 128, 512, 0), ("vocab store", 4096, 5004, 1024, 0),
          ("vocab top2", 4096, 5004, 1024, 2), ("vocab top8", 4096, 5004, 1024, 8), ("vocab top16", 4096, 5004, 1024, 16),
          ("vocab top32", 4096, 5004, 1024, 32), ("vocab16 bw16", 4096, 5004, 1024, 32),
          ("vocab bw4 B32", 128, 5004, 1024, 8), ("enc B32", 10624, 2048, 512, 0)]
for name, M, N, K, slots in SHAPES:
    ms = np.zeros(1, dtype=np.float32)
    check(lib.asr_bench_gemm(m._h, M, N, K, iters, slots, _cabi.fptr(ms), None), "asr_bench_gemm")
    gf = 2.0 * M * N * K / 1e9
    print(f"{name:14s} M={M:6d} N={N:5d} K={K:5d}  {ms[0]*1e3:9.1f} us  {gf/ms[0]:8.1f} TFLOP/s(alg)  "
          f"out {M*N*4/1e6:7.1f} MB", flush=True)
