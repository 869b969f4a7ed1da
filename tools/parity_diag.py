#!/usr/bin/env python
"""Print the parity statistics of the full-size oracle comparisons (tests/gpu_checks.py) as JSON lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_checks as G  # noqa: E402

which = sys.argv[1:] or ["config2", "config3", "config4", "bench", "config3_full", "wide600", "wide700", "beam16"]
runs = {
    "config2": lambda: G.check_config_shape(32, 4, [10.0] * 32),
    "config3": lambda: G.check_config_shape(12, 16, [2.0, 20.0, 3.7, 11.3, 7.9, 16.4, 2.6, 13.0, 5.5, 19.2, 9.1, 4.4],
                                            seed0=3300, eos_bias=9.0, wseed=77),
    "config4": lambda: G.check_config_shape(32, 8, [10.0] * 32, lm_seed=7, seed0=3600),
    "bench": G.check_bench_shape,
    "config3_full": G.check_config3_full,
    "wide600": lambda: G.check_wide_recurrence(600),
    "wide700": lambda: G.check_wide_recurrence(700),
    "beam16": lambda: G.check_beam("beam16"),
}
for name in which:
    r = runs[name]()
    print(json.dumps({"check": name, **{k: v for k, v in r.items()}}, default=str), flush=True)
