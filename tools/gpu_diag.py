#!/usr/bin/env python
"""Run every GPU parity check and print / save the metrics (no assertions).  Used on the GPU box:
    python tools/gpu_diag.py [check-name-substring ...]  -> gpurun_out/diag.json"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from tests import gpu_checks as G  # noqa: E402


def main():
    sel = sys.argv[1:]
    checks = [("features", G.check_features, ()),
              ("encoder_kat", G.check_encoder_kat, ()),
              ("encoder_greedy3", G.check_encoder, ("greedy3",)),
              ("encoder_beam4", G.check_encoder, ("beam4",)),
              ("gemm_tc", G.check_gemm, ()),
              ("lm", G.check_lm, ()),
              ("greedy1", G.check_greedy, ("greedy1",)),
              ("greedy3", G.check_greedy, ("greedy3",)),
              ("greedy3_plain", G.check_greedy, ("greedy3_plain",)),
              ("beam4", G.check_beam, ("beam4",)),
              ("beam4es", G.check_beam, ("beam4es",)),
              ("beam16", G.check_beam, ("beam16",)),
              ("beam8lm", G.check_beam, ("beam8lm",)),
              ("beam4_plain", G.check_beam, ("beam4_plain",)),
              ("fused", G.check_fused, ()),
              ("batch_invariance", G.check_batch_invariance, ())]
    out = {}
    print(torch.cuda.get_device_name(0), flush=True)
    for name, fn, args in checks:
        if sel and not any(s in name for s in sel):
            continue
        t0 = time.time()
        try:
            out[name] = fn(*args)
        except Exception as e:  # keep going: first-run diagnostics
            out[name] = {"EXCEPTION": repr(e), "trace": traceback.format_exc()[-1500:]}
        torch.cuda.synchronize()
        print(f"== {name}  ({time.time() - t0:.1f}s)")
        for k, v in out[name].items():
            print(f"   {k}: {v}")
        sys.stdout.flush()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "_".join(sel) if sel else "all"
    with open(os.path.join(ROOT, "gpurun_out", f"diag_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
