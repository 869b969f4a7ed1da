#!/usr/bin/env python
"""gpurun_out/pass_<tag>.csv (ncu --metrics time, DRAM bytes, pipe activity of EVERY launch of one pass of the bench
workload, tools/prof_cmd.sh) -> profiles/<tag>_pass_metrics.csv (one row per launch), profiles/<tag>_pass_summary.md
(per kernel) and profiles/<tag>_roofline.json (what bench.py's `roofline.traffic` reads).
    python tools/summarize_pass.py r02"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = [l for l in open(os.path.join(ROOT, "gpurun_out", f"pass_{tag}.csv")) if not l.startswith("==")]
ids = collections.OrderedDict()
for r in csv.DictReader(rows):
    d = ids.setdefault(r["ID"], {"kernel": r["Kernel Name"].split("(")[0].replace("void ", ""), "grid": r["Grid Size"],
                                 "block": r["Block Size"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u, m = r["Metric Unit"], r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        m = "time_us"
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        m = "dram_read_bytes" if "read" in m else "dram_write_bytes"
    else:
        m = {"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
             "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
             "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct"}.get(m, m)
    d[m] = v
out = os.path.join(ROOT, "profiles")
cols = ["launch", "kernel", "grid", "block", "time_us", "dram_read_bytes", "dram_write_bytes", "tensor_pipe_pct",
        "issue_active_pct", "xu_pipe_pct"]
with open(os.path.join(out, f"{tag}_pass_metrics.csv"), "w", newline="") as f:
    f.write(f"# one pass of the bench workload (tools/prof_step.py: 512 x 10 s, bw=8, 16-bit PCM resident), every launch;\n"
            f"# ncu --metrics ... --clock-control none (cold-cache, serialised: shares, not absolutes)\n")
    w = csv.DictWriter(f, fieldnames=cols, extrasaction="ignore")
    w.writeheader()
    for i, d in enumerate(ids.values()):
        w.writerow({"launch": i, **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()}})
agg = collections.OrderedDict()
for d in ids.values():
    a = agg.setdefault(d["kernel"], dict(n=0, t=0.0, rd=0.0, wr=0.0, tp=0.0, xu=0.0, iss=0.0))
    t = d.get("time_us", 0.0)
    a["n"] += 1
    a["t"] += t
    a["rd"] += d.get("dram_read_bytes", 0.0)
    a["wr"] += d.get("dram_write_bytes", 0.0)
    a["tp"] += t * d.get("tensor_pipe_pct", 0.0)
    a["xu"] += t * d.get("xu_pipe_pct", 0.0)
    a["iss"] += t * d.get("issue_active_pct", 0.0)
tot = sum(a["t"] for a in agg.values())
with open(os.path.join(out, f"{tag}_pass_summary.md"), "w") as f:
    f.write(f"# one pass of the bench workload under ncu ({len(ids)} launches, {tot / 1e3:.2f} ms of kernel time), per kernel;\n"
            "# DRAM = dram__bytes_read.sum + dram__bytes_write.sum; pipe figures are time-weighted means\n\n")
    f.write("| kernel | launches | ms | share | DRAM GB | GB/s | tensor pipe % | XU pipe % | issue active % |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        t = max(a["t"], 1e-9)
        f.write(f"| `{k}` | {a['n']} | {a['t'] / 1e3:.3f} | {100 * a['t'] / tot:.1f}% | {(a['rd'] + a['wr']) / 1e9:.3f} | "
                f"{(a['rd'] + a['wr']) / t / 1e3:.0f} | {a['tp'] / t:.1f} | {a['xu'] / t:.1f} | {a['iss'] / t:.1f} |\n")
gemm = {k: a for k, a in agg.items() if "gemm_split_pair_kernel" in k}
roof = {"source": f"profiles/{tag}_pass_metrics.csv", "workload": "512 x 10 s, bw=8, one pass",
        "gemm_launches_per_pass": sum(a["n"] for a in gemm.values()),
        "gemm_dram_bytes_per_pass": sum(a["rd"] + a["wr"] for a in gemm.values()),
        "gemm_ms_per_pass_under_ncu": sum(a["t"] for a in gemm.values()) / 1e3,
        "per_kernel": {k: {"launches": a["n"], "ms": a["t"] / 1e3, "dram_bytes": a["rd"] + a["wr"]} for k, a in agg.items()}}
json.dump(roof, open(os.path.join(out, f"{tag}_roofline.json"), "w"), indent=1)
print("wrote", f"{tag}_pass_metrics.csv, {tag}_pass_summary.md, {tag}_roofline.json")
