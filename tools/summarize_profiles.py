#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ into the small committed summaries under profiles/.

    python tools/summarize_profiles.py r01
      gpurun_out/launches_<tag>.csv       -> profiles/<tag>_launches.md  (per-kernel totals / shares)
      gpurun_out/prof_<tag>_*.ncu-rep     -> profiles/<tag>_kernels.csv  (key metrics per captured launch)
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_max_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def launches():
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("asr::", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, {tag}: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1` (one engine;\n"
                "# 512 x 10 s utterances per pass, bw=8; 7 device-resident passes, 4 end-to-end passes, 1 stage-timed pass;\n"
                "# tools/prof_cmd.sh.  With the default two engines the same kernels run, interleaved from two streams),\n"
                "# `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000` after the same command exited 0\n"
                "# without ncu.  Per-launch times are cold-cache and serialised: compare SHARES with stage_ms.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"| **total** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.2f} | | |\n")
    print("wrote", f"{tag}_launches.md", "total ms", tot / 1e3)


def kernels():
    rows_out = []
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_*.ncu-rep")))
    raws = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_*_raw.csv")))   # exported on the GPU box
    for rep in reps + raws:
        if rep.endswith(".csv"):
            raw = open(rep).read()
        else:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        units = rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        # ncu picks a unit per report: normalise durations to ms and DRAM bytes to GB
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3,
                 "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
        for r in rows[2:]:
            d = {"report": os.path.basename(rep), "kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
            for k in KEEP:
                v = r[idx[k]] if k in idx else ""
                if k in idx and (k.startswith("gpu__time_duration") or k.startswith("dram__bytes")) and v:
                    u = units[idx[k]]
                    if u in scale:
                        v = "%.6f" % (float(v.replace(",", "")) * scale[u])
                d[k] = v
            rows_out.append(d)
    if not rows_out:
        return
    with open(os.path.join(out_dir, f"{tag}_kernels.csv"), "w", newline="") as f:
        f.write("# units: gpu__time_duration.sum ms, dram__bytes_* GB, the rest as named (pct / counts)\n")
        if tag == "r01":
            f.write("# prof_r01_all_raw: one pass captured before feat_write_kernel wrote the split operand itself - its\n"
                    "# feat_write_kernel and 0.23 ms split_operand_kernel<0> rows are that earlier pair; prof_r01_fw_raw is the\n"
                    "# fused feat_write_kernel (6 B per value written, no split pass).  All other kernels are unchanged.\n")
        w = csv.DictWriter(f, fieldnames=["report", "kernel"] + KEEP)
        w.writeheader()
        w.writerows(rows_out)
    print("wrote", f"{tag}_kernels.csv", len(rows_out), "launches")


launches()
kernels()
