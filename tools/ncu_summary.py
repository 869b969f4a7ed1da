#!/usr/bin/env python
"""Key metrics of .ncu-rep files (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep [--csv out.csv]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_max_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        yield {h: (v, u) for h, u, v in zip(hdr, units, r)}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_csv = sys.argv[sys.argv.index("--csv") + 1] if "--csv" in sys.argv else None
    if out_csv:
        args.remove(out_csv)
    table = []
    for path in args:
        for r in rows_of(path):
            name = r["Kernel Name"][0]
            rec = {"report": path.split("/")[-1], "kernel": name[:80]}
            for w in WANT:
                if w in r:
                    v, u = r[w]
                    rec[w + (f" [{u}]" if u else "")] = v
            table.append(rec)
            print(rec["report"], rec["kernel"])
            for k, v in rec.items():
                if k not in ("report", "kernel"):
                    print(f"    {k:90s} {v}")
    if out_csv and table:
        keys = []
        for t in table:
            for k in t:
                if k not in keys:
                    keys.append(k)
        with open(out_csv, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            w.writerows(table)


if __name__ == "__main__":
    main()
