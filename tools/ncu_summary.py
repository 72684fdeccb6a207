#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  tools/ncu_summary.py launches <launches.csv> <out.md>     per-kernel totals of a `--metrics gpu__time_duration.sum` pass
  tools/ncu_summary.py full <prof.ncu-rep> <out.md>         key counters of a `--set full` capture (reads it with `ncu -i`)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[i_unit], 1e-6)
        t = tot.setdefault(r[i_name].split("(")[0], [0, 0.0])
        t[0] += 1
        t[1] += v
    s = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{path}` (cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"{sum(v[0] for v in tot.values())} launches, {s:.3f} ms total\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for n, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{n[:90]}` | {c} | {ms:.3f} | {ms / s:.4f} |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{path}`\n")
        for r in rows[2:]:
            f.write(f"\n## `{r[idx['Kernel Name']][:110]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "-nan", "nan"):
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
