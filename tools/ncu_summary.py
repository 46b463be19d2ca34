#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  launches <launches.csv> <out.md>            per-kernel count / total / mean of gpu__time_duration
  full <report.ncu-rep> <out.md> [<json>]     key metrics of each profiled launch (+ dram bytes per launch as JSON)
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and\n"
                "serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / total:.1f}% |\n")


def full(rep, out, js=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        kn = hdr.index("Kernel Name")
        for r in data:
            f.write(f"## {r[kn][:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    if js and data:
        def val(r, k):
            i = hdr.index(k)
            v = float(r[i].replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
        r = data[-1]
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench                       # the hash of the kernel's sources: bench.py quotes the capture only for this build
        with open(js, "w") as f:
            json.dump({"dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                       "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                       "kernel": r[hdr.index("Kernel Name")][:100], "source": rep,
                       "kernel_src_sha": bench.headline_kernel_fingerprint()}, f)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
