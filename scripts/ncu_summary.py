#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_launches.csv and <tag>_prof_*.ncu-rep into profiles/<tag>_*.txt
(the tracked evidence; gpurun_out/ is scratch).  usage: scripts/ncu_summary.py <tag>"""
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

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]

launches = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.isfile(launches):
    rows = list(csv.reader(open(launches)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: gpurun_out/{tag}_launches.csv, {sum(len(v) for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v):10.1f} us  total={sum(v)/1e3:9.3f} ms  share={sum(v)/tot:.4f}\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.txt")).read())

for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_prof_*.ncu-rep"))):
    name = os.path.basename(rep)[:-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    with open(os.path.join(out_dir, f"{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source gpurun_out/{name}.ncu-rep\n")
        for r in rows[2:]:
            f.write(f"--- {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"  {w:80s} {r[i]:>16s} {units[i]}\n")
    print(open(os.path.join(out_dir, f"{name}.txt")).read())
