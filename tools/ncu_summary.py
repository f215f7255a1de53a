#!/usr/bin/env python
"""Condenses ncu output into the small CSV/markdown files kept under profiles/.

  tools/ncu_summary.py full  <report.ncu-rep> <out.csv>     selected --set full metrics, one row per captured launch
  tools/ncu_summary.py list  <launches.csv>   <out.md>      per-kernel totals and share of the launch list
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    for r in data:
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:30], "ms", d.get("gpu__time_duration.sum"), "dramR", d.get("dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")],
              "dramW", d.get("dram__bytes_write.sum"))


def launch_list(src, out):
    tot = OrderedDict()
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r["Metric Value"]) / 1e6
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(f"ncu launch list ({src}): gpu__time_duration.sum per kernel, --clock-control none; times are serialised and cold-cache\n\n")
        f.write("| kernel | launches | total ms | share | share without k_synth |\n|---|---|---|---|---|\n")
        nosynth = total - tot.get("k_synth", [0, 0.0])[1]
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {ms:.3f} | {ms / total:.3f} | {'' if k == 'k_synth' else f'{ms / nosynth:.3f}'} |\n")
    print(open(out).read())


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
