"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: summarize_ncu.py launches <launches.csv> <out.md> | full <report.ncu-rep> <out.md>"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        v = float(row[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(row[ui], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\nper-launch times are cold-cache and serialised: compare SHARES\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.2f} | {ms / tot:.3f} |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for d in data:
            get = lambda k: d[hdr.index(k)] if k in hdr else "n/a"
            f.write(f"## `{get('Kernel Name')}`  grid {get('Grid Size')} block {get('Block Size')}\n\n| metric | unit | value |\n|---|---|---:|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {units[hdr.index(k)]} | {d[hdr.index(k)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
