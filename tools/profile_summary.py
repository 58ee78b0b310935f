#!/usr/bin/env python
"""Summarise an ncu report (brought back in gpurun_out/) into a small markdown table for profiles/.
usage: tools/profile_summary.py <report.ncu-rep> <out.md> [title]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu data wavefronts %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
ki = hdr.index("Kernel Name")
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none` (cold-cache, serialised replays: compare shares, not absolutes).\n\n")
    f.write("| kernel | " + " | ".join(n for _, n in idx) + " |\n|---|" + "---|" * len(idx) + "\n")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("orbk::", "").replace("<unnamed>::", "")
        vals = []
        for i, _ in idx:
            v = r[i].replace(",", "")
            try:
                x = float(v)
                vals.append(f"{x:.1f}" if abs(x) < 1e6 and x != int(x) else f"{int(x)}")
            except ValueError:
                vals.append(v)
        f.write(f"| {name} | " + " | ".join(vals) + " |\n")
print(open(out).read())
