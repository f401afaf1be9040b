#!/usr/bin/env python
"""NOTE: pass the mangled name of ONE instantiation (k_sweepILb0 = sse::k_sweep<false>); a substring that matches
both instantiations maps addresses to the wrong lines.
Summarise an ncu --set full report: key counters, stall breakdown, samples by CUDA source line.
usage: summarize_ncu.py <report.ncu-rep> <lib.so> <kernel-substring> "<command that produced it>" > profiles/xxx.txt"""
import csv, os, subprocess, sys
rep, so, kname, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("# ncu --set full --clock-control none summary")
print("# command:", cmd)
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.per_cycle_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic']
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h:72s} {u:16s} {v}")
items = []
for h, u, v in zip(hdr, units, vals):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try:
            items.append((float(v.replace(',', '')), h))
        except ValueError:
            pass
tot = sum(x for x, _ in items) or 1
print("# warp stall sampling (share of samples)")
for x, h in sorted(items, reverse=True)[:10]:
    print(f"{100 * x / tot:6.2f}%  {h}")
print("# samples / executed warp-instructions by CUDA source line")
here = os.path.dirname(os.path.abspath(__file__))
print(subprocess.run([sys.executable, os.path.join(here, "ncu_by_line.py"), rep, so, kname, "40"], capture_output=True, text=True).stdout)
