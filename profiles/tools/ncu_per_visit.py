#!/usr/bin/env python
"""Per-visit figures of sse::k_sweep from an ncu --set full capture of ONE profiled sse_advance launch, for bench.py's
roofline.traffic / roofline.issue_frac (profiles/r2_ksweep_ncu.json).
usage: ncu_per_visit.py <report.ncu-rep> <walkers> <visits_per_walker_in_the_launch> <L> "<command>" > profiles/r2_ksweep_ncu.json
The profiled launch gives every walker the same visit budget, so its visit count is walkers x budget."""
import csv
import json
import subprocess
import sys

rep, walkers, budget, L, cmd = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return v * scale


visits = walkers * budget
out = {
    "L": L, "source": f"ncu --set full capture: {cmd}", "visits_in_launch": visits,
    "dram_bytes_per_visit": (metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")) / visits,
    "dram_bytes_read": metric("dram__bytes_read.sum"), "dram_bytes_write": metric("dram__bytes_write.sum"),
    "warp_instructions_per_visit": metric("smsp__inst_executed.sum") / visits,
    "duration_ms": metric("gpu__time_duration.sum") * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(
        units[hdr.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms"), 1e-6),
    "l2_hit_rate_pct": metric("lts__t_sector_hit_rate.pct"),
    "note": "per-visit figures include the streaming phases of the sweeps the launch passes through (diagonal update, record build, "
            "measurement); algorithmic bytes per visit at this size: 64 (worm) + (12 M + 16 n)/V = 64 + 19",
}
print(json.dumps(out, indent=1))
