#!/usr/bin/env python
"""Excerpt of the SASS of sse::k_sweep<false>: the worm-visit loop (the loop around the record load LDG.E.128.STRONG.GPU),
with the CUDA source line of every instruction, the ptxas resource line, and a count of local-memory instructions in it.
usage: sass_worm_loop.py <libsse_b200.so> > profiles/r2_worm_loop_sass.txt"""
import os
import re
import subprocess
import sys
import tempfile

so = os.path.abspath(sys.argv[1])
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
fn, cur, items = None, None, []
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        fn = ln.strip().split(".text.")[1].split()[0]
        continue
    if fn is None or "k_sweepILb0" not in fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        continue
    m = re.match(r"^(\.L_x_\d+):", ln.strip())
    if m:
        items.append(("L", m.group(1), None))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        items.append(("I", int(m.group(1), 16), m.group(2).strip(), cur))
lab, ins = {}, []
for it in items:
    if it[0] == "L":
        lab[it[1]] = len(ins)
    else:
        ins.append(it)
best = None
for i, it in enumerate(ins):
    m = re.search(r"BRA\b.*`\((\.L_x_\d+)\)", it[2])
    if m and m.group(1) in lab and lab[m.group(1)] <= i:
        s = lab[m.group(1)]
        body = ins[s:i + 1]
        if (any("LDG.E.128.STRONG" in b[2] for b in body) and sum((b[3] or "").startswith("sse_worm.cuh") for b in body) > 20
                and (best is None or len(body) < len(best))):
            best = body
print("# sse::k_sweep<false>, sm_100a: smallest loop around the worm's record load (lane_visit, csrc/sse_worm.cuh)")
for ln in res.splitlines():
    if "k_sweepILb0" in ln or (ln.strip().startswith("REG") and "k_sweepILb0" in prev):
        print("#", ln.strip())
    prev = ln
loc = [b for b in best if re.search(r"\b(LDL|STL)", b[2])]
print(f"# {len(best)} instructions, {len(loc)} local-memory instructions, "
      f"{sum('LDG' in b[2] for b in best)} LDG, {sum('STG' in b[2] or re.match(r'(@!?P\\d+ +)?ST\\.E', b[2]) is not None for b in best)} ST, "
      f"{sum('LDS' in b[2] for b in best)} LDS")
def in_visit(b):
    m = re.match(r"sse_worm.cuh:(\d+)", b[3] or "")
    return m is not None and 168 <= int(m.group(1)) <= 201
idx = [i for i, b in enumerate(best) if in_visit(b)]
print(f"# below: lane_visit (sse_worm.cuh:168-201) = instructions {idx[0]}..{idx[-1]} of the loop; the rest of the loop is the per-lane\n"
      "# slow paths (worm start / end, sweep end, walker hand-over) and the budget bookkeeping")
for b in best[idx[0]:idx[-1] + 1]:
    print(f"/*{b[1]:05x}*/ {b[2]:70s} // {b[3]}")
