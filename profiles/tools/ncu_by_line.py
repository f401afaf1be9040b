#!/usr/bin/env python
"""NOTE: pass the mangled name of ONE instantiation (k_sweepILb0 = sse::k_sweep<false>); a substring that matches
both instantiations maps addresses to the wrong lines.
Aggregate an ncu source-page CSV (SASS view) by CUDA source line, using nvdisasm -g line info.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [top]
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
cur = None
infn = False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) <= iexec or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key, _ = line_of.get(a - base, ((None, 0), ""))
    s, e = int(r[isamp] or 0), int(r[iexec] or 0)
    agg[key][0] += s
    agg[key][1] += e
    for i in stall_cols:
        if r[i] and r[i] != "0":
            agg[key][2][hdr[i]] += int(r[i])
    tot_s += s
    tot_e += e
print(f"total samples {tot_s}, total warp-instructions {tot_e}")
srcs = {}
def text(key):
    if not key or not key[0]:
        return ""
    for d in (os.path.dirname(os.path.abspath(so)), os.path.join(os.path.dirname(os.path.abspath(so)), "../../include")):
        p = os.path.join(d, key[0])
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            return srcs[p][key[1] - 1].strip()[:90]
    return ""
print(f"{'samples%':>8} {'inst%':>7}  line  top-stalls / source")
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    stalls = ",".join(f"{k[6:]}:{v * 100 // max(s, 1)}" for k, v in st.most_common(3))
    print(f"{100 * s / tot_s:8.2f} {100 * e / tot_e:7.2f}  {key[0]}:{key[1]}  [{stalls}]  {text(key)}")
