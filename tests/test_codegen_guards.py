"""Static guards on the compiled sm_100a code of sse::k_sweep<false> (no GPU needed: cuobjdump + nvdisasm on the built
library).  They pin the code-generation properties the measured throughput depends on and that a harmless-looking source
change can silently break (round 2: reference parameters of __noinline__ helpers put the accept thresholds of the chunk
loop into local memory, 26 % of that loop's stall samples):
  * the kernel is built with 128 registers (one 512-thread CTA per SM), as csrc/sse_capi.cu's launch bounds intend;
  * no local-memory instruction (LDL/STL) inside the worm-visit path or inside the chunk loops of the diagonal update;
  * the streaming pass prefetches with LDGSTS (cp.async) and the worm visit is one 16-byte load + one 4-byte store."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stochasticseriesexpansion.jl_b200", "csrc", "libsse_b200.so")
KERNEL = "k_sweepILb0"  # sse::k_sweep<false>

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvdisasm") is None,
                                reason="CUDA binary utilities not installed")


@pytest.fixture(scope="module")
def kernel():
    subprocess.check_call(["make", "-C", os.path.dirname(LIB)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    fn, cur, labels, ins = None, None, {}, []
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            fn = ln.strip().split(".text.")[1].split()[0]
            continue
        if fn is None or KERNEL not in fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^(\.L_x_\d+):", ln.strip())
        if m:
            labels[m.group(1)] = len(ins)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    loops = []
    for i, (_, text, _) in enumerate(ins):
        m = re.search(r"BRA\b.*`\((\.L_x_\d+)\)", text)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            loops.append(ins[labels[m.group(1)]:i + 1])
    regs = None
    lines = res.splitlines()
    for j, ln in enumerate(lines):
        if KERNEL in ln and j + 1 < len(lines):
            m = re.search(r"REG:(\d+)", lines[j + 1])
            regs = int(m.group(1)) if m else None
    assert ins, "kernel not found in the library"
    return {"ins": ins, "loops": loops, "regs": regs}


def _local(body):
    return [b for b in body if re.search(r"\b(LDL|STL)\b|\b(LDL|STL)\.", b[1])]


def test_register_budget(kernel):
    assert kernel["regs"] == 128, kernel["regs"]


def test_worm_visit_path_has_no_local_memory(kernel):
    cands = [b for b in kernel["loops"] if any("LDG.E.128.STRONG" in x[1] for x in b)
             and sum(1 for x in b if x[2] and x[2][0] == "sse_worm.cuh") > 20]
    assert cands, "worm loop not found"
    loop = min(cands, key=len)
    visit = [x for x in loop if x[2] and x[2][0] == "sse_worm.cuh" and 168 <= x[2][1] <= 201]
    first, last = loop.index(visit[0]), loop.index(visit[-1])
    path = loop[first:last + 1]  # lane_visit including the inlined helpers between its lines
    assert not _local(path), _local(path)
    assert sum("LDG.E.128" in x[1] for x in path) == 1      # the record
    assert sum(bool(re.search(r"\bSTG\.E\b", x[1])) for x in path) == 1  # the new op code
    assert len(path) < 220, len(path)


def test_chunk_loops_have_no_local_memory_and_prefetch_asynchronously(kernel):
    chunk = [b for b in kernel["loops"] if 1 <= sum("LDGSTS" in x[1] for x in b) <= 3 and len(b) > 300]
    assert len(chunk) >= 3, [len(b) for b in chunk]  # diagonal update with / without measurement, stand-alone measure pass
    for body in chunk:
        assert not _local(body), (len(body), _local(body)[:4])
    sizes = sorted(len(b) for b in chunk)
    assert sizes[0] < 450 and sizes[1] < 760 and sizes[-1] < 1050, sizes  # measured builds: 369 / 647 / 899 instructions
