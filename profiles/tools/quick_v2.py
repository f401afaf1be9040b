"""Quick throughput probe of the persistent sweep kernel (no torch): 2D Heisenberg L x L at beta, W walkers, thermalised by
beta doubling, then timed sse_advance launches of a fixed visit budget per walker.
usage: quick_v2.py L beta W doublings per_level therm_sweeps budget_visits n_launches [worm_warps stream_warps] [out.jsonl]
       env SSE_PROBE_SHAPES="ww,sw,level;ww,sw,level;..." times every listed launch shape on the same thermalised batch
       env SSE_PROBE_LIB=path/to/libsse_b200_wNN.so uses a tuning build of the library
       env SSE_PROBE_DESYNC=visits: budget of the untimed launch that takes the walkers out of step (default: budget_visits)"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import os

import sse_b200 as S
from sse_b200 import capi
from sse_b200.walkers import DeviceModel, Walkers

if os.environ.get("SSE_PROBE_LIB"):  # tuning builds of the library (csrc/Makefile `variants`)
    capi.LIB_PATH = os.path.abspath(os.environ["SSE_PROBE_LIB"])

L, beta, W, dbl, per_level, therm = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
budget, launches = int(float(sys.argv[7])), int(sys.argv[8])
ww, sw = (int(sys.argv[9]), int(sys.argv[10])) if len(sys.argv) > 10 else (0, 0)
out = sys.argv[11] if len(sys.argv) > 11 else None
model = S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(L, L)), J=1.0, measure=["magnetization", "staggered_magnetization"]))
dm = DeviceModel(model=model)
n_est = 0.71 * beta * 2 * L * L
t0 = time.time()
wk = Walkers(dm, np.full(W, 1.0 / beta), m_capacity=int(4 * n_est) + 16384, n_capacity=int(1.08 * n_est) + 4096, seed=7)
wk.set_launch_shape(ww, sw)
print(f"L={L} beta={beta} W={W}: {wk.device_bytes() / 1e9:.1f} GB, {wk.device_bytes() / W / 1e6:.2f} MB per walker", flush=True)
wk.thermalize_by_beta_doubling(dbl, sweeps_per_level=per_level)
t1 = time.time()
wk.sweep(therm, thermalized=False)
t2 = time.time()
c = wk.fetch_counters(reset=True)
print(f"setup: doubling {t1 - t0:.1f} s, {therm} sweeps at target {t2 - t1:.1f} s ({c['visits'] / max(1e-9, t2 - t0):.3e} visits/s overall), "
      f"mean n {c['sum_n'] / max(1, c['sweeps']):.0f}", flush=True)
wk.advance(int(float(os.environ.get("SSE_PROBE_DESYNC", budget))), thermalized=True)  # de-synchronise the walkers
wk.fetch_counters(reset=True)
shapes = [(ww, sw, None)]
if os.environ.get("SSE_PROBE_SHAPES"):
    shapes = [tuple(int(x) for x in t.split(",")) for t in os.environ["SSE_PROBE_SHAPES"].split(";")]
for (sww, ssw, level) in shapes:
    if level is not None:
        os.environ["SSE_B200_SMEM_LEVEL"] = str(level)
    wk.set_launch_shape(sww, ssw)
    for i in range(launches):
        t = time.time()
        wk.advance(budget, thermalized=True, measure=(i % 2 == 1))
        dt = time.time() - t
        c = wk.fetch_counters(reset=True)
        clk = 1.965e9
        line = dict(L=L, beta=beta, walkers=W, budget=budget, measure=bool(i % 2), seconds=dt, visits_per_s=c["visits"] / dt,
                    sweeps=c["sweeps"], mean_n=c["sum_n"] / max(1, c["sweeps"]), mean_M=c["sum_M"] / max(1, c["sweeps"]),
                    visits_per_sweep=c["visits"] / max(1, c["sweeps"]),
                    lane_occupancy=c["lane_iters"] / max(1, 32 * c["warp_iters"]),
                    worm_iter_cycles=c["cycles_worm"] / max(1, c["warp_iters"]),
                    stream_busy_warps_per_cta=(c["cycles_build"] + c["cycles_finish"]) / (dt * clk) / min(W, 148),
                    build_cycles_per_task=c["cycles_build"] / max(1, c["tasks"]), finish_cycles_per_task=c["cycles_finish"] / max(1, c["tasks"]),
                    shape=(sww, ssw, level))
        print(json.dumps(line), flush=True)
        if out:
            open(out, "a").write(json.dumps(line) + "\n")
wk.finish_sweeps(thermalized=True)
print("flags:", int((wk.get_flags() & 7).sum()), "energy check n*T/N:", float(wk.num_operators().mean() / beta / (L * L)))
