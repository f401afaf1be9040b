#!/usr/bin/env python
"""Short device timing of the launch shapes (1 / 2 / 4 walkers per warp) on one thermalised batch, without torch:
host clock around sse_sweep + sse_sync, work from the device counters.  Written for the last 100 GPU-seconds of
round 1; every result line is flushed to gpurun_out/ as soon as it exists.
usage: quick_multi.py L beta walkers doublings sweeps_per_level therm timed_sweeps shapes(e.g. 2,4,1) [out.jsonl] [variants(e.g. 0,2,4)]
Every (shape, variant) pair is timed on the same batch (variants = sse_dbg_set_variant bits: 2 no prefetch, 4 no hint pass)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import sse_b200 as S  # noqa: E402
from sse_b200.walkers import DeviceModel, Walkers  # noqa: E402


def main():
    L, beta, W, doublings, per_level, therm, timed = [int(x) for x in sys.argv[1:8]]
    shapes = [int(x) for x in sys.argv[8].split(",")]
    out = sys.argv[9] if len(sys.argv) > 9 else os.path.join(ROOT, "gpurun_out", "quick_multi.jsonl")
    variants = [int(x) for x in sys.argv[10].split(",")] if len(sys.argv) > 10 else [0]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    t_start = time.time()
    model = S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(L, L)), J=1.0, measure=["magnetization"]))
    dm = DeviceModel(model=model)
    n_est = 0.75 * beta * 2 * L * L
    wk = Walkers(dm, np.full(W, 1.0 / beta), m_capacity=int(3.6 * n_est), n_capacity=int(1.7 * n_est), seed=7)
    wk.set_walkers_per_warp(shapes[0])
    wk.thermalize_by_beta_doubling(doublings, sweeps_per_level=per_level, final_sweeps=therm)
    setup = time.time() - t_start
    for k, variant in [(k, v) for k in shapes for v in variants]:
        wk.set_walkers_per_warp(k)
        dm.dbg_set_variant(variant)
        wk.sweep(1, thermalized=True)  # warm-up launch of this shape
        wk.fetch_counters(reset=True)
        t0 = time.perf_counter()
        wk.sweep(timed, thermalized=True)
        dt = time.perf_counter() - t0
        c = wk.fetch_counters(reset=True)
        cyc = c["cycles_diag_build"] + c["cycles_worm"] + c["cycles_commit_measure"]
        line = dict(L=L, beta=beta, walkers=W, walkers_per_warp=k, variant=variant, timed_sweeps=timed, seconds=dt,
                    visits_per_s=c["visits"] / dt, walker_sweeps_per_s=c["sweeps"] / dt, mean_n=c["sum_n"] / c["sweeps"],
                    mean_M=c["sum_M"] / c["sweeps"], visits_per_sweep=c["visits"] / c["sweeps"],
                    worm_cycle_share=c["cycles_worm"] / max(1, cyc), setup_s=setup,
                    note="host clock around sse_sweep+sse_sync, %d doublings x %d sweeps + %d sweeps thermalisation "
                         "(compare shapes and variants, not absolute numbers)" % (doublings, per_level, therm))
        with open(out, "a") as f:
            f.write(json.dumps(line) + "\n")
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
