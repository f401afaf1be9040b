#!/usr/bin/env python
"""Thermalisation cost on the CPU oracle: the reference's cold start (Carlo.init! + unthermalised sweeps) against
beta doubling as Walkers.thermalize_by_beta_doubling does it (levels at controller attenuation 0.1, avg_wl doubled at
each level).  Prints worm-vertex visits per sweep and the controller state.  The numbers quoted in DESIGN.md §4 come
from `python profiles/tools/therm_study.py 32 32 5`.     usage: therm_study.py L beta doublings"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from helpers import heisenberg_square  # noqa: E402
from oracle import OracleModel, OracleWalker  # noqa: E402
from test_zz_beta_doubling import _oracle_thermalize_by_doubling  # noqa: E402


def follow(w, n_sweeps, marks, tag):
    tot = 0
    for s in range(1, n_sweeps + 1):
        c0 = w.fetch_counters()["visits"]
        w.sweep(1)
        v = w.fetch_counters()["visits"] - c0
        tot += v
        if s in marks:
            st = w.get_state()
            print(f"{tag} sweep {s}: n={st['num_operators']} num_worms={st['num_worms']:.1f} avg_wl={st['avg_worm_length']:.0f} "
                  f"visits this sweep={v} cumulative={tot}", flush=True)


def main():
    L, beta, k = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
    om = OracleModel(model=heisenberg_square(L, False, measure=("magnetization",)))
    ow = OracleWalker(om, 1 / beta, seed=5, walker_id=0)
    ow.init()
    follow(ow, 600, (1, 2, 5, 10, 20, 50, 100, 200, 300, 400, 600), "cold start,")
    t0 = time.time()
    fw = _oracle_thermalize_by_doubling(om, 1 / beta, 5, 1, k, 10, 0)
    st = fw.get_state()
    print(f"doubling ({k} levels x 10 sweeps, {time.time() - t0:.1f} s): n={st['num_operators']} num_worms={st['num_worms']:.1f} "
          f"avg_wl={st['avg_worm_length']:.0f}", flush=True)
    follow(fw, 100, (1, 5, 10, 20, 50, 100), "after doubling,")


if __name__ == "__main__":
    main()
