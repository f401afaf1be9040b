#!/usr/bin/env python
"""tests/test_gpu_statistics.py::test_bani2v2o8_published_results_gpu evaluated WITHOUT a GPU: the CPU oracle follows the
same Philox streams as the device walkers bit for bit (tests/test_zz_beta_doubling.py), so running the test's protocol on
the oracle gives the z-scores the GPU test will see.  Used to check a protocol change when no GPU time is left.
usage: bani_protocol_on_oracle.py L sweeps replicas seed therm [processes] > profiles/r2_bani_protocol_L<L>.json"""
import json
import os
import sys
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

L, SWEEPS, REPLICAS, SEED, THERM = (int(x) for x in sys.argv[1:6])
PROCS = int(sys.argv[6]) if len(sys.argv) > 6 else os.cpu_count()
DOUBLINGS = 3
_ctx = {}


def _setup():
    if not _ctx:
        from helpers import bani_honeycomb
        from oracle import OracleModel

        model = bani_honeycomb(L)
        golden = [t for t in json.load(open(os.path.join(ROOT, "tests", "golden", "bani2v2o8_golden.json")))["tasks"] if t["L"] == L]
        _ctx.update(model=model, om=OracleModel(model), golden=golden, Ts=[t["T"] for t in golden])
    return _ctx


def one_walker(w):
    from test_zz_beta_doubling import _oracle_thermalize_by_doubling

    c = _setup()
    T = float(c["Ts"][w // REPLICAS])
    fw = _oracle_thermalize_by_doubling(c["om"], T, SEED, w, DOUBLINGS, max(10, THERM // 10), 0)
    fw.sweep(THERM, thermalized=False)
    fw.sweep(SWEEPS, thermalized=True, measure=True)  # one bin per replica
    sums, counts = fw.fetch_accumulators()
    return w, sums.tolist(), counts.tolist()


def main():
    from mcstats import evaluate, obs_names

    c = _setup()
    nT = len(c["Ts"])
    with Pool(PROCS) as pool:
        res = sorted(pool.imap_unordered(one_walker, range(nT * REPLICAS), chunksize=1))
    names = obs_names(c["model"])
    zs = {}
    for it, t in enumerate(c["golden"]):
        bins = {}
        for i, n in enumerate(names):
            vals = []
            for w in range(it * REPLICAS, (it + 1) * REPLICAS):
                _, sums, counts = res[w]
                cnt = counts[1] if n == "WormLengthFraction" else counts[0]
                vals.append(sums[i] / max(cnt, 1))
            bins[n] = np.array(vals)
        r = evaluate(c["model"], bins)
        for name in ("Energy", "OperatorCount", "AbsMag", "Mag2", "Mag4", "MagChi", "BinderRatio", "SpecificHeat"):
            mean, err = r[name]
            gm, ge = t[name]
            zs.setdefault(name, []).append(float((mean - gm) / np.hypot(err, ge)))
    allz = np.concatenate([np.array(v) for v in zs.values()])
    out = {"L": L, "sweeps": SWEEPS, "replicas": REPLICAS, "seed": SEED, "therm": THERM, "n_z": int(len(allz)),
           "max_abs_z": float(np.abs(allz).max()), "mean_z": float(allz.mean()), "std_z": float(allz.std()),
           "frac_below_3": float(np.mean(np.abs(allz) < 3.0)),
           "passes": bool(np.all(np.abs(allz) < 4.5) and abs(allz.mean()) < 0.6 and allz.std() < 1.6 and np.mean(np.abs(allz) < 3.0) > 0.98),
           "z": {k: np.round(v, 2).tolist() for k, v in zs.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
