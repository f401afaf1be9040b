"""Writes tests/golden/trajectory_digests.json: SHA-256 digests of whole oracle trajectories (operators, state, stream
position, controller values and accumulated observables after init! + 60 unthermalised + 20 measured sweeps) for fixed
(model, T, seed, walker id).  Oracle and kernels agree bit for bit, so the digests pin BOTH against silent drift of what
they share: the random-stream contract (include/sse_rng.h), the table generator and the flattened model layout.
Entries marked lp_dependent rest on the LP optimum scipy-HiGHS returns (vertex_data.py) and are only meaningful with
the scipy version recorded in the file.   usage: python tests/golden/make_trajectory_digests.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import scipy  # noqa: E402

CASES = [  # (model class, T, seed, walker id, lp_dependent)
    ("heisenberg_det", 0.25, 11, 0, False),
    ("heisenberg_det", 1.0, 11, 5, False),
    ("heisenberg_eof", 0.25, 12, 1, True),
    ("spin1_dz", 0.4, 13, 2, True),
    ("mixed_honeycomb", 0.3, 14, 3, True),
    ("dimer_bilayer", 0.5, 15, 4, True),
]


def digest_of(state: dict, sums, counts) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(state["operators"], dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(state["state"], dtype=np.uint8).tobytes())
    h.update(np.array([state["num_operators"], state["rng_draws"]], dtype=np.uint64).tobytes())
    h.update(np.array([state["num_worms"], state["avg_worm_length"]], dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(counts, dtype=np.int64).tobytes())
    # the sums are compared to 1e-12 elsewhere (summation order differs between host and device); keep them out of the hash
    return h.hexdigest()


def run_oracle(name, T, seed, wid):
    from helpers import MODEL_CLASSES
    from oracle import OracleModel, OracleWalker

    om = OracleModel(model=MODEL_CLASSES[name]())
    ow = OracleWalker(om, T, seed=seed, walker_id=wid)
    ow.init()
    ow.sweep(60, thermalized=False)
    ow.sweep(20, thermalized=True, measure=True)
    sums, counts = ow.fetch_accumulators()
    return digest_of(ow.get_state(), sums, counts)


def main():
    out = dict(scipy=scipy.__version__, recipe="init! + 60 unthermalised + 20 measured sweeps, Philox stream", cases=[])
    for name, T, seed, wid, lp in CASES:
        out["cases"].append(dict(model=name, T=T, seed=seed, walker_id=wid, lp_dependent=lp, sha256=run_oracle(name, T, seed, wid)))
        print(out["cases"][-1])
    with open(os.path.join(HERE, "trajectory_digests.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
