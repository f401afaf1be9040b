"""One-off soak run (not part of the suite): the emulated kernel source against the oracle over many seeds,
temperatures, model classes and launch shapes.  Usage: python tests/emu/soak.py [rounds]  ->  one line per case,
exit code 1 on the first mismatch.  Every third case thermalises through sse_advance with random visit budgets (walkers
parked between sweeps, between worms and inside worms) instead of sse_sweep.  Outcomes: profiles/r1_emu_soak.txt (round 1),
profiles/r2_emu_soak.txt (round-2 kernel)."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

from sse_b200 import capi  # noqa: E402

capi.LIB_PATH = os.path.join(HERE, "libsse_b200_emu.so")

from helpers import MODEL_CLASSES  # noqa: E402
from oracle import OracleModel, OracleWalker  # noqa: E402
from sse_b200.capi import model_desc_from_model  # noqa: E402
from sse_b200.walkers import DeviceModel, Walkers  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    rng = np.random.default_rng(2026)
    t0 = time.time()
    cases = 0
    for name, make in MODEL_CLASSES.items():
        model = make()
        desc, keep, sd = model_desc_from_model(model)
        dm = DeviceModel(model=model, desc=desc, keep=keep, sse_data=sd)
        om = OracleModel(desc=desc, keep=keep, sse_data=sd)
        for r in range(rounds):
            for ci, chains in enumerate(((1, 1), (2, 3), (1, 8))):  # launch shapes (worm warps, stream warps)
                W = int(rng.integers(3, 14))
                Ts = rng.uniform(0.15, 2.5, size=W)  # below ~0.1 the S=1 model launches 1e8-visit worms early on
                seed = int(rng.integers(1, 2**40))
                off = int(rng.integers(0, 1000))
                n_th, n_ms = int(rng.integers(5, 40)), int(rng.integers(1, 12))
                level = (r + 2 * ci) % 3  # shared-memory level of the stream warps: 0 = all global, 1 = state + tags, 2 = + vfirst/vlast
                os.environ["SSE_B200_SMEM_LEVEL"] = str(level)
                gw = Walkers(dm, Ts, m_capacity=16384, seed=seed, walker_id_offset=off)
                gw.set_launch_shape(*chains)
                gw.init()
                parked = (r + ci) % 3 == 2
                if parked:  # random visit budgets, then complete the sweeps in flight: every walker did its own number of sweeps
                    for _ in range(int(rng.integers(2, 6))):
                        gw.advance(int(rng.integers(50, 5000)), thermalized=False)
                    gw.finish_sweeps(thermalized=False)
                    done, in_flight = gw.progress()
                    assert not in_flight.any()
                else:
                    gw.sweep(n_th, thermalized=False)
                    done = np.full(W, n_th)
                gw.sweep(n_ms, thermalized=True, measure=True)
                sums, counts = gw.fetch_accumulators()
                for i in range(W):
                    ow = OracleWalker(om, float(Ts[i]), seed=seed, walker_id=off + i)
                    ow.init()
                    ow.sweep(int(done[i]), thermalized=False)
                    ow.sweep(n_ms, thermalized=True, measure=True)
                    a, b = gw.get_state(i), ow.get_state()
                    ok = (a["num_operators"] == b["num_operators"] and np.array_equal(a["operators"], b["operators"])
                          and np.array_equal(a["state"], b["state"]) and a["rng_draws"] == b["rng_draws"]
                          and a["num_worms"] == b["num_worms"])
                    osums, ocounts = ow.fetch_accumulators()
                    ok = ok and np.array_equal(counts[i], ocounts) and np.allclose(sums[i], osums, rtol=1e-12, atol=1e-300)
                    if not ok:
                        print(f"MISMATCH {name} shape={chains} seed={seed} walker={i} T={Ts[i]}")
                        sys.exit(1)
                cases += W
                print(f"ok {name:16s} shape={chains} level={level} walkers={W:2d} sweeps={'advance' if parked else n_th}+{n_ms} seed={seed}", flush=True)
    print(f"soak ok: {cases} walker runs bit-identical to the oracle in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
