"""Known-answer digests of whole trajectories (tests/golden/trajectory_digests.json, written by
tests/golden/make_trajectory_digests.py): the oracle, the emulated kernel source and the GPU must all reproduce them.
They pin what oracle and kernels share — the random-stream contract of include/sse_rng.h, the table generator, the
flattened model layout — against silent drift that a pure oracle-vs-kernel comparison cannot see."""
import json
import os
import sys

import numpy as np
import pytest
import scipy

from test_emu_parity import emu, emu_built  # noqa: F401  (fixtures)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN_DIR)
from make_trajectory_digests import digest_of, run_oracle  # noqa: E402

GOLD = json.load(open(os.path.join(GOLDEN_DIR, "trajectory_digests.json")))
CASES = [c for c in GOLD["cases"] if not c["lp_dependent"] or GOLD["scipy"] == scipy.__version__]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['model']}-T{c['T']}")
def test_oracle_reproduces_digest(case):
    assert run_oracle(case["model"], case["T"], case["seed"], case["walker_id"]) == case["sha256"]


def _device_digests(shape):
    from helpers import MODEL_CLASSES
    from sse_b200.walkers import DeviceModel, Walkers

    for case in CASES:
        dm = DeviceModel(model=MODEL_CLASSES[case["model"]]())
        # the walker of interest sits in the middle of a small batch
        wid = case["walker_id"]
        gw = Walkers(dm, np.full(wid + 3, case["T"]), m_capacity=16384, seed=case["seed"])
        gw.set_launch_shape(*shape)
        gw.init()
        gw.sweep(60, thermalized=False)
        gw.sweep(20, thermalized=True, measure=True)
        sums, counts = gw.fetch_accumulators()
        assert digest_of(gw.get_state(wid), sums[wid], counts[wid]) == case["sha256"], case


def test_emu_reproduces_digests(emu):
    _device_digests((1, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(0, 0), (1, 1), (8, 8)])
def test_gpu_reproduces_digests(shape):
    _device_digests(shape)
