"""Pins the CPU oracle (oracle/sse_oracle.cpp) against every golden vector / known answer the reference's
own tests hold for the hot path.  CPU only."""
import json
import os

import numpy as np
import pytest

import sse_b200 as S
from ed import run_ed
from helpers import bani_honeycomb, heisenberg_square, isconsistent, mixed_honeycomb, random_stream
from mcstats import run_oracle_task
from oracle import OracleModel, OracleWalker
from sse_b200.capi import build_model_desc
from sse_b200.estimators import all_magnetization_estimators
from sse_b200.sse_data import SSEBond, SSEData
from sse_b200.util import opercode, vertex_code
from sse_b200.vertex_data import make_vertex_data

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bani2v2o8_golden.json")


def test_vertex_list_golden_vector():
    """test/test_vertex_list.jl:1-30 (exact link arrays)"""
    sp, sz = S.spin_operators(2)
    H = np.kron(sz, sz) + 0.5 * (np.kron(sp, sp.T) + np.kron(sp.T, sp))
    vd = make_vertex_data((2, 2), H, energy_offset_factor=0.0)
    sd = SSEData([vd], [SSEBond(1, (1, 2)), SSEBond(1, (2, 3)), SSEBond(1, (1, 3))])
    flat = sd.flatten()
    flat["n_sites"] = 4
    flat["site_dim"] = np.array([2, 2, 2, 2], dtype=np.uint8)
    desc, keep = build_model_desc(flat, 4, None)
    om = OracleModel(desc=desc, keep=keep, sse_data=sd)
    v = vertex_code(False, 1)
    ops = np.array([0, 0, opercode(1, v), 0, opercode(1, v), opercode(2, v), opercode(3, v)], dtype=np.uint64)
    w = OracleWalker(om, 1.0)
    w.set_state(dict(num_operators=4, operators=ops, state=np.ones(4, dtype=np.uint8), T=1.0))
    w.make_vertex_list()
    vert, vf, vl = w.get_vertex_list()
    expected = -np.ones((7, 4, 2), dtype=np.int64)
    expected[2] = [(3, 7), (3, 6), (1, 5), (2, 5)]
    expected[4] = [(3, 3), (4, 3), (1, 7), (1, 6)]
    expected[5] = [(4, 5), (4, 7), (2, 3), (2, 7)]
    expected[6] = [(3, 5), (4, 6), (1, 3), (2, 6)]
    assert np.array_equal(vert, expected)
    assert vl.tolist() == [[3, 7], [3, 6], [4, 7], [-1, -1]]
    assert vf.tolist() == [[1, 3], [2, 3], [2, 6], [-1, -1]]


def test_worm_traverse_consistency():
    """test/test_sse.jl:30-60"""
    model = heisenberg_square(4, True)
    om = OracleModel(model)
    sd = om.sse_data
    vd = sd.get_vertex_data(1)
    for ops in ([opercode(1, int(vd.diagonal_vertices[2]))],
                [opercode(1, vertex_code(True, 1)), opercode(1, vertex_code(True, 1))]):
        ops = np.array(ops, dtype=np.uint64)
        ls = vd.get_leg_state((int(ops[0]) & ((1 << 25) - 1)) >> 1)
        state = np.ones(16, dtype=np.uint8)
        state[sd.bonds[0].sites[0] - 1], state[sd.bonds[0].sites[1] - 1] = ls[0], ls[1]
        w = OracleWalker(om, 0.1, seed=3)
        w.set_state(dict(num_operators=len(ops), operators=ops, state=state, T=0.1))
        w.make_vertex_list()
        w.worm_traverse(1, 1, 1)
        st = w.get_state()
        _, vf, _ = w.get_vertex_list()
        state0 = np.ones(16, dtype=np.int64)
        for s in range(16):
            if vf[s, 0] > 0:
                op = int(st["operators"][vf[s, 1] - 1])
                state0[s] = sd.get_vertex_data(op >> 26).get_leg_state((op & ((1 << 25) - 1)) >> 1)[vf[s, 0] - 1]
        assert isconsistent(st["operators"], state0, sd)


def test_sse_mc_1000_sweeps_consistent():
    """test/test_sse.jl:63-94: mixed spin-1/2 / spin-1 honeycomb 4x4, T=0.1, 1000 sweeps"""
    model = mixed_honeycomb(4)
    om = OracleModel(model)
    w = OracleWalker(om, 0.1, seed=17)
    w.init()
    for s in range(1000):
        if s % 50 == 0:
            st = w.get_state()
            assert isconsistent(st["operators"], st["state"], om.sse_data)
        w.sweep(1, thermalized=s > 100)
    st = w.get_state()
    assert isconsistent(st["operators"], st["state"], om.sse_data)
    dims = np.array([x.dim for x in om.sse_data.sites])
    assert np.all(st["state"] > 0) and np.all(st["state"] <= dims)
    assert w.flags == 0


def test_injected_stream_equals_philox_stream():
    """The two stream kinds are the same contract: injecting the Philox words reproduces the Philox run."""
    from sse_b200 import capi  # noqa: F401
    import ctypes as C

    model = bani_honeycomb(3)
    om = OracleModel(model)
    a = OracleWalker(om, 0.4, seed=9, walker_id=5)
    a.init()
    a.sweep(12)
    n = a.rng_draws
    # regenerate the words through the shared header via the oracle itself: a second walker consuming an
    # injected copy of walker a's stream must land in the same state
    b = OracleWalker(om, 0.4, seed=9, walker_id=5)
    words = np.array([_philox(9, 5, k) for k in range(n)], dtype=np.uint64)
    b.set_injected_stream(words)
    b.init()
    b.sweep(12)
    assert not b.stream_exhausted and b.rng_draws == n
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa["operators"], sb["operators"]) and np.array_equal(sa["state"], sb["state"])


def _philox(seed, walker, k):
    """Pure-Python draw k of the Philox stream (include/sse_rng.h): block k>>1, word pair k&1."""
    M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    j, half = k >> 1, k & 1
    c = [j & mask, (j >> 32) & mask, walker & mask, (walker >> 32) & mask]
    k0, k1 = seed & mask, (seed >> 32) & mask
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & mask, (p0 >> 32) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    return (c[2] | (c[3] << 32)) if half else (c[0] | (c[1] << 32))


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors): counter/key all zero and all ones."""
    def block(ctr, key):
        M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
        c, (k0, k1) = list(ctr), key
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & mask, (p0 >> 32) ^ c[3] ^ k1, p0 & mask]
            k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
        return c
    assert block([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    assert block([f, f, f, f], (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox(0, 0, 0) == 0x6627e8d5 | (0xe169c58d << 32)
    assert _philox(0, 0, 1) == 0xbc57ac4c | (0x9b00dbd8 << 32)


ED_JOBS = {
    # test/test_jobs.jl:4-30
    "magnet_square": lambda: S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(2, 4)), J=1.23, hz=-0.2,
                                                measure=all_magnetization_estimators(2))),
    # test/test_jobs.jl:32-67
    "honeycomb": lambda: S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.honeycomb, size=(2, 2)),
                                            parameter_map=dict(S=["Sa", "Sb"], J=["J1", "J2", "J3"]), J1=1.0, J2=0.5, J3=1.0,
                                            d=0.2, Dz=0.2, Dx=0.5, Sa=0.5, Sb=1, measure=all_magnetization_estimators(2))),
    # test/test_jobs.jl:69-110
    "fully_frustrated_bilayer": lambda: S.ClusterModel(dict(
        lattice=dict(unitcell=S.UnitCells.fully_frust_square_bilayer, size=(2, 2)), cluster_bases=(S.ClusterBases.dimer,),
        measure_quantum_numbers=[dict(name="", quantum_number=2)],
        parameter_map=dict(S=["Sa", "Sb"], J=["JD", "JxP", "JxP", "JyP", "JyP", "JxX", "JxX", "JyX", "JyX"]),
        JD=1, JxP=0.55, JyP=0.6, JxX=0.7, JyX=0.3, d=-0.2, Dz=0.2, Dx=0.5, Sa=0.5, Sb=0.5)),
}


@pytest.mark.parametrize("job", list(ED_JOBS))
def test_ed_compare_oracle(job):
    """test/test_ed_compare.jl:15-62 on the oracle (fewer sweeps than the reference's 40000 to keep the CPU suite
    short): every observable ED provides within 4 sigma at every temperature."""
    model = ED_JOBS[job]()
    om = OracleModel(model)
    Ts = np.linspace(0.04, 4.0, 7)
    ests = model.get_opstring_estimators() if job != "fully_frustrated_bilayer" else []  # test/ed/cluster.jl:4-14
    ed = run_ed(model, Ts, ests)
    worst = 0.0
    for it, T in enumerate(Ts):
        res = run_oracle_task(om, model, float(T), sweeps=12000, therm=2000, binsize=200, seed=124535, walker_id=it)
        for name, vals in ed.items():
            mean, err = res[name]
            z = abs(mean - vals[it]) / (err if err > 0 else 1e-8)
            worst = max(worst, z)
            assert z <= 4.5, f"{job} T={T:.3f} {name}: MC {mean} +- {err} vs ED {vals[it]} (z={z:.2f})"
    assert worst > 0


def test_bani2v2o8_published_points_oracle():
    """docs/src/bani2v2o8.results.json (BASELINE config 4): the oracle reproduces the reference's published
    observables on a subset of the L=10 tasks (the GPU test covers all 40 points)."""
    golden = json.load(open(GOLDEN))["tasks"]
    model = bani_honeycomb(10)
    om = OracleModel(model)
    zs = []
    for t in golden:
        if t["L"] != 10 or t["T"] < 1.2 or int(t["task"][-2:]) % 3 != 0:
            continue
        res = run_oracle_task(om, model, t["T"], sweeps=12000, therm=1500, binsize=200, seed=7, walker_id=int(t["task"][-2:]))
        for name in ("Energy", "OperatorCount", "AbsMag", "Mag2", "MagChi"):
            mean, err = res[name]
            gm, ge = t[name]
            zs.append((mean - gm) / np.hypot(err, ge))
    zs = np.array(zs)
    assert len(zs) >= 20
    assert np.all(np.abs(zs) < 4.5), zs
    assert abs(zs.mean()) < 1.0 and zs.std() < 1.8
