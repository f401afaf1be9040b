"""Replays golden dumps produced by julia/dump_golden.jl from the REAL reference.

No dump can be generated in the build image (julia is absent), so tests/golden/reference_dump/ is empty and the
reference-replay tests skip; the replay machinery itself is exercised on a synthetic dump written by the oracle in
the same JSON schema (that self-check pins nothing about the reference).  When a maintainer drops real dumps into
tests/golden/reference_dump/, the oracle (CPU test) and the CUDA path (gpu test) must match them bit for bit."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import OracleModel, OracleWalker
from sse_b200.capi import build_model_desc, model_desc_from_model

DUMP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_dump")
DUMPS = sorted(glob.glob(os.path.join(DUMP_DIR, "*.json")))


def _desc_from_tables(t):
    """tables of julia/SSEB200.jl `flatten` -> sse_model_desc"""
    flat = dict(
        n_sites=len(t["site_dim"]), site_dim=t["site_dim"], n_bonds=len(t["bond_type"]), bond_type=t["bond_type"],
        bond_sites=t["bond_sites"], n_types=len(t["voff"]) - 1, type_dims=t["type_dims"], type_vertex_off=t["voff"],
        type_diag_off=t["doff"], n_vertices=t["voff"][-1], weights=t["weights"], signs=t["signs"],
        leg_states=t["leg_states"], diag_vertices=t["diag"], max_worm=t["max_worm"], trans_offset=t["trans_offset"],
        trans_count=t["trans_count"], n_outcomes=len(t["out_cumprob"]), out_cumprob=t["out_cumprob"],
        out_target=t["out_target"], out_leg=t["out_leg"], out_worm=t["out_worm"], energy_offset=t.get("energy_offset", 0.0))
    est = None
    if t.get("est") and t.get("n_est", 0):
        est = np.array(t["est"], dtype=np.float64).reshape(t["n_est"], flat["n_sites"], t["max_dim"])
    return build_model_desc(flat, t.get("norm_site_count", flat["n_sites"]), est)


def _replay(dump, make_backend):
    stream = np.array(dump["stream"], dtype=np.uint64)
    for step in dump["steps"]:
        b = step["before"]
        be = make_backend(stream)
        start = dict(num_operators=b["num_operators"], operators=np.array(b["operators"], dtype=np.uint64),
                     state=np.array(b["state"], dtype=np.uint8), T=dump["T"], num_worms=b["num_worms"],
                     avg_worm_length=b["avg_worm_length"], rng_draws=b["draws_before"])
        be.set(start)
        a = step["after_diagonal_update"]
        st = be.diagonal_update()
        assert np.array_equal(st["operators"], np.array(a["operators"], dtype=np.uint64))
        assert np.array_equal(st["state"], np.array(a["state"], dtype=np.uint8))
        assert st["num_operators"] == a["num_operators"] and st["rng_draws"] == a["draws"]
        v, vf, vl = be.vertex_list(len(a["operators"]))
        assert np.array_equal(v.reshape(-1, 2), np.array(step["vertex_list"]["vertices"], dtype=np.int64))
        assert np.array_equal(vf, np.array(step["vertex_list"]["v_first"], dtype=np.int64))
        assert np.array_equal(vl, np.array(step["vertex_list"]["v_last"], dtype=np.int64))
        w = step["after_worm_update"]
        st = be.worm_update()
        assert np.array_equal(st["operators"], np.array(w["operators"], dtype=np.uint64))
        assert np.array_equal(st["state"], np.array(w["state"], dtype=np.uint8))
        assert st["rng_draws"] == w["draws"]
        assert st["num_worms"] == pytest.approx(w["num_worms"], rel=1e-12)  # tanh: libm vs sse_tanh
        assert st["avg_worm_length"] == pytest.approx(w["avg_worm_length"], rel=1e-12)


class _OracleBackend:
    def __init__(self, om, T, stream):
        self.w = OracleWalker(om, T)
        self.w.set_injected_stream(stream)

    def set(self, s):
        self.w.set_state(s)

    def diagonal_update(self):
        self.w.diagonal_update()
        return self.w.get_state()

    def vertex_list(self, M):
        self.w.make_vertex_list()
        return self.w.get_vertex_list()

    def worm_update(self):
        self.w.worm_update(False)
        return self.w.get_state()


class _GpuBackend:
    def __init__(self, dm, T, stream):
        from sse_b200.walkers import Walkers

        self.g = Walkers(dm, [T], m_capacity=max(8192, 4 * len(stream) // 100))
        self.stream = stream

    def set(self, s):
        self.g.set_injected_stream(self.stream[None, :])
        self.g.set_state(0, s)

    def diagonal_update(self):
        self.g.dbg_diagonal_update()
        return self.g.get_state(0)

    def vertex_list(self, M):
        self.g.dbg_make_vertex_list()
        return self.g.dbg_get_vertex_list(0, M)

    def worm_update(self):
        self.g.dbg_worm_update(False)
        return self.g.get_state(0)


def _synthetic_dump(model, T, sweeps=4, seed=3):
    """Same schema as julia/dump_golden.jl, written by the oracle (self-check of the replay machinery only)."""
    desc, keep, sd = model_desc_from_model(model)
    om = OracleModel(desc=desc, keep=keep, sse_data=sd)
    stream = np.random.default_rng(seed).integers(0, 2**64, size=200_000, dtype=np.uint64)
    w = OracleWalker(om, T)
    w.set_injected_stream(stream)
    w.init()
    steps = []
    for _ in range(sweeps):
        b = w.get_state()
        before = dict(operators=b["operators"].tolist(), state=b["state"].tolist(), num_operators=b["num_operators"],
                      num_worms=b["num_worms"], avg_worm_length=b["avg_worm_length"], draws_before=b["rng_draws"])
        w.diagonal_update()
        a = w.get_state()
        w.make_vertex_list()
        v, vf, vl = w.get_vertex_list()
        w.worm_update(False)
        ww = w.get_state()
        steps.append(dict(
            before=before,
            after_diagonal_update=dict(operators=a["operators"].tolist(), state=a["state"].tolist(),
                                       num_operators=a["num_operators"], draws=a["rng_draws"]),
            vertex_list=dict(vertices=v.reshape(-1, 2).tolist(), v_first=vf.tolist(), v_last=vl.tolist()),
            after_worm_update=dict(operators=ww["operators"].tolist(), state=ww["state"].tolist(), num_worms=ww["num_worms"],
                                   avg_worm_length=ww["avg_worm_length"], draws=ww["rng_draws"])))
    return dict(name="selfcheck", T=T, stream=stream[: w.rng_draws].tolist(), steps=steps), (desc, keep, sd)


def test_replay_machinery_selfcheck_oracle():
    from helpers import MODEL_CLASSES

    dump, (desc, keep, sd) = _synthetic_dump(MODEL_CLASSES["spin1_dz"](), 0.3)
    dump = json.loads(json.dumps(dump))  # through JSON like a real dump
    om = OracleModel(desc=desc, keep=keep, sse_data=sd)
    _replay(dump, lambda stream: _OracleBackend(om, dump["T"], stream))


@pytest.mark.gpu
def test_replay_machinery_selfcheck_gpu():
    from helpers import MODEL_CLASSES
    from sse_b200.walkers import DeviceModel

    dump, (desc, keep, sd) = _synthetic_dump(MODEL_CLASSES["dimer_bilayer"](), 0.3)
    dump = json.loads(json.dumps(dump))
    dm = DeviceModel(desc=desc, keep=keep, sse_data=sd)
    _replay(dump, lambda stream: _GpuBackend(dm, dump["T"], stream))


@pytest.mark.skipif(not DUMPS, reason="no reference dumps (julia/dump_golden.jl cannot run in the build image)")
@pytest.mark.parametrize("path", DUMPS)
def test_oracle_matches_reference_dump(path):
    dump = json.load(open(path))
    desc, keep = _desc_from_tables(dump["tables"])
    om = OracleModel(desc=desc, keep=keep, sse_data=None)
    _replay(dump, lambda stream: _OracleBackend(om, dump["T"], stream))


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="no reference dumps (julia/dump_golden.jl cannot run in the build image)")
@pytest.mark.parametrize("path", DUMPS)
def test_gpu_matches_reference_dump(path):
    from sse_b200.walkers import DeviceModel

    dump = json.load(open(path))
    desc, keep = _desc_from_tables(dump["tables"])
    dm = DeviceModel(desc=desc, keep=keep, sse_data=None)
    _replay(dump, lambda stream: _GpuBackend(dm, dump["T"], stream))
