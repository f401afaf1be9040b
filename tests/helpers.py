"""Shared fixtures for the parity tests: model factories for the model classes SURVEY.md §7 asks for
(S=1/2 Heisenberg deterministic; S=1 Dz stochastic 17-vertex table; dimer-basis 4x4 table; mixed-spin
honeycomb of test/test_sse.jl:63-94) and the reference's `isconsistent` invariant."""
import numpy as np

import sse_b200 as S
from sse_b200.util import op_bond, op_vertex


def heisenberg_chain(L, measure=("magnetization", "staggered_magnetization"), **kw):
    return S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.chain, size=(L,)), J=1.0, measure=list(measure),
                              s_half_deterministic=True, **kw))


def heisenberg_square(L, deterministic=True, measure=("magnetization", "staggered_magnetization")):
    return S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(L, L)), J=1.0, measure=list(measure),
                              s_half_deterministic=deterministic))


def bani_honeycomb(L):
    return S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.honeycomb, size=(L, L)), S=1, J=1.0,
                              Dz=0.04556 / 8.07, measure=["magnetization"]))


def mixed_honeycomb(L=4):
    """test/test_sse.jl:63-77"""
    return S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.honeycomb, size=(L, L)), J=1.4, S1=0.5, S2=1,
                              parameter_map=dict(S=("S1", "S2")), measure=[]))


def dimer_bilayer(L, JD=0.5, JP=1.0):
    """test/test_jobs.jl:139-167 scaled to L"""
    return S.ClusterModel(dict(lattice=dict(unitcell=S.UnitCells.fully_frust_square_bilayer, size=(L, L)),
                               cluster_bases=(S.ClusterBases.dimer,),
                               measure_quantum_numbers=[dict(name="", quantum_number=2)],
                               parameter_map=dict(S=["Sa", "Sb"], J=["JD"] + ["JP"] * 8), JD=JD, JP=JP, Sa=0.5, Sb=0.5))


MODEL_CLASSES = {
    "heisenberg_det": lambda: heisenberg_square(4, True),
    "heisenberg_eof": lambda: heisenberg_square(4, False),
    "spin1_dz": lambda: bani_honeycomb(3),
    "mixed_honeycomb": lambda: mixed_honeycomb(3),
    "dimer_bilayer": lambda: dimer_bilayer(3),
}


def isconsistent(operators, state0, sse_data) -> bool:
    """test/test_sse.jl:5-28 — leg states chain correctly through the operator string."""
    state = np.array(state0, dtype=np.int64).copy()
    for op in operators:
        op = int(op)
        if op == 0:
            continue
        b = sse_data.bonds[op_bond(op) - 1]
        sites = [s - 1 for s in b.sites]
        ls = sse_data.get_vertex_data(op_bond(op)).get_leg_state(op_vertex(op))
        dims = [sse_data.sites[s].dim for s in sites]
        if not all(state[s] <= d for s, d in zip(sites, dims)):
            return False
        if list(state[sites]) != [int(x) for x in ls[: len(sites)]]:
            return False
        state[sites] = ls[len(sites):]
    return True


def random_stream(rng, n):
    return rng.integers(0, 2**64, size=n, dtype=np.uint64)
