"""Host-side builder of the abstract-loop probability tables (`VertexData`).

Restates /root/reference/src/vertex_data.jl:13-88,129-416 on numpy + scipy's HiGHS
(`scipy.optimize.linprog(method="highs")`) so the kernels have inputs without Julia.  The tables are
an INPUT to the device path; for dim>2 the LP optimum HiGHS returns is not
unique, so only the structural properties the reference tests pin (detailed balance, normalisation,
the unique S=1/2 solution, test/test_vertex_data.jl:1-135) are guaranteed to agree with a Julia run.

Conventions kept from the reference: everything 1-based (vertex indices, leg indices, worm indices,
state indices), legs 1..NSites = bottom (ket of row index i), NSites+1..2NSites = top.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import linprog

from .util import (
    INVALID_VERTEX_CODE,
    join_idx,
    split_idx,
    vertex_code,
    worm_action,
    worm_count,
    worm_inverse,
)


def site_of_leg(leg: int, num_sites: int) -> int:
    """vertex_data.jl:190"""
    return leg - num_sites if leg > num_sites else leg


def calc_energy_offset(H: np.ndarray, energy_offset_factor: float) -> float:
    """vertex_data.jl:129-135"""
    d = np.diag(H)
    hmin, hmax = float(d.min()), float(d.max())
    epsilon = (hmax - hmin) * energy_offset_factor
    return -(hmax + epsilon)


def construct_vertices(dims, H: np.ndarray, energy_offset: float, tolerance: float):
    """Enumerate non-zero matrix elements column-major (vertex_data.jl:137-169).

    Returns (diagonal_vertices [prod(dims)] of VertexCode ints, weights, leg_states [2N, nv], signs).
    """
    nsites = len(dims)
    total = H.shape[0]
    diagonal_vertices = np.full(total, INVALID_VERTEX_CODE, dtype=np.int64)
    weights = []
    leg_states = []
    signs = []
    rdims = tuple(reversed(dims))
    # CartesianIndices(bond_hamiltonian) iterates the first index (row i) fastest
    for j in range(1, total + 1):
        for i in range(1, total + 1):
            w = -H[i - 1, j - 1]
            if i == j:
                w -= energy_offset
            if abs(w) > tolerance:
                si = tuple(reversed(split_idx(rdims, i)))
                sj = tuple(reversed(split_idx(rdims, j)))
                if i == j:
                    reversed_i = join_idx(dims, si)
                    diagonal_vertices[reversed_i - 1] = vertex_code(True, len(weights) + 1)
                leg_states.append(si + sj)
                weights.append(abs(w))
                signs.append(1 if w >= 0 else -1)
    ls = np.array(leg_states, dtype=np.uint8).reshape(-1, 2 * nsites).T.copy()  # [leg, vertex]
    return diagonal_vertices, np.array(weights, dtype=np.float64), ls, np.array(signs, dtype=np.int8)


def wrap_vertex_idx(leg_states: np.ndarray, vidx):
    """vertex_data.jl:171-187: VertexCode with the diagonal flag derived from the leg states."""
    if vidx is None:
        return INVALID_VERTEX_CODE
    nl = leg_states.shape[0]
    ls = leg_states[:, vidx - 1]
    diagonal = bool(np.all(ls[: nl // 2] == ls[nl // 2:]))
    return vertex_code(diagonal, vidx)


class _VertexFinder:
    """dict lookup replacing the linear `findfirst` of vertex_apply_change (vertex_data.jl:210)."""

    def __init__(self, leg_states: np.ndarray):
        self.map = {}
        for v in range(leg_states.shape[1] - 1, -1, -1):  # keep the FIRST match like findfirst
            self.map[bytes(leg_states[:, v])] = v + 1

    def __call__(self, ls: np.ndarray):
        return self.map.get(bytes(ls))


def vertex_apply_change(leg_states, dims, vertex, step_in, step_out, finder=None):
    """vertex_data.jl:192-211 (1-based vertex / legs / worms); returns vertex index or None."""
    nsites = len(dims)
    new = leg_states[:, vertex - 1].copy()
    leg_in, worm_in = step_in
    leg_out, worm_out = step_out
    dim_in = dims[site_of_leg(leg_in, nsites) - 1]
    dim_out = dims[site_of_leg(leg_out, nsites) - 1]
    new[leg_in - 1] = worm_action(worm_in, int(new[leg_in - 1]), dim_in)
    new[leg_out - 1] = worm_action(worm_out, int(new[leg_out - 1]), dim_out)
    if finder is None:
        finder = _VertexFinder(leg_states)
    return finder(new)


def construct_transitions(weights, leg_states, max_worm_count, dims, tolerance, lp_tolerance):
    """Directed-loop LP per worm-connected vertex class (vertex_data.jl:213-416).

    Returns (trans_offset [leg, worm, vertex] int64 (1-based offset, -1 invalid),
             trans_length [leg, worm, vertex] (= count-1), cumprobs, targets (VertexCode), step_outs [(leg, worm)]).
    """
    nsites = len(dims)
    leg_count = 2 * nsites
    nv = len(weights)
    trans_offset = -np.ones((leg_count, max_worm_count, nv), dtype=np.int64)
    trans_length = np.zeros((leg_count, max_worm_count, nv), dtype=np.int64)
    cumprobs: list[float] = []
    targets_out: list[int] = []
    step_outs: list[tuple[int, int]] = []

    steps = []
    inv_steps = {}
    step_idx = {}
    for worm in range(1, max_worm_count + 1):
        for leg in range(1, leg_count + 1):
            dim = dims[site_of_leg(leg, nsites) - 1]
            if worm <= worm_count(dim):
                steps.append((leg, worm))
                step_idx[(leg, worm)] = len(steps) - 1
                inv_steps[(leg, worm)] = (leg, worm_inverse(worm, dim))
    ns = len(steps)

    def inverse(vc):
        return (inv_steps[vc[1]], inv_steps[vc[0]])

    variables = []
    var_index = {}
    for s_in in steps:
        for s_out in steps:
            vc = (s_in, s_out)
            if inverse(vc) not in var_index:
                var_index[vc] = len(variables)
                variables.append(vc)
    nvar = len(variables)

    def find_var(s_in, s_out):
        vc = (s_in, s_out)
        if vc in var_index:
            j = var_index[vc]
            ivc = inverse(vc)
            if ivc in var_index:
                j = min(j, var_index[ivc])
            return j
        return var_index[inverse(vc)]

    cost = np.array([1.0 if vc == inverse(vc) else 0.0 for vc in variables])  # discourage bounces

    A = np.zeros((ns, nvar))
    for r, step in enumerate(steps):
        for j, vc in enumerate(variables):
            if vc[0] == step or inverse(vc)[0] == step:
                A[r, j] = 1.0

    finder = _VertexFinder(leg_states)
    used_inaccurate_truncations = False

    while True:
        v = None
        step_in = None
        for empty_v in range(1, nv + 1):
            for s in steps:
                if trans_offset[s[0] - 1, s[1] - 1, empty_v - 1] < 0:
                    v = empty_v
                    step_in = inv_steps[s]
                    break
            if v is not None:
                break
        if v is None:
            break

        targets = {}
        constraints = np.zeros(ns)
        for s_out in steps:
            t = vertex_apply_change(leg_states, dims, v, step_in, s_out, finder)
            targets[s_out] = t
            constraints[step_idx[s_out]] = weights[t - 1] if t is not None else 0.0

        res = linprog(
            cost,
            A_eq=A,
            b_eq=constraints,
            bounds=(0, None),
            method="highs",
            options={
                "primal_feasibility_tolerance": lp_tolerance,
                "dual_feasibility_tolerance": lp_tolerance,
                "presolve": True,
            },
        )
        if res.status != 0:
            raise RuntimeError(f"transition probability optimization failed: {res.message}")
        solution = res.x

        for s_in in steps:
            if targets[s_in] is None:
                continue
            c = constraints[step_idx[s_in]]
            norm = 1.0 if c == 0 else c
            offset = len(cumprobs) + 1
            length = -1
            acc = 0.0
            for s_out in steps:
                var = find_var(s_in, s_out)
                prob = solution[var] / norm
                if solution[var] < -lp_tolerance * 10:
                    raise AssertionError("LP solution negative beyond tolerance")
                if prob < 0.0:
                    used_inaccurate_truncations = prob < -tolerance
                    prob = 0.0
                if prob > 1.0:
                    used_inaccurate_truncations = prob > 1 + tolerance
                    prob = 1.0
                if prob > tolerance / ns:
                    # cumsum! over the slice (vertex_data.jl:401-402): sequential left-to-right sum
                    acc = acc + prob if length >= 0 else prob
                    cumprobs.append(acc)
                    tgt = targets[inv_steps[s_out]]
                    code = wrap_vertex_idx(leg_states, tgt)
                    assert code != INVALID_VERTEX_CODE
                    targets_out.append(code)
                    step_outs.append(s_out)
                    length += 1
            tv = targets[s_in]
            trans_offset[s_in[0] - 1, s_in[1] - 1, tv - 1] = offset
            trans_length[s_in[0] - 1, s_in[1] - 1, tv - 1] = length
            if length >= 0 and abs(cumprobs[-1] - 1) > tolerance:
                warnings.warn(f"normalization error: {cumprobs[-1]} != 1")

    if used_inaccurate_truncations:
        warnings.warn("had to truncate some probabilities in a possibly inaccurate way!")

    return (
        trans_offset,
        trans_length,
        np.array(cumprobs, dtype=np.float64),
        np.array(targets_out, dtype=np.int64),
        np.array(step_outs, dtype=np.int64).reshape(-1, 2),
    )


@dataclass
class VertexData:
    """Mirror of `VertexData{NSites}` (vertex_data.jl:13-28).  All indices 1-based as in the reference."""

    energy_offset: float
    dims: tuple
    diagonal_vertices: np.ndarray  # [prod(dims)] VertexCode ints (INVALID_VERTEX_CODE = invalid)
    signs: np.ndarray  # int8 [nv]
    weights: np.ndarray  # f64 [nv]
    trans_offset: np.ndarray  # int64 [leg, worm, vertex], 1-based offset into the flat arrays, -1 invalid
    trans_length: np.ndarray  # int64 [leg, worm, vertex], count-1
    transition_cumprobs: np.ndarray
    transition_targets: np.ndarray  # VertexCode ints
    transition_step_outs: np.ndarray  # [n, 2] (leg, worm)
    leg_states: np.ndarray  # uint8 [leg, vertex]
    _cache_key: tuple = field(default=None, repr=False)

    @property
    def nsites(self) -> int:
        return len(self.dims)

    def vertex_count(self) -> int:
        return len(self.weights)

    def get_diagonal_vertex(self, compound_state_idx: int) -> int:
        return int(self.diagonal_vertices[compound_state_idx - 1])

    def get_vertex_weight(self, v: int) -> float:
        return 0.0 if v >= INVALID_VERTEX_CODE - 1 else float(self.weights[(v >> 1) - 1])

    def get_sign(self, v: int) -> int:
        return int(self.signs[(v >> 1) - 1])

    def get_leg_state(self, v: int) -> np.ndarray:
        return self.leg_states[:, (v >> 1) - 1]

    def scatter(self, v: int, leg_in: int, worm_in: int, random: float):
        """vertex_data.jl:106-125"""
        vi = v >> 1
        off = int(self.trans_offset[leg_in - 1, worm_in - 1, vi - 1])
        ln = int(self.trans_length[leg_in - 1, worm_in - 1, vi - 1])
        for out in range(off, off + ln + 1):
            if random < self.transition_cumprobs[out - 1]:
                leg_out, worm_out = self.transition_step_outs[out - 1]
                return int(leg_out), int(worm_out), int(self.transition_targets[out - 1])
        return -1, -1, INVALID_VERTEX_CODE


_VD_CACHE: dict = {}


def make_vertex_data(
    dims,
    bond_hamiltonian,
    energy_offset_factor: float = 0.25,
    tolerance: float = 1e-7,
    lp_tolerance: float = 1e-10,
) -> VertexData:
    """`VertexData(dims, H; energy_offset_factor, tolerance, lp_tolerance)` (vertex_data.jl:48-88)."""
    dims = tuple(int(d) for d in dims)
    H = np.asarray(bond_hamiltonian, dtype=np.float64)
    key = (dims, H.tobytes(), float(energy_offset_factor), float(tolerance), float(lp_tolerance))
    if key in _VD_CACHE:
        return _VD_CACHE[key]
    nsites = len(dims)
    assert nsites >= 1
    total_dim = int(np.prod(dims))
    assert H.shape == (total_dim, total_dim)
    energy_offset = calc_energy_offset(H, energy_offset_factor)
    max_worm_count = max(worm_count(d) for d in dims)
    diagonal_vertices, weights, leg_states, signs = construct_vertices(dims, H, energy_offset, tolerance)
    assert leg_states.shape == (2 * nsites, len(weights))
    to, tl, cp, tg, so = construct_transitions(weights, leg_states, max_worm_count, dims, tolerance, lp_tolerance)
    vd = VertexData(energy_offset, dims, diagonal_vertices, signs, weights, to, tl, cp, tg, so, leg_states)
    _VD_CACHE[key] = vd
    return vd
