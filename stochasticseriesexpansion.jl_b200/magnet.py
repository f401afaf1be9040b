"""`MagnetModel`: arbitrary-spin XXZ magnet -> bond Hamiltonians -> SSEData (host model setup).

Mirror of /root/reference/src/models/magnet/magnet.jl (:1-211) and the `AbstractModel` interface of
/root/reference/src/abstract_model.jl:17-54 (`generate_sse_data`, `get_opstring_estimators`,
`leg_count`, `normalization_site_count`).  Parameters are a plain dict keyed by the reference's
task-parameter names (`lattice`, `S`, `J`, `d`, `hz`, `Dz`, `Dx`, `parameter_map`, `measure`).
"""
from __future__ import annotations

from dataclasses import dataclass
from fractions import Fraction

import numpy as np

from .estimators import MagnetizationEstimator
from .lattice import Lattice, neel_vector
from .operators import spin_operators
from .sse_data import SSEBond, SSEData
from .vertex_data import make_vertex_data


@dataclass(frozen=True)
class MagnetBondParams:
    J: float
    d: float
    Dx: tuple
    Dz: tuple
    hz: tuple


@dataclass(frozen=True)
class MagnetSiteParams:
    spin_states: int


class ParameterMap:
    def __init__(self, mapping):
        self.map = mapping

    def get(self, name: str, index: int) -> str:
        """get_parameter (magnet.jl:68-79); index is 1-based."""
        if self.map is None:
            return name
        res = self.map
        if name not in res or not (1 <= index <= len(res[name])):
            return name
        return res[name][index - 1]


def generate_bond_hamiltonian(bond: MagnetBondParams, sites, deterministic_half: bool = False):
    """magnet.jl:133-165 -> ((dimi, dimj), H, energy_offset_factor)."""
    dimi, dimj = sites[0].spin_states, sites[1].spin_states
    splusi, szi = spin_operators(dimi)
    splusj, szj = spin_operators(dimj)
    idi, idj = np.eye(dimi), np.eye(dimj)
    H = (
        bond.J * (0.5 * (np.kron(splusi.T, splusj) + np.kron(splusi, splusj.T)) + (1 + bond.d) * np.kron(szi, szj))
        + bond.hz[0] * np.kron(idi, szj)
        + bond.hz[1] * np.kron(szi, idj)
        + bond.Dx[0] / 4 * np.kron((splusi + splusi.T) @ (splusi + splusi.T), idj)
        + bond.Dx[1] / 4 * np.kron(idi, (splusj + splusj.T) @ (splusj + splusj.T))
        + bond.Dz[0] * np.kron(szi @ szi, idj)
        + bond.Dz[1] * np.kron(idi, szj @ szj)
    )
    energy_offset_factor = 0.25
    # magnet.jl:159-162 intends "use the deterministic solution for S == 1//2", but its test
    # `bond.hz == 0 && bond.Dz == 0 && bond.Dx == 0` compares TUPLES with the integer 0, which is
    # `false` in Julia (generic `==` falls back to `===`), so the reference always keeps 0.25.
    # Default = that literal behaviour; `deterministic_half=True` (task parameter
    # `s_half_deterministic`) selects the intended energy_offset_factor = 0 tables instead.
    if deterministic_half and dimi == 2 and dimj == 2 and _iszero(bond.hz) and _iszero(bond.Dz) and _iszero(bond.Dx):
        energy_offset_factor = 0.0
    return (dimi, dimj), H, energy_offset_factor


def _iszero(t) -> bool:
    return all(x == 0 for x in t)


class MagnetModel:
    """magnet.jl:56-120"""

    LEG_COUNT = 4

    def __init__(self, params: dict):
        self.lattice = Lattice(params["lattice"])
        assert len(self.lattice.bonds) > 0
        pm = ParameterMap(params.get("parameter_map"))
        lat = self.lattice

        def split_site(param, bond, default):
            iuc, _ = lat.split_idx(bond.i)
            juc, _ = lat.split_idx(bond.j)
            first = params.get(pm.get(param, iuc), default) / lat.uc.sites[iuc - 1].coordination
            second = params.get(pm.get(param, juc), default) / lat.uc.sites[juc - 1].coordination
            return (float(first), float(second))

        self.bond_params = [
            MagnetBondParams(
                float(params[pm.get("J", bond.type)]),
                float(params.get(pm.get("d", bond.type), 0.0)),
                split_site("Dx", bond, 0.0),
                split_site("Dz", bond, 0.0),
                split_site("hz", bond, 0.0),
            )
            for bond in lat.bonds
        ]
        uc_site_params = [
            MagnetSiteParams(int(Fraction(params.get(pm.get("S", i), Fraction(1, 2))) * 2 + 1))
            for i in range(1, len(lat.uc.sites) + 1)
        ]
        self.site_params = uc_site_params * int(np.prod(lat.Ls))
        self.opstring_estimators = gen_opstring_estimators(lat, params)
        self.deterministic_half = bool(params.get("s_half_deterministic", False))

    # --- AbstractModel interface ---------------------------------------------------------------
    @classmethod
    def leg_count(cls) -> int:
        return cls.LEG_COUNT

    def normalization_site_count(self) -> int:
        return self.lattice.site_count()

    def get_opstring_estimators(self):
        return self.opstring_estimators

    def generate_sse_data(self) -> SSEData:
        """magnet.jl:169-184 (zips unit-cell bonds with the FIRST len(uc.bonds) bond params)."""
        lat = self.lattice
        vertex_data = []
        for uc_bond, bond in zip(lat.uc.bonds, self.bond_params):
            dims, H, eof = generate_bond_hamiltonian(
                bond, (self.site_params[uc_bond.iuc - 1], self.site_params[uc_bond.juc - 1]), self.deterministic_half
            )
            vertex_data.append(make_vertex_data(dims, H, energy_offset_factor=eof))
        bonds = [SSEBond(b.type, (b.i, b.j)) for b in lat.bonds]
        return SSEData(vertex_data, bonds)

    # --- MagnetizationEstimator plug points ----------------------------------------------------
    def magnetization_state(self, tag, site_idx: int, state_idx: int) -> float:
        """magnet.jl:122-129"""
        return (self.site_params[site_idx - 1].spin_states - 1) * 0.5 - state_idx + 1

    def magnetization_lattice_site_idx(self, sse_site_idx: int):
        return sse_site_idx

    def staggered_sign(self, q, stagger_uc, site_idx: int) -> int:
        return self.lattice.staggered_sign(q, stagger_uc, site_idx)

    def site_dim(self, sse_site_idx: int) -> int:
        return self.site_params[sse_site_idx - 1].spin_states


def gen_opstring_estimators(lattice: Lattice, params: dict):
    """magnet.jl:186-209"""
    ests = []
    for est in params.get("measure", []):
        if est == "magnetization":
            q = tuple(False for _ in range(lattice.dimension))
            ests.append(MagnetizationEstimator(q, False, "", None))
        elif est == "staggered_magnetization":
            neel = neel_vector(lattice.uc)
            if neel is None:
                raise ValueError(
                    "selected :staggered_magnetization measurement, but lattice does not have a Neel vector."
                )
            q, stagger_uc = neel
            ests.append(MagnetizationEstimator(q, stagger_uc, "Stag", None))
        elif isinstance(est, MagnetizationEstimator):
            ests.append(est)
        else:
            raise ValueError(f"Unrecognized measure option '{est}'")
    return ests
