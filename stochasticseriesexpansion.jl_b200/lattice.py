"""Unit cells and lattices: host geometry consumed as a flat bond list + per-site sign tables.

Mirror of /root/reference/src/models/common/lattice.jl (UCBond/UCSite/UnitCell :17-131, neel_vector
:134-150, Lattice :223-243, staggered_sign :257-267, UnitCells :280-331).  Site and bond indices are
1-based here exactly as in the reference; the flattening to 0-based happens once in `sse_data.flatten`.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from .util import join_idx, split_idx


@dataclass(frozen=True)
class UCBond:
    iuc: int
    jd: tuple
    juc: int


@dataclass
class UCSite:
    pos: tuple
    sublattice_sign: int = 0
    coordination: int = 0


def calculate_uc_signs(bonds, num_sites: int):
    """lattice.jl:62-93 (bipartite sublattice signs from intra-cell bonds)."""
    signs = [0] * num_sites
    signs[0] = 1
    tries = 0
    while any(s == 0 for s in signs) or tries > len(bonds) ** 2:
        progressed = False
        for b in bonds:
            if any(x != 0 for x in b.jd):
                continue
            i, j = b.iuc - 1, b.juc - 1
            if signs[i] != 0 and signs[j] == 0:
                signs[j] = -signs[i]
                progressed = True
            elif signs[j] != 0 and signs[i] == 0:
                signs[i] = -signs[j]
                progressed = True
            elif signs[i] == signs[j]:
                signs = [1] * num_sites  # lattice not bipartite
                break
            tries += 1
        if not progressed:
            # the reference would spin here when a cell site is only reachable through inter-cell
            # bonds; it resolves to all-ones below, so break out explicitly
            break
    if any(s == 0 for s in signs):
        signs = [1] * num_sites
    return signs


def calculate_uc_coordinations(bonds, num_sites: int):
    c = [0] * num_sites
    for b in bonds:
        c[b.iuc - 1] += 1
        c[b.juc - 1] += 1
    return c


class UnitCell:
    def __init__(self, lattice_vectors, sites, bonds):
        self.lattice_vectors = np.asarray(lattice_vectors, dtype=np.float64)
        signs = calculate_uc_signs(bonds, len(sites))
        coords = calculate_uc_coordinations(bonds, len(sites))
        self.sites = [UCSite(tuple(s.pos), sg, c) for s, sg, c in zip(sites, signs, coords)]
        self.bonds = list(bonds)

    @property
    def dimension(self) -> int:
        return len(self.bonds[0].jd)


def neel_vector(uc: UnitCell):
    """lattice.jl:134-150 -> (q tuple of bools, stagger_uc) or None."""
    D = uc.dimension
    for stagger_uc in (False, True):
        # Iterators.product: first factor fastest
        for q_rev in itertools.product((False, True), repeat=D):
            q = tuple(reversed(q_rev))
            ok = True
            for bond in uc.bonds:
                si = uc.sites[bond.iuc - 1].sublattice_sign ** int(stagger_uc)
                sj = uc.sites[bond.juc - 1].sublattice_sign ** int(stagger_uc) * (-1) ** (
                    sum(int(a) * int(b) for a, b in zip(bond.jd, q))
                )
                if si == sj:
                    ok = False
                    break
            if ok:
                return q, stagger_uc
    return None


@dataclass(frozen=True)
class LatticeBond:
    type: int
    i: int
    j: int


@dataclass(frozen=True)
class LatticeSite:
    iuc: int
    ix: tuple


class Lattice:
    """lattice.jl:204-243.  `Lattice(uc, Ls)` or `Lattice(params)` with keys unitcell/size."""

    def __init__(self, uc, Ls=None):
        if Ls is None:
            p = uc
            uc = p["unitcell"] if isinstance(p, dict) else p.unitcell
            Ls = p["size"] if isinstance(p, dict) else p.size
        self.uc = uc
        self.Ls = tuple(int(L) for L in Ls)
        dims = (len(uc.sites),) + self.Ls
        self.bonds = []
        self.sites = []
        # Iterators.product([1:L for L in Ls]...): first dimension fastest
        for r_rev in itertools.product(*[range(1, L + 1) for L in reversed(self.Ls)]):
            r = tuple(reversed(r_rev))
            for bond_type, b in enumerate(uc.bonds, start=1):
                i = join_idx(dims, (b.iuc,) + r)
                rj = tuple((x + d - 1) % L + 1 for x, d, L in zip(r, b.jd, self.Ls))
                j = join_idx(dims, (b.juc,) + rj)
                self.bonds.append(LatticeBond(bond_type, i, j))
            for iuc in range(1, len(uc.sites) + 1):
                self.sites.append(LatticeSite(iuc, r))

    @property
    def dimension(self) -> int:
        return len(self.Ls)

    def split_idx(self, site_idx: int):
        r = split_idx((len(self.uc.sites),) + self.Ls, site_idx)
        return r[0], r[1:]

    def site_count(self) -> int:
        return len(self.uc.sites) * int(np.prod(self.Ls))

    def staggered_sign(self, ordering_vector, stagger_uc: bool, site_idx: int) -> int:
        """lattice.jl:257-267"""
        s = self.sites[site_idx - 1]
        sign = self.uc.sites[s.iuc - 1].sublattice_sign if stagger_uc else 1
        sign *= 1 - 2 * (sum(int(q) * x for q, x in zip(ordering_vector, s.ix)) % 2)
        return sign


class UnitCells:
    """Predefined unit cells (lattice.jl:280-331) plus a 1-D chain for BASELINE config 0."""

    square = UnitCell(
        [[1.0, 0.0], [0.0, 1.0]],
        [UCSite((0.0, 0.0))],
        [UCBond(1, (0, 1), 1), UCBond(1, (1, 0), 1)],
    )
    columnar_dimer = UnitCell(
        [[1.0, 0.0], [0.0, 2.0]],
        [UCSite((0.0, 0.0)), UCSite((0.0, 0.5))],
        [UCBond(1, (0, 0), 2), UCBond(1, (1, 0), 1), UCBond(2, (1, 0), 2), UCBond(2, (0, 1), 1)],
    )
    honeycomb = UnitCell(
        [[np.sqrt(3) / 2, np.sqrt(3 / 2)], [-0.5, 0.5]],
        [UCSite((0.0, 0.0)), UCSite((1 / 3, 1 / 3))],
        [UCBond(1, (0, 0), 2), UCBond(2, (0, 1), 1), UCBond(2, (1, 0), 1)],
    )
    triangle = UnitCell(
        [[1.0, -0.5], [0.0, np.sqrt(3 / 2)]],
        [UCSite((0.0, 0.0))],
        [UCBond(1, (0, 1), 1), UCBond(1, (1, 0), 1), UCBond(1, (1, 1), 1)],
    )
    fully_frust_square_bilayer = UnitCell(
        [[1.0, 0.0], [0.0, 1.0]],
        [UCSite((0.0, 0.0)), UCSite((0.0, 0.0))],
        [
            UCBond(1, (0, 0), 2),
            UCBond(1, (0, 1), 1),
            UCBond(2, (0, 1), 2),
            UCBond(1, (1, 0), 1),
            UCBond(2, (1, 0), 2),
            UCBond(1, (0, 1), 2),
            UCBond(2, (0, 1), 1),
            UCBond(1, (1, 0), 2),
            UCBond(2, (1, 0), 1),
        ],
    )
    # not in the reference ("no predefined chain exists"): 1-D chain, one bond per cell
    chain = UnitCell([[1.0]], [UCSite((0.0,))], [UCBond(1, (1,), 1)])
