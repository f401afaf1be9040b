"""`ClusterModel`: a MagnetModel rewritten in a multi-spin cluster basis (e.g. the dimer basis).

Mirror of /root/reference/src/models/cluster/cluster.jl (:1-323).  Host-side model setup only: it
produces the dim-4x4 bond tables BASELINE config 5 (bilayer Heisenberg in the dimer basis) needs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .estimators import MagnetizationEstimator
from .magnet import MagnetModel, ParameterMap, generate_bond_hamiltonian
from .sse_data import SSEBond, SSEData
from .vertex_data import make_vertex_data


@dataclass
class ClusterBasis:
    """cluster.jl:8-11: quantum numbers per basis state + unitary with the basis states as columns."""

    quantum_numbers: list
    transformation: np.ndarray


class ClusterBases:
    _s = 1 / np.sqrt(2)
    dimer = ClusterBasis(
        [(0.0, 0.0), (1.0, 1.0), (1.0, 0.0), (1.0, -1.0)],
        np.array([[0, 1, 0, 0], [_s, 0, _s, 0], [-_s, 0, _s, 0], [0, 0, 0, 1]], dtype=np.float64),
    )


def lift_twobody_operator(op: np.ndarray, site_dims, sites):
    """Embed a two-site operator (kron convention: leftmost factor slowest) at 1-based positions
    `sites` of a product space with local dimensions `site_dims` (cluster.jl:93-126)."""
    site_dims = tuple(int(d) for d in site_dims)
    n = len(site_dims)
    a, b = sites[0] - 1, sites[1] - 1
    da, db = site_dims[a], site_dims[b]
    op4 = np.asarray(op).reshape(da, db, da, db)
    total = int(np.prod(site_dims))
    res = np.zeros(site_dims + site_dims, dtype=op4.dtype)
    for I in np.ndindex(*site_dims):
        for ka in range(da):
            for kb in range(db):
                K = list(I)
                K[a] = ka
                K[b] = kb
                res[I + tuple(K)] = op4[I[a], I[b], ka, kb]
    return res.reshape(total, total)


def build_cluster_hamiltonians(cluster_ids, uc_bonds, bonds, site_params):
    """cluster.jl:128-199 -> (intracluster list, intercluster dict keyed by (iuc, juc, jd))."""
    uniq = list(dict.fromkeys(cluster_ids))
    clusters = [[i for i, c in enumerate(cluster_ids) if c == cid] for cid in uniq]
    cluster_dims = [tuple(site_params[s].spin_states for s in cl) for cl in clusters]
    ordering = [0] * len(cluster_ids)
    for cl in clusters:
        for i, s in enumerate(cl, start=1):
            ordering[s] = i
    intra = [np.zeros((int(np.prod(d)), int(np.prod(d)))) for d in cluster_dims]
    inter: dict = {}
    for uc_bond, bond in zip(uc_bonds, bonds):
        (_, _), H, _ = generate_bond_hamiltonian(
            bond, (site_params[uc_bond.iuc - 1], site_params[uc_bond.juc - 1])
        )
        ci, cj = cluster_ids[uc_bond.iuc - 1], cluster_ids[uc_bond.juc - 1]
        if ci == cj and all(x == 0 for x in uc_bond.jd):
            if uc_bond.iuc == uc_bond.juc:
                raise ValueError("found bond in Magnet connecting site to itself... not supported by ClusterModel")
            intra[ci - 1] += lift_twobody_operator(
                H, cluster_dims[ci - 1], (ordering[uc_bond.iuc - 1], ordering[uc_bond.juc - 1])
            )
        else:
            key = (ci, cj, tuple(uc_bond.jd))
            dims = cluster_dims[ci - 1] + cluster_dims[cj - 1]
            dim = int(np.prod(dims))
            Hij = inter.setdefault(key, np.zeros((dim, dim)))
            Hij += lift_twobody_operator(
                H, dims, (ordering[uc_bond.iuc - 1], len(cluster_dims[ci - 1]) + ordering[uc_bond.juc - 1])
            )
    return intra, inter


def absorb_intracluster_hamiltonians(intra, inter):
    """cluster.jl:201-222"""
    coord = [sum((k[0] == i) + (k[1] == i) for k in inter) for i in range(1, len(intra) + 1)]
    out = {}
    for key, Hij in inter.items():
        i, j = key[0] - 1, key[1] - 1
        out[key] = (
            Hij
            + np.kron(intra[i], np.eye(intra[j].shape[0]) / coord[i])
            + np.kron(np.eye(intra[i].shape[0]) / coord[j], intra[j])
        )
    return out


class ClusterModel:
    """cluster.jl:25-62.  Parameters: `inner_model` (class, default MagnetModel), `cluster_bases`,
    `measure_quantum_numbers` = list of dicts {name, quantum_number (1-based)}, optional `cluster_id`."""

    LEG_COUNT = 4

    def __init__(self, params: dict):
        inner_cls = params.get("inner_model", MagnetModel)
        self.inner_model = inner_cls(params)
        self.basis = tuple(params["cluster_bases"])
        pm = ParameterMap(params.get("parameter_map"))
        lat = self.inner_model.lattice
        uc_ids = [params.get(pm.get("cluster_id", i), 1) for i in range(1, len(lat.uc.sites) + 1)]
        if len(set(uc_ids)) != len(self.basis):
            raise ValueError(
                f"Number of cluster bases ({len(self.basis)}) does not match number of distinct cluster ids ({len(set(uc_ids))})."
            )
        self.uc_cluster_ids = uc_ids
        self.cluster_ids = uc_ids * int(np.prod(lat.Ls))
        q = tuple(False for _ in range(lat.dimension))
        self.opstring_estimators = [
            MagnetizationEstimator(q, False, str(m["name"]), int(m["quantum_number"]))
            for m in params["measure_quantum_numbers"]
        ]
        self._sse_data = None

    @classmethod
    def leg_count(cls) -> int:
        return cls.LEG_COUNT

    def normalization_site_count(self) -> int:
        return self.inner_model.normalization_site_count()

    def get_opstring_estimators(self):
        return self.opstring_estimators

    def magnetization_state(self, tag, site_idx: int, state_idx: int) -> float:
        """cluster.jl:66-73 (tag = 1-based quantum-number index)."""
        return float(self.basis[self.cluster_ids[site_idx - 1] - 1].quantum_numbers[state_idx - 1][tag - 1])

    def magnetization_lattice_site_idx(self, sse_site_idx: int):
        return sse_site_idx

    def staggered_sign(self, q, stagger_uc, site_idx: int) -> int:
        return 1  # cluster.jl:75

    def site_dim(self, sse_site_idx: int) -> int:
        return self.generate_sse_data().sites[sse_site_idx - 1].dim

    def generate_sse_data(self) -> SSEData:
        """generate_cluster_sse_data (cluster.jl:245-306)."""
        if self._sse_data is not None:
            return self._sse_data
        mag = self.inner_model
        lat = mag.lattice
        n_uc = len(lat.uc.sites)
        intra, inter = build_cluster_hamiltonians(self.uc_cluster_ids, lat.uc.bonds, mag.bond_params, mag.site_params)
        num_clusters = len(intra)
        keys = list(inter.keys())
        bonds = []
        seen = set()
        for bond in lat.bonds:
            uc_bond = lat.uc.bonds[bond.type - 1]
            iuc = self.uc_cluster_ids[uc_bond.iuc - 1]
            juc = self.uc_cluster_ids[uc_bond.juc - 1]
            i = num_clusters * ((bond.i - 1) // n_uc) + iuc
            j = num_clusters * ((bond.j - 1) // n_uc) + juc
            if iuc == juc and all(x == 0 for x in uc_bond.jd):
                continue
            bond_type = keys.index((iuc, juc, tuple(uc_bond.jd))) + 1
            b = SSEBond(bond_type, (i, j))
            if b not in seen:
                seen.add(b)
                bonds.append(b)
        absorbed = absorb_intracluster_hamiltonians(intra, inter)
        vertex_data = []
        for key in keys:
            H = absorbed[key]
            U = np.kron(self.basis[key[0] - 1].transformation, self.basis[key[1] - 1].transformation)
            Hc = U.T @ H @ U
            Hc = np.where(np.abs(Hc) < 1e-14, 0.0, Hc)
            vertex_data.append(
                make_vertex_data((intra[key[0] - 1].shape[0], intra[key[1] - 1].shape[0]), Hc, energy_offset_factor=0.25)
            )
        self._sse_data = SSEData(vertex_data, bonds)
        return self._sse_data
