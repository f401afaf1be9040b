"""`SSEData`: the read-only input of the sweep, and its flattening to the C-ABI table format.

Mirror of /root/reference/src/sse_data.jl:1-71.  `flatten()` produces the POD arrays of
`include/sse_b200.h:sse_model_desc`: 0-based sites/bonds/types, vertex
indices made local-1-based per type (0 = invalid), one flat outcome list for all types.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .util import INVALID_VERTEX_CODE
from .vertex_data import VertexData


@dataclass(frozen=True)
class SSESite:
    dim: int


@dataclass(frozen=True)
class SSEBond:
    """SSEBond(type, (i, j, ...)) — 1-based type and site indices (sse_data.jl:9-13)."""

    type: int
    sites: tuple


def generate_sites_from_bonds(vertex_data, bonds):
    """sse_data.jl:44-63"""
    max_site = max(max(b.sites) for b in bonds)
    dims = [0] * max_site
    for b in bonds:
        for is_, s in enumerate(b.sites):
            new_dim = vertex_data[b.type - 1].dims[is_]
            if dims[s - 1] == 0:
                dims[s - 1] = new_dim
            elif dims[s - 1] != new_dim:
                raise ValueError(
                    "SSEData: site dimensions set by VertexData are inconsistent!\n"
                    f"Site {s}: {dims[s - 1]}, Bond {b.type}: {new_dim}"
                )
    return [SSESite(d) for d in dims]


class SSEData:
    """SSEData(vertex_data, bonds) (sse_data.jl:34-42)."""

    def __init__(self, vertex_data, bonds):
        self.vertex_data: list[VertexData] = list(vertex_data)
        self.bonds: list[SSEBond] = list(bonds)
        self.energy_offset = float(sum(self.vertex_data[b.type - 1].energy_offset for b in self.bonds))
        self.sites = generate_sites_from_bonds(self.vertex_data, self.bonds)
        self.nsites_per_bond = len(self.bonds[0].sites)

    def get_vertex_data(self, bond_idx: int) -> VertexData:
        return self.vertex_data[self.bonds[bond_idx - 1].type - 1]

    # ------------------------------------------------------------------------------------------
    def flatten(self) -> dict:
        """POD arrays for `sse_model_desc`.  Everything 0-based except states (1-based, as stored in
        `leg_states`) and worms (1-based); local vertex index 0 means "invalid"."""
        ns = self.nsites_per_bond
        if ns != 2:
            raise NotImplementedError("the B200 sweep backend supports 2-site bonds (leg_count == 4) only")
        nl = 2 * ns
        n_types = len(self.vertex_data)
        max_worm = max(max(d - 1 for d in vd.dims) for vd in self.vertex_data)
        max_worm = max(max_worm, 1)

        site_dim = np.array([s.dim for s in self.sites], dtype=np.uint8)
        bond_type = np.array([b.type - 1 for b in self.bonds], dtype=np.int32)
        bond_sites = np.array([[s - 1 for s in b.sites] for b in self.bonds], dtype=np.int32).reshape(-1)

        type_dims = np.array([vd.dims for vd in self.vertex_data], dtype=np.int32).reshape(-1)
        voff = np.zeros(n_types + 1, dtype=np.int32)
        doff = np.zeros(n_types + 1, dtype=np.int32)
        for t, vd in enumerate(self.vertex_data):
            voff[t + 1] = voff[t] + vd.vertex_count()
            doff[t + 1] = doff[t] + len(vd.diagonal_vertices)
        nv = int(voff[-1])

        weights = np.concatenate([vd.weights for vd in self.vertex_data]).astype(np.float64)
        signs = np.concatenate([vd.signs for vd in self.vertex_data]).astype(np.int8)
        leg_states = np.concatenate([vd.leg_states.T.reshape(-1) for vd in self.vertex_data]).astype(np.uint8)
        is_diag = np.zeros(nv, dtype=np.uint8)
        diag_vertices = np.zeros(int(doff[-1]), dtype=np.int32)
        trans_offset = -np.ones(nv * max_worm * nl, dtype=np.int32)
        trans_count = np.zeros(nv * max_worm * nl, dtype=np.int32)
        out_cumprob, out_target, out_leg, out_worm = [], [], [], []
        for t, vd in enumerate(self.vertex_data):
            ls = vd.leg_states
            is_diag[voff[t]:voff[t + 1]] = np.all(ls[:ns] == ls[ns:], axis=0)
            for c, code in enumerate(vd.diagonal_vertices):
                diag_vertices[doff[t] + c] = 0 if code >= INVALID_VERTEX_CODE - 1 else (int(code) >> 1)
            base = len(out_cumprob)
            out_cumprob.extend(vd.transition_cumprobs.tolist())
            out_target.extend((vd.transition_targets >> 1).tolist())
            out_leg.extend((vd.transition_step_outs[:, 0] - 1).tolist() if len(vd.transition_step_outs) else [])
            out_worm.extend(vd.transition_step_outs[:, 1].tolist() if len(vd.transition_step_outs) else [])
            nleg, nworm, nvert = vd.trans_offset.shape
            for v in range(nvert):
                for w in range(nworm):
                    for leg in range(nleg):
                        off = int(vd.trans_offset[leg, w, v])
                        if off < 0:
                            continue
                        idx = ((voff[t] + v) * max_worm + w) * nl + leg
                        trans_offset[idx] = base + off - 1
                        trans_count[idx] = int(vd.trans_length[leg, w, v]) + 1

        return dict(
            n_sites=len(self.sites),
            site_dim=site_dim,
            n_bonds=len(self.bonds),
            bond_type=bond_type,
            bond_sites=bond_sites,
            n_types=n_types,
            type_dims=type_dims,
            type_vertex_off=voff,
            type_diag_off=doff,
            n_vertices=nv,
            weights=weights,
            signs=signs,
            leg_states=leg_states,
            is_diag=is_diag,
            diag_vertices=diag_vertices,
            max_worm=int(max_worm),
            trans_offset=trans_offset,
            trans_count=trans_count,
            n_outcomes=len(out_cumprob),
            out_cumprob=np.array(out_cumprob, dtype=np.float64),
            out_target=np.array(out_target, dtype=np.int32),
            out_leg=np.array(out_leg, dtype=np.int32),
            out_worm=np.array(out_worm, dtype=np.int32),
            energy_offset=float(self.energy_offset),
        )
