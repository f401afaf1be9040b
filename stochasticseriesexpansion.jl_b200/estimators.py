"""Operator-string estimators: the table-driven `MagnetizationEstimator`.

Mirror of /root/reference/src/abstract_opstring_estimator.jl:30-69 and
/root/reference/src/models/common/magnetization_estimator.jl:32-258.  On the device an estimator is
a dense table `value[site, state] = staggered_sign(site) * magnetization_state(site, state)`; the
per-operator accumulation (`init`/`measure`/`result`, :96-230) runs in the measurement kernel.
Arbitrary user `measure` callbacks cannot run on the device: only estimators that
can be expressed as such a table are supported.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

OBS_NAMES = ("Mag", "AbsMag", "Mag2", "Mag4", "MagChi", "BinderRatio")


@dataclass(frozen=True)
class MagnetizationEstimator:
    """Type parameters of `MagnetizationEstimator{OrderingVector,StaggerUC,Model,Prefix,Tag}`."""

    ordering_vector: tuple
    stagger_uc: bool
    prefix: str
    tag: object = None

    def obs_symbols(self):
        """magnetization_estimator_obs_symbols (:188-202): (plain names, sign-multiplied names)."""
        return (
            {n.lower(): self.prefix + n for n in OBS_NAMES},
            {n.lower(): "Sign" + self.prefix + n for n in OBS_NAMES},
        )

    def value_table(self, model, n_sites: int, max_dim: int) -> np.ndarray:
        """[n_sites, max_dim] f64: sign(site) * m(site, state) for 1-based state = column+1."""
        tab = np.zeros((n_sites, max_dim), dtype=np.float64)
        for site in range(1, n_sites + 1):
            lsite = model.magnetization_lattice_site_idx(site)
            if lsite is None:
                continue
            sg = model.staggered_sign(self.ordering_vector, self.stagger_uc, lsite)
            for state in range(1, model.site_dim(site) + 1):
                tab[site - 1, state - 1] = sg * model.magnetization_state(self.tag, lsite, state)
        return tab


def magnetization_estimator_standard_prefix(q, stagger_uc: bool) -> str:
    """:172-180"""
    if not any(q) and not stagger_uc:
        return ""
    names = "".join(chr((ord("X") - ord("A") + i) % 26 + ord("A")) for i, qi in enumerate(q) if qi)
    return "Stag" + names + ("uc" if stagger_uc else "")


def all_magnetization_estimators(dimension: int, tag=None):
    """:55-70 — array comprehension: stagger_uc fastest, then q (first factor fastest)."""
    out = []
    for q_rev in itertools.product((False, True), repeat=dimension):
        q = tuple(reversed(q_rev))
        for stagger_uc in (False, True):
            out.append(
                MagnetizationEstimator(q, stagger_uc, magnetization_estimator_standard_prefix(q, stagger_uc), tag)
            )
    return out


def register_evaluables(est: MagnetizationEstimator, evaluator) -> None:
    """:236-258"""
    symbols, signsymbols = est.obs_symbols()
    for obs in ("mag", "absmag", "mag2", "mag4", "magchi"):
        evaluator.evaluate(symbols[obs], lambda so, s: so / s, (signsymbols[obs], "Sign"))

    def binder(smag4, smag2, sign):
        if smag2 == 0 and smag4 == 0:
            return 0.0
        return smag2 ** 2 / smag4 / sign

    evaluator.evaluate(symbols["binderratio"], binder, (signsymbols["mag4"], signsymbols["mag2"], "Sign"))
