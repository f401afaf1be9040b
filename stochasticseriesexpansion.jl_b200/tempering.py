"""Replica exchange (parallel tempering) over the walkers of a batch — host-side decisions only.

The reference exposes two hooks to Carlo's parallel-tempering wrapper (src/sse.jl:390-405):
`parallel_tempering_log_weight_ratio(mc, :T, T_new) = -n * log(T_new / T)` and
`parallel_tempering_change_parameter!(mc, :T, T_new)`.  With a batch of walkers on one GPU the exchange needs no
Carlo/MPI machinery: the operator counts come back with one `sse_get_num_operators`, the swap decisions are made
here with the reference's weight ratio, and the new temperatures go down with one `sse_set_temperature`.
Configurations never move; only the temperature labels do (as in Carlo)."""
from __future__ import annotations

import numpy as np


def log_weight_ratio(n, T_old, T_new):
    """src/sse.jl:395, vectorised."""
    return -np.asarray(n, dtype=np.float64) * np.log(np.asarray(T_new, dtype=np.float64) / np.asarray(T_old, dtype=np.float64))


def swap_decisions(n, T, order, parity: int, uniforms):
    """One sweep of neighbour swaps.  `order` lists walker indices sorted by temperature; pairs
    (order[i], order[i+1]) with i % 2 == parity are proposed.  Accept with min(1, exp(lw_a + lw_b)), where
    lw_x is walker x's log weight ratio for taking the other's temperature.  Returns the new temperature array."""
    n = np.asarray(n, dtype=np.float64)
    T = np.array(T, dtype=np.float64)
    order = np.asarray(order)
    k = 0
    for i in range(parity, len(order) - 1, 2):
        a, b = order[i], order[i + 1]
        lw = log_weight_ratio(n[a], T[a], T[b]) + log_weight_ratio(n[b], T[b], T[a])
        if np.log(max(float(uniforms[k]), 1e-300)) < lw:
            T[a], T[b] = T[b], T[a]
        k += 1
    return T


class ReplicaExchange:
    """Drives neighbour swaps for a `Walkers` batch whose walkers sit on a temperature ladder."""

    def __init__(self, walkers, seed: int = 0):
        self.walkers = walkers
        self.rng = np.random.default_rng(seed)
        self.parity = 0
        self.proposed = 0
        self.accepted = 0

    def step(self):
        n = self.walkers.num_operators()
        T = np.array(self.walkers.T, dtype=np.float64)
        order = np.argsort(T, kind="stable")
        npairs = max(0, (len(order) - self.parity) // 2)
        u = self.rng.random(max(npairs, 1))
        T_new = swap_decisions(n, T, order, self.parity, u)
        self.proposed += npairs
        self.accepted += int(np.count_nonzero(T_new != T) // 2)
        self.parity ^= 1
        self.walkers.set_temperature(T_new)
        return T_new
