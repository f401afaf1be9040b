"""Replica exchange (parallel tempering) over the walkers of a batch: decided on the host (ReplicaExchange) or on the
device (DeviceReplicaExchange, sse_pt_exchange).

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

    def step(self, allow_open_bin: bool = False):
        if not allow_open_bin:  # the accumulators are indexed by walker: a bin must not span a change of temperature
            _, counts = self.walkers.fetch_accumulators(reset=False)
            if counts.any():
                raise RuntimeError("a bin is open: flush the accumulators before exchanging temperatures")
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


def pt_uniforms(seed: int, step: int, n: int) -> np.ndarray:
    """The uniforms sse_pt_exchange(seed, step) uses for its pairs 0..n-1 (draw i of the Philox stream (seed, step))."""
    import ctypes as C

    from . import capi

    out = np.zeros(max(n, 1))
    capi.check(capi.lib().sse_pt_uniforms(int(seed), int(step), int(n), out.ctypes.data_as(capi.f64p)))
    return out[:n]


class DeviceReplicaExchange:
    """Neighbour swaps decided on the device (sse_pt_set_ladder / sse_pt_exchange): per exchange step one small kernel and a
    4-byte read-back instead of a round trip of operator counts and temperatures.  Measurements belong to a TEMPERATURE,
    not to a walker: `rank_of_walker()` gives the group index to pass to Walkers.reduce_bins, and bins must be flushed
    before every exchange step (step() refuses to run on a non-empty bin unless told otherwise)."""

    def __init__(self, walkers, seed: int = 0):
        self.walkers = walkers
        self.seed = int(seed)
        self.step_index = 0
        self.parity = 0
        self.proposed = 0
        self.accepted = 0
        order = np.argsort(np.asarray(walkers.T, dtype=np.float64), kind="stable")
        walkers.pt_set_ladder(order)

    def rank_of_walker(self) -> np.ndarray:
        ladder = self.walkers.pt_get_ladder()
        rank = np.empty(len(ladder), dtype=np.int32)
        rank[ladder] = np.arange(len(ladder), dtype=np.int32)
        return rank

    def step(self, allow_open_bin: bool = False) -> int:
        if not allow_open_bin:
            _, counts = self.walkers.fetch_accumulators(reset=False)
            if counts.any():
                raise RuntimeError("a bin is open: flush the accumulators (reduce_bins / fetch_accumulators) before exchanging "
                                   "temperatures, or a bin mixes measurements taken at different temperatures")
        n = self.walkers._n_ladder
        pairs = max(0, (n - self.parity) // 2)
        acc = self.walkers.pt_exchange(self.parity, self.seed, self.step_index)
        self.proposed += pairs
        self.accepted += acc
        self.parity ^= 1
        self.step_index += 1
        return acc
