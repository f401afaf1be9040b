"""Minimal stand-in for the parts of Carlo.jl the reference's `MC` talks to.

Carlo.jl itself (scheduler, MPI, HDF5 checkpoints, CLI) is out of scope; what the hot path needs from it
is tiny: `MCContext` with `is_thermalized` / `measure!` and binned accumulators, the `sweep!`/`measure!`
run loop, and the jackknife `evaluate!` used by `register_evaluables` (src/sse.jl:111-134).  Observables
here are VECTORS over the walkers of a batch (one entry per walker), which Carlo supports natively.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


class MCContext:
    """ctx.sweeps / is_thermalized(ctx) / measure!(ctx, name, value) with `binsize` internal bins."""

    def __init__(self, params: dict):
        self.sweeps = 0
        self.thermalization_sweeps = int(params.get("thermalization", 0))
        self.binsize = int(params.get("binsize", 1))
        self._acc: dict = OrderedDict()  # name -> [sum, count]
        self.bins: dict = OrderedDict()  # name -> list of bin means (arrays over walkers)

    def is_thermalized(self) -> bool:
        return self.sweeps > self.thermalization_sweeps

    def measure(self, name: str, value):
        v = np.asarray(value, dtype=np.float64)
        a = self._acc.get(name)
        if a is None:
            a = self._acc[name] = [np.zeros_like(v), 0]
        a[0] = a[0] + v
        a[1] += 1
        if a[1] >= self.binsize:
            self.bins.setdefault(name, []).append(a[0] / a[1])
            self._acc[name] = [np.zeros_like(v), 0]

    def add_bin(self, name: str, mean):
        """Append one externally accumulated bin (device-side accumulation path)."""
        self.bins.setdefault(name, []).append(np.asarray(mean, dtype=np.float64))

    def bin_array(self, name: str) -> np.ndarray:
        return np.array(self.bins[name])  # [n_bins, ...]


class Evaluator:
    """`evaluate!(f, eval, :Name, (:Inputs...))` with jackknife error propagation over bins.

    `bins[name]` is an array [n_samples] of independent bin means."""

    def __init__(self, bins: dict):
        self.bins = {k: np.asarray(v, dtype=np.float64) for k, v in bins.items()}
        self.results: dict = OrderedDict()
        for k, v in self.bins.items():
            n = len(v)
            self.results[k] = (float(v.mean()), float(v.std(ddof=1) / np.sqrt(n)) if n > 1 else float("nan"))

    def evaluate(self, name: str, func, inputs):
        if any(i not in self.bins for i in inputs):
            return
        xs = [self.bins[i] for i in inputs]
        n = min(len(x) for x in xs)
        xs = [x[:n] for x in xs]
        sums = [x.sum() for x in xs]
        full = func(*[s / n for s in sums])
        if n < 2:
            self.results[name] = (float(full), float("nan"))
            return
        jk = np.array([func(*[(s - x[j]) / (n - 1) for s, x in zip(sums, xs)]) for j in range(n)])
        mean_jk = jk.mean()
        err = np.sqrt((n - 1) / n * np.sum((jk - mean_jk) ** 2))
        bias_corrected = n * full - (n - 1) * mean_jk
        self.results[name] = (float(bias_corrected), float(err))

    def __getitem__(self, name):
        return self.results[name]


def run(mc, params: dict, ctx: MCContext | None = None, fused: bool = True) -> MCContext:
    """Carlo's run loop for one task: init!, then `sweeps` x (sweep!; measure! once thermalised).

    fused=True keeps the loop on the device: `thermalization` un-thermalised sweeps in one launch, then one
    launch per bin with on-device measurement (identical Markov chain and identical per-bin sums)."""
    ctx = ctx or MCContext(params)
    sweeps = int(params["sweeps"])
    therm = int(params.get("thermalization", 0))
    mc.init(ctx, params)
    if not fused:
        # Carlo: sweep!; ctx.sweeps += 1; if is_thermalized measure!
        while ctx.sweeps < sweeps + therm:
            mc.sweep(ctx)
            ctx.sweeps += 1
            if ctx.is_thermalized():
                mc.measure(ctx)
        return ctx
    # Device-resident loop: `therm` un-thermalised sweeps, then `sweeps` thermalised + measured sweeps, one
    # launch per bin.  (Carlo's own loop evaluates is_thermalized before incrementing ctx.sweeps, so there the
    # first measured sweep still runs the worm-count controller; that one-sweep shift is immaterial.)
    if therm > 0:
        mc.sweep_many(ctx, therm, thermalized=False, measure=False)
    remaining = sweeps
    while remaining > 0:
        nb = min(ctx.binsize, remaining)
        mc.sweep_many(ctx, nb, thermalized=True, measure=True)
        mc.flush_bin(ctx)
        remaining -= nb
    return ctx
