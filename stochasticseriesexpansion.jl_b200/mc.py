"""`MC`: the Carlo.AbstractMC implementation backed by the B200 sweep kernels.

Mirror of `mutable struct MC <: AbstractMC` and its Carlo methods in /root/reference/src/sse.jl
(MC(params) :26-45, init! :47-60, sweep! :62-68, measure! :70-87, write/read_checkpoint :89-107,
register_evaluables :111-134, parallel-tempering hooks :390-405) for a BATCH of independent walkers:
`params["T"]` may be a scalar or a vector (one walker per entry), `params["n_walkers"]` replicates a
scalar T.  Each method is one C-ABI call (see julia/SSEB200.jl for the Julia twin).
"""
from __future__ import annotations

import numpy as np

from . import estimators as _est
from .walkers import OBS_FIXED, DeviceModel, Walkers


def operator_count_bound(sse_data, T_min: float) -> float:
    """Upper bound on the mean operator count: <n> = beta * sum_b <eps_b - H_b> <= beta * sum_b lambda_max(eps_b - H_b).
    The shifted bond operator is rebuilt from the vertex table (weights with signs between leg states)."""
    lam = []
    for vd in sse_data.vertex_data:
        d0, d1 = vd.dims
        Wm = np.zeros((d0 * d1, d0 * d1))
        ls = vd.leg_states.astype(np.int64) - 1
        rows = ls[0] + d0 * ls[1]
        cols = ls[2] + d0 * ls[3]
        Wm[rows, cols] = vd.weights * vd.signs
        lam.append(float(np.linalg.eigvalsh(0.5 * (Wm + Wm.T)).max()))
    total = sum(lam[b.type - 1] for b in sse_data.bonds)
    return total / T_min


def default_capacity(sse_data, T_min: float):
    """(m_capacity, n_capacity) that the string growth rule M <- 1.5 M + 100 while n >= M/2 (src/sse.jl:138-145)
    cannot exceed: M <= 3 n + 100 in the worst case, n <= the spectral bound (+ fluctuations)."""
    nb = operator_count_bound(sse_data, T_min)
    nb = nb + 8.0 * np.sqrt(nb) + 64
    return int(3.0 * nb + 1124), int(min(nb + 256, (1 << 22) - 1))


class MC:
    def __init__(self, params: dict):
        self.params = params
        self.model = params["model"](params)  # sse.jl:27
        self.dmodel = DeviceModel(self.model)  # generate_sse_data + estimator tables -> device (sse.jl:28)
        T = np.atleast_1d(np.asarray(params["T"], dtype=np.float64))
        if T.size == 1 and int(params.get("n_walkers", 1)) > 1:
            T = np.repeat(T, int(params["n_walkers"]))
        self.T = T
        m_def, n_def = default_capacity(self.dmodel.sse_data, float(T.min()))
        m_cap = int(params.get("m_capacity", m_def))
        n_cap = int(params.get("n_capacity", n_def))
        self.walkers = Walkers(
            self.dmodel,
            T,
            m_capacity=m_cap,
            n_capacity=n_cap,
            seed=int(params.get("seed", 0)),
            walker_id_offset=int(params.get("walker_id_offset", 0)),
            device=int(params.get("device", -1)),
            target_worm_length_fraction=float(params.get("target_worm_length_fraction", 2.0)),  # sse.jl:34
            num_worms_attenuation_factor=float(params.get("num_worms_attenuation_factor", 0.01)),  # sse.jl:35
            init_num_worms=float(params.get("init_num_worms", 5)),  # sse.jl:37
        )
        self.obs_names = self.dmodel.observable_names()

    @property
    def n_walkers(self) -> int:
        return self.walkers.n_walkers

    # --- Carlo.AbstractMC ----------------------------------------------------------------------
    def init(self, ctx, params: dict):
        """Carlo.init! (sse.jl:47-60).  Optional extension (default off): `beta_doublings = k` starts the walkers 2**k
        times hotter and grows them with sse_double_beta (`beta_doubling_sweeps` sweeps per level, default 10) before
        Carlo's own thermalisation sweeps begin."""
        ik = dict(init_opstring_cutoff=int(params.get("init_opstring_cutoff", -1)),
                  diagonal_warmup_sweeps=int(params.get("diagonal_warmup_sweeps", 5)))
        k = int(params.get("beta_doublings", 0))
        if k > 0:
            self.walkers.thermalize_by_beta_doubling(k, sweeps_per_level=int(params.get("beta_doubling_sweeps", 10)),
                                                     init_kwargs=ik)
        else:
            self.walkers.init(**ik)

    def sweep(self, ctx):
        """Carlo.sweep! (sse.jl:62-68)"""
        self.walkers.sweep(1, thermalized=ctx.is_thermalized(), measure=False)

    def measure(self, ctx):
        """Carlo.measure! (sse.jl:70-87): one vector observable (over walkers) per name."""
        obs = self.walkers.measure()
        for i, name in enumerate(self.obs_names):
            if name == "WormLengthFraction":
                # pushed by worm_update itself in the reference (sse.jl:200-202)
                if np.all(np.isfinite(obs[:, i])):
                    ctx.measure(name, obs[:, i])
            else:
                ctx.measure(name, obs[:, i])

    # device-resident variants used by carlo.run(fused=True)
    def sweep_many(self, ctx, n_sweeps: int, thermalized: bool, measure: bool, sync: bool = True):
        self.walkers.sweep(n_sweeps, thermalized=thermalized, measure=measure, sync=sync)
        ctx.sweeps += n_sweeps

    def flush_bin(self, ctx):
        """Turn the device accumulators into one Carlo bin per observable."""
        sums, counts = self.walkers.fetch_accumulators(reset=True)
        for i, name in enumerate(self.obs_names):
            c = counts[:, 1] if name == "WormLengthFraction" else counts[:, 0]
            if np.all(c > 0):
                ctx.add_bin(name, sums[:, i] / c)

    def write_checkpoint(self) -> dict:
        """Carlo.write_checkpoint (sse.jl:89-107): the reference's five fields per walker (+ stream position)."""
        return {"walkers": self.walkers.get_states()}

    def read_checkpoint(self, data: dict):
        self.walkers.set_states(list(data["walkers"]))

    def parallel_tempering_log_weight_ratio(self, parameter: str, new_value):
        """sse.jl:390-396"""
        if parameter != "T":
            raise ValueError(f"unsupported parallel tempering parameter {parameter}")
        nv = np.broadcast_to(np.asarray(new_value, dtype=np.float64), (self.n_walkers,))
        return self.walkers.pt_log_weight_ratio(nv)

    def parallel_tempering_change_parameter(self, parameter: str, new_value):
        """sse.jl:398-405"""
        if parameter != "T":
            raise ValueError(f"unsupported parallel tempering parameter {parameter}")
        nv = np.ascontiguousarray(np.broadcast_to(np.asarray(new_value, dtype=np.float64), (self.n_walkers,)))
        self.walkers.set_temperature(nv)
        self.T = nv

    # --- evaluables ------------------------------------------------------------------------------
    @staticmethod
    def register_evaluables(evaluator, params: dict, model=None):
        """Carlo.register_evaluables (sse.jl:111-134)."""
        model = model or params["model"](params)
        for est in model.get_opstring_estimators():
            _est.register_evaluables(est, evaluator)
        nsc = model.normalization_site_count()
        evaluator.evaluate("Energy", lambda se, s: se / s, ("SignEnergy", "Sign"))
        evaluator.evaluate(
            "SpecificHeat",
            lambda sn2, sn, s: (sn2 / s - sn * sn / s ** 2 - sn / s) / nsc,
            ("SignOperatorCount2", "SignOperatorCount", "Sign"),
        )


def evaluate_walker(ctx, mc: MC, walker: int):
    """Jackknife results {name: (mean, error)} of one walker of the batch from the ctx bins."""
    from .carlo import Evaluator

    bins = {k: np.array(v)[:, walker] for k, v in ctx.bins.items()}
    ev = Evaluator(bins)
    MC.register_evaluables(ev, mc.params, mc.model)
    return ev.results


def evaluate_group(ctx, mc: MC, walkers) -> dict:
    """Pool the bins of several independent walkers at the same temperature (they are independent chains)."""
    from .carlo import Evaluator

    idx = list(walkers)
    bins = {k: np.array(v)[:, idx].reshape(-1) for k, v in ctx.bins.items()}
    ev = Evaluator(bins)
    MC.register_evaluables(ev, mc.params, mc.model)
    return ev.results
