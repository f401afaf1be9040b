"""Object wrappers over the C ABI handles: `DeviceModel` (sse_model) and `Walkers` (sse_walkers).

Thin: every method is one C call (see capi.py / include/sse_b200.h).  The Carlo-facing mirror of the
reference's `MC` lives in mc.py on top of this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import WalkerState, WalkersOpts, check, f64p, i64p, u32p, u64p, u8p

OBS_FIXED = ["Sign", "OperatorCount", "SignOperatorCount", "SignOperatorCount2", "SignEnergy", "WormLengthFraction"]
OBS_PER_EST = ["Mag", "AbsMag", "Mag2", "Mag4", "MagChi"]


class DeviceModel:
    """sse_model: the flattened SSEData + estimator tables resident on the GPU."""

    def __init__(self, model=None, desc=None, keep=None, sse_data=None):
        if desc is None:
            desc, keep, sse_data = capi.model_desc_from_model(model)
        self.model, self.desc, self._keep, self.sse_data = model, desc, keep, sse_data
        self.n_sites = int(desc.n_sites)
        self.n_bonds = int(desc.n_bonds)
        self.n_estimators = int(desc.n_estimators)
        self.handle = C.c_void_p()
        self.L = capi.lib()
        check(self.L.sse_model_create(C.byref(desc), C.byref(self.handle)))

    def walker_bytes(self, m_capacity: int, n_capacity: int) -> int:
        """Device bytes per walker at these capacities (for sizing a batch to the GPU's memory)."""
        return int(self.L.sse_walker_bytes(self.handle, int(m_capacity), int(n_capacity)))

    def observable_names(self):
        names = list(OBS_FIXED)
        ests = self.model.get_opstring_estimators() if self.model is not None else []
        for e in range(self.n_estimators):
            prefix = ests[e].prefix if e < len(ests) else f"Est{e}"
            names += [f"Sign{prefix}{o}" for o in OBS_PER_EST]
        return names

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.L.sse_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Walkers:
    """sse_walkers: a batch of independent walkers (each one reference `MC`, src/sse.jl:6-24)."""

    def __init__(self, dmodel: DeviceModel, T, m_capacity: int, n_capacity: int | None = None, seed: int = 0,
                 walker_id_offset: int = 0, device: int = -1, target_worm_length_fraction: float = 2.0,
                 num_worms_attenuation_factor: float = 0.01, init_num_worms: float = 5.0):
        self.dmodel = dmodel
        self.L = capi.lib()
        self.T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
        self.n_walkers = len(self.T)
        o = WalkersOpts()
        o.n_walkers = self.n_walkers
        o.T = self.T.ctypes.data_as(f64p)
        o.m_capacity = int(m_capacity)
        o.n_capacity = int(n_capacity if n_capacity is not None else min(m_capacity, (1 << 22) - 1))
        o.device = device
        o.seed = seed
        o.walker_id_offset = walker_id_offset
        o.target_worm_length_fraction = target_worm_length_fraction
        o.num_worms_attenuation_factor = num_worms_attenuation_factor
        o.init_num_worms = init_num_worms
        self.m_capacity = o.m_capacity
        self.n_capacity = o.n_capacity
        self.target_worm_length_fraction = float(target_worm_length_fraction)
        self.num_worms_attenuation_factor = float(num_worms_attenuation_factor)
        self.handle = C.c_void_p()
        check(self.L.sse_walkers_create(dmodel.handle, C.byref(o), C.byref(self.handle)))
        self.n_obs = int(self.L.sse_n_observables(self.handle))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.L.sse_walkers_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- Carlo.AbstractMC surface ------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        check(self.L.sse_set_stream(self.handle, C.c_void_p(cuda_stream)))

    def grow_capacity(self, m_capacity: int, n_capacity: int):
        """Move every walker to larger arrays (sse_grow_capacity); a pending m_capacity overflow is cleared."""
        check(self.L.sse_grow_capacity(self.handle, int(m_capacity), int(n_capacity)))
        self.m_capacity = int(m_capacity)
        self.n_capacity = int(n_capacity)

    def device_bytes(self) -> int:
        return int(self.L.sse_device_bytes(self.handle))

    def init(self, init_opstring_cutoff: int = -1, diagonal_warmup_sweeps: int = 5):
        check(self.L.sse_init(self.handle, init_opstring_cutoff, diagonal_warmup_sweeps))

    def sweep(self, n_sweeps: int = 1, thermalized: bool = False, measure: bool = False, sync: bool = True, auto_grow: bool = False):
        """Carlo.sweep! x n_sweeps for every walker.  auto_grow (needs sync): when a string outgrows m_capacity — the
        reference would resize it (sse.jl:138-145) — the arrays are doubled with sse_grow_capacity and the call is resumed
        (sse_continue_sweeps), and n_capacity grows ahead of the operator count; the trajectories are unchanged."""
        check(self.L.sse_sweep(self.handle, n_sweeps, int(thermalized), int(measure)))
        if not sync:
            return
        if not auto_grow:
            self.sync()
            return
        while True:
            try:
                self.sync()
                break
            except capi.SSEError as e:
                if "m_capacity" not in str(e):
                    raise
                self.grow_capacity(2 * self.m_capacity, self.n_capacity)
                check(self.L.sse_continue_sweeps(self.handle, int(thermalized), int(measure)))
        if self.num_operators().max() > 0.8 * self.n_capacity:
            self.grow_capacity(self.m_capacity, min((1 << 22) - 1, int(1.5 * self.n_capacity)))

    def advance(self, visit_budget: int, max_sweeps: int = 2**31 - 1, thermalized: bool = False, measure: bool = False,
                sync: bool = True):
        """Free-running sweeps (sse_advance): every walker does `visit_budget` worm visits (or `max_sweeps` sweeps) and is
        parked wherever it is; the next advance()/sweep() resumes there.  Same Markov chain per walker as sweep()."""
        check(self.L.sse_advance(self.handle, int(max_sweeps), int(visit_budget), int(thermalized), int(measure)))
        if sync:
            self.sync()

    def finish_sweeps(self, thermalized: bool = False, measure: bool = False, sync: bool = True):
        """Complete the sweeps advance() left in flight (needed before get_state / measure / double_beta)."""
        check(self.L.sse_finish_sweeps(self.handle, int(thermalized), int(measure)))
        if sync:
            self.sync()

    def progress(self):
        """(sweeps_done[n_walkers], in_flight[n_walkers]): completed sweeps since init/set_state, parked inside a sweep?"""
        sd = np.zeros(self.n_walkers, dtype=np.uint64)
        fl = np.zeros(self.n_walkers, dtype=np.uint8)
        check(self.L.sse_get_progress(self.handle, sd.ctypes.data_as(u64p), fl.ctypes.data_as(u8p)))
        return sd, fl.astype(bool)

    def sync(self):
        check(self.L.sse_sync(self.handle))

    def measure(self) -> np.ndarray:
        out = np.zeros((self.n_walkers, self.n_obs))
        check(self.L.sse_measure(self.handle, out.ctypes.data_as(f64p)))
        return out

    def fetch_accumulators(self, reset: bool = True):
        sums = np.zeros((self.n_walkers, self.n_obs))
        counts = np.zeros((self.n_walkers, 2), dtype=np.int64)
        check(self.L.sse_fetch_accumulators(self.handle, sums.ctypes.data_as(f64p), counts.ctypes.data_as(i64p), int(reset)))
        return sums, counts

    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclUniqueId for comm_init (rank 0 creates it, the host broadcasts it)."""
        buf = C.create_string_buffer(128)
        check(capi.lib().sse_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """Join the NCCL communicator used by reduce_bins (one rank per GPU)."""
        assert len(unique_id) == 128
        check(self.L.sse_comm_init(self.handle, C.create_string_buffer(unique_id, 128), int(rank), int(nranks)))

    def reduce_bins(self, group=None, n_groups: int = 1, reset: bool = True):
        """One bin: per-group sums of the accumulators over this rank's walkers and, after comm_init, over all ranks
        (sse_reduce_bins).  Returns (sums[n_groups, n_obs], counts[n_groups, 2])."""
        g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        sums = np.zeros((n_groups, self.n_obs))
        counts = np.zeros((n_groups, 2), dtype=np.int64)
        check(self.L.sse_reduce_bins(self.handle, None if g is None else g.ctypes.data_as(capi.i32p), int(n_groups),
                                     sums.ctypes.data_as(f64p), counts.ctypes.data_as(i64p), int(reset)))
        return sums, counts

    def accumulators_device_ptr(self):
        s, c = C.c_void_p(), C.c_void_p()
        check(self.L.sse_accumulators_device_ptr(self.handle, C.byref(s), C.byref(c)))
        return s.value, c.value

    def fetch_counters(self, reset: bool = False) -> dict:
        out = np.zeros(16, dtype=np.uint64)
        check(self.L.sse_fetch_counters(self.handle, out.ctypes.data_as(u64p), int(reset)))
        names = ("visits", "sweeps", "sum_n", "sum_M", "cycles_build", "cycles_worm", "cycles_finish", "cycles_idle",
                 "lane_iters", "warp_iters", "tasks")
        return {k: int(out[i]) for i, k in enumerate(names)}

    def get_state(self, walker: int) -> dict:
        ops = np.zeros(self.m_capacity + 32, dtype=np.uint64)
        state = np.zeros(self.dmodel.n_sites, dtype=np.uint8)
        st = WalkerState()
        st.operators = ops.ctypes.data_as(u64p)
        st.operators_len = len(ops)
        st.state = state.ctypes.data_as(u8p)
        check(self.L.sse_get_state(self.handle, walker, C.byref(st)))
        return dict(num_operators=int(st.num_operators), avg_worm_length=float(st.avg_worm_length),
                    num_worms=float(st.num_worms), operators=ops[: st.operators_len].copy(), state=state,
                    rng_draws=int(st.rng_draws), T=float(st.T))

    def get_states(self, first: int = 0, count: int | None = None) -> list:
        """Checkpoint data of `count` walkers from `first` with one round of device-to-host copies (sse_get_states)."""
        count = self.n_walkers - first if count is None else count
        arr = (WalkerState * count)()
        check(self.L.sse_get_states(self.handle, first, count, arr))  # sizes
        bufs = []
        for j in range(count):
            ops = np.zeros(max(int(arr[j].operators_len), 1), dtype=np.uint64)
            state = np.zeros(self.dmodel.n_sites, dtype=np.uint8)
            arr[j].operators = ops.ctypes.data_as(u64p)
            arr[j].operators_len = len(ops)
            arr[j].state = state.ctypes.data_as(u8p)
            bufs.append((ops, state))
        check(self.L.sse_get_states(self.handle, first, count, arr))
        return [dict(num_operators=int(st.num_operators), avg_worm_length=float(st.avg_worm_length), num_worms=float(st.num_worms),
                     operators=ops[: st.operators_len].copy(), state=state, rng_draws=int(st.rng_draws), T=float(st.T))
                for st, (ops, state) in zip(arr, bufs)]

    @staticmethod
    def _fill_state(st, s: dict):
        ops = np.ascontiguousarray(s["operators"], dtype=np.uint64)
        state = np.ascontiguousarray(s["state"], dtype=np.uint8)
        st.num_operators = int(s["num_operators"])
        st.avg_worm_length = float(s.get("avg_worm_length", 1.0))
        st.num_worms = float(s.get("num_worms", 5.0))
        st.operators = ops.ctypes.data_as(u64p)
        st.operators_len = len(ops)
        st.state = state.ctypes.data_as(u8p)
        st.rng_draws = int(s.get("rng_draws", 0))
        st.T = float(s["T"])
        return ops, state  # keep the buffers alive until the call returns

    def set_state(self, walker: int, s: dict):
        st = WalkerState()
        keep = self._fill_state(st, s)
        check(self.L.sse_set_state(self.handle, walker, C.byref(st)))
        del keep

    def set_states(self, states: list, first: int = 0):
        """Restore walkers first .. first+len(states)-1 with one round of host-to-device copies (sse_set_states); all
        states are validated before anything is copied."""
        arr = (WalkerState * len(states))()
        keep = [self._fill_state(arr[j], s) for j, s in enumerate(states)]
        check(self.L.sse_set_states(self.handle, first, len(states), arr))
        del keep

    def get_flags(self) -> np.ndarray:
        f = np.zeros(self.n_walkers, dtype=np.uint32)
        check(self.L.sse_get_flags(self.handle, f.ctypes.data_as(u32p)))
        return f

    def num_operators(self) -> np.ndarray:
        n = np.zeros(self.n_walkers, dtype=np.int64)
        check(self.L.sse_get_num_operators(self.handle, n.ctypes.data_as(i64p)))
        return n

    def pt_log_weight_ratio(self, new_T) -> np.ndarray:
        nt = np.ascontiguousarray(new_T, dtype=np.float64)
        out = np.zeros(self.n_walkers)
        check(self.L.sse_pt_log_weight_ratio(self.handle, nt.ctypes.data_as(f64p), out.ctypes.data_as(f64p)))
        return out

    def set_temperature(self, T):
        t = np.ascontiguousarray(T, dtype=np.float64)
        assert len(t) == self.n_walkers
        check(self.L.sse_set_temperature(self.handle, t.ctypes.data_as(f64p)))
        self.T = t

    def temperatures(self) -> np.ndarray:
        t = np.zeros(self.n_walkers)
        check(self.L.sse_get_temperatures(self.handle, t.ctypes.data_as(f64p)))
        return t

    def pt_set_ladder(self, walker_at_rank):
        """Name the walkers of a temperature ladder in rank order (device-side replica exchange, sse_pt_exchange)."""
        o = np.ascontiguousarray(walker_at_rank, dtype=np.int32)
        check(self.L.sse_pt_set_ladder(self.handle, o.ctypes.data_as(capi.i32p), len(o)))
        self._n_ladder = len(o)

    def pt_get_ladder(self) -> np.ndarray:
        o = np.zeros(self._n_ladder, dtype=np.int32)
        check(self.L.sse_pt_get_ladder(self.handle, o.ctypes.data_as(capi.i32p)))
        return o

    def pt_exchange(self, parity: int, seed: int, step: int) -> int:
        """One round of neighbour swaps decided on the device; returns the number of accepted pairs.  self.T is refreshed."""
        acc = C.c_int32(0)
        check(self.L.sse_pt_exchange(self.handle, int(parity), int(seed), int(step), C.byref(acc)))
        self.T = self.temperatures()
        return int(acc.value)

    def set_launch_shape(self, worm_warps: int = 0, stream_warps: int = 0):
        """Launch shape of sweep()/advance(): warps per CTA that chase worms (one lane = one walker) and warps that run the
        streaming phases (one warp = one walker); 0 = automatic.  Results are bit-identical for every shape."""
        check(self.L.sse_set_launch_shape(self.handle, int(worm_warps), int(stream_warps)))

    def set_controller(self, target_worm_length_fraction: float | None = None, num_worms_attenuation_factor: float | None = None):
        """The worm-count controller's two parameters (sse.jl:34-35), changeable between launches."""
        if target_worm_length_fraction is not None:
            self.target_worm_length_fraction = float(target_worm_length_fraction)
        if num_worms_attenuation_factor is not None:
            self.num_worms_attenuation_factor = float(num_worms_attenuation_factor)
        check(self.L.sse_set_controller(self.handle, self.target_worm_length_fraction, self.num_worms_attenuation_factor))

    def double_beta(self):
        """Thermalisation aid (not in the reference): (state, S_M) -> (state, S_M S_M) at T/2 for every walker; the
        controller's average worm length doubles with it."""
        check(self.L.sse_double_beta(self.handle))
        self.T = self.T / 2.0

    def thermalize_by_beta_doubling(self, doublings: int, sweeps_per_level: int = 10, final_sweeps: int = 0,
                                    attenuation: float = 0.1, init_kwargs: dict | None = None):
        """Reach the target temperatures self.T from 2**doublings times hotter walkers: init! at T*2**doublings, then
        `sweeps_per_level` unthermalised sweeps and one doubling per level, then `final_sweeps` at the target.
        During the levels the worm-count controller runs with `attenuation` instead of the reference's 0.01 so that it
        follows the worm length, which grows by more than 2x per level: with 0.01 the worm count stays tuned for the hot
        levels and the first cold sweeps cost 50x an equilibrium sweep (measured; the reference's own cold start has the
        same transient).  The chain that follows is the reference's Markov chain; only its starting point and the
        controller's starting values differ from Carlo.init!, so thermalisation sweeps at the target temperature are
        still the caller's responsibility."""
        target = self.T.copy()
        atten0 = self.num_worms_attenuation_factor
        self.set_temperature(target * 2.0 ** doublings)
        self.init(**(init_kwargs or {}))
        self.set_controller(num_worms_attenuation_factor=attenuation)
        for _ in range(doublings):
            self.sweep(sweeps_per_level, thermalized=False)
            self.double_beta()
        self.set_controller(num_worms_attenuation_factor=atten0)
        # T halves exactly (a power of two), so the target is recovered bit for bit
        assert np.array_equal(self.T, target)
        if final_sweeps:
            self.sweep(final_sweeps, thermalized=False)

    # --- parity hooks ------------------------------------------------------------------------------
    def set_injected_stream(self, stream):
        if stream is None:
            check(self.L.sse_set_injected_stream(self.handle, None, 0))
            return
        s = np.ascontiguousarray(stream, dtype=np.uint64)
        assert s.ndim == 2 and s.shape[0] == self.n_walkers
        check(self.L.sse_set_injected_stream(self.handle, s.ctypes.data_as(u64p), s.shape[1]))

    def dbg_diagonal_update(self):
        check(self.L.sse_dbg_diagonal_update(self.handle))

    def dbg_make_vertex_list(self):
        check(self.L.sse_dbg_make_vertex_list(self.handle))

    def dbg_worm_update(self, thermalized: bool = False):
        check(self.L.sse_dbg_worm_update(self.handle, int(thermalized)))

    def dbg_worm_traverse(self, l0: int, p0: int, wormfunc0: int) -> np.ndarray:
        out = np.zeros(self.n_walkers, dtype=np.int64)
        check(self.L.sse_dbg_worm_traverse(self.handle, l0, p0, wormfunc0, out.ctypes.data_as(i64p)))
        return out

    def dbg_get_vertex_list(self, walker: int, M: int):
        v = np.zeros((M, 4, 2), dtype=np.int64)
        vf = np.zeros((self.dmodel.n_sites, 2), dtype=np.int64)
        vl = np.zeros((self.dmodel.n_sites, 2), dtype=np.int64)
        check(self.L.sse_dbg_get_vertex_list(self.handle, walker, v.ctypes.data_as(i64p), M, vf.ctypes.data_as(i64p),
                                             vl.ctypes.data_as(i64p)))
        return v, vf, vl
