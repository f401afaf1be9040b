"""ctypes wrapper of oracle/libsse_oracle.so — TEST INFRASTRUCTURE ONLY (see sse_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sse_b200.capi import ModelDesc, WalkerState, f64p, i64p, u64p, u8p, model_desc_from_model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsse_oracle.so")
_lib = None


def build(force: bool = False):
    src = os.path.join(_HERE, "sse_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])


def use_native_build() -> bool:
    """bench.py's CPU legs only: compile the oracle for THIS host's cores (-march=native, still -ffp-contract=off) into
    oracle/_native/ and load that instead of the portable build (SURVEY.md 8d asks for -march=native; the portable .so
    that travels with the repo is x86-64-v3).  Must be called before the first lib().  False if the compile fails."""
    global LIB_PATH, _lib
    out_dir = os.path.join(_HERE, "_native")
    out = os.path.join(out_dir, "libsse_oracle.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-o", out, os.path.join(_HERE, "sse_oracle.cpp")], stderr=subprocess.DEVNULL)
    except Exception:
        return False
    LIB_PATH = out
    _lib = None
    return True


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_model_create.restype = vp
        L.oracle_model_create.argtypes = [C.POINTER(ModelDesc)]
        L.oracle_model_destroy.argtypes = [vp]
        L.oracle_walker_create.restype = vp
        L.oracle_walker_create.argtypes = [vp, C.c_double, C.c_int32, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double]
        L.oracle_walker_destroy.argtypes = [vp]
        L.oracle_set_injected_stream.argtypes = [vp, u64p, C.c_int64]
        L.oracle_set_temperature.argtypes = [vp, C.c_double]
        L.oracle_sweep_capped.restype = C.c_int32
        L.oracle_sweep_capped.argtypes = [vp, C.c_int32, C.c_int32, C.c_uint64]
        L.oracle_rng_draws.restype = C.c_uint64
        L.oracle_rng_draws.argtypes = [vp]
        L.oracle_set_rng_draws.argtypes = [vp, C.c_uint64]
        L.oracle_stream_exhausted.restype = C.c_int32
        L.oracle_stream_exhausted.argtypes = [vp]
        L.oracle_flags.restype = C.c_uint32
        L.oracle_flags.argtypes = [vp]
        L.oracle_init.argtypes = [vp, C.c_int64, C.c_int32]
        L.oracle_sweep.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
        L.oracle_diagonal_update.argtypes = [vp]
        L.oracle_make_vertex_list.argtypes = [vp]
        L.oracle_worm_update.argtypes = [vp, C.c_int32]
        L.oracle_worm_traverse.restype = C.c_int64
        L.oracle_worm_traverse.argtypes = [vp, C.c_int32, C.c_int64, C.c_int32]
        L.oracle_measure.argtypes = [vp, f64p]
        L.oracle_n_obs.restype = C.c_int32
        L.oracle_n_obs.argtypes = [vp]
        L.oracle_fetch_accumulators.argtypes = [vp, f64p, i64p, C.c_int32]
        L.oracle_fetch_counters.argtypes = [vp, u64p, C.c_int32]
        L.oracle_opstring_length.restype = C.c_int64
        L.oracle_opstring_length.argtypes = [vp]
        L.oracle_get_state.argtypes = [vp, C.POINTER(WalkerState)]
        L.oracle_set_state.argtypes = [vp, C.POINTER(WalkerState)]
        L.oracle_get_vertex_list.argtypes = [vp, i64p, i64p, i64p]
        L.oracle_bench.argtypes = [vp, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, f64p]
        L.oracle_bench2.argtypes = [vp, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, f64p]
        _lib = L
    return _lib


class OracleModel:
    def __init__(self, model=None, desc=None, keep=None, sse_data=None):
        if desc is None:
            desc, keep, sse_data = model_desc_from_model(model)
        self.desc, self._keep, self.sse_data = desc, keep, sse_data
        self.n_sites = desc.n_sites
        self.handle = lib().oracle_model_create(C.byref(desc))

    def __del__(self):
        if getattr(self, "handle", None):
            lib().oracle_model_destroy(self.handle)
            self.handle = None

    def bench(self, T: float, n_threads: int, therm: int, sweeps: int, seed: int = 1, doublings: int = 0, per_level: int = 10):
        out = np.zeros(6)
        lib().oracle_bench2(self.handle, T, n_threads, doublings, per_level, therm, sweeps, seed, out.ctypes.data_as(f64p))
        return dict(seconds=out[0], visits=out[1], walker_sweeps=out[2], mean_n=out[3], mean_M=out[4], thread_seconds=out[5])


RNG_PHILOX, RNG_INJECTED, RNG_XOSHIRO = 0, 1, 2


class OracleWalker:
    """One reference-layout `MC` (src/sse.jl:6-24) on the CPU."""

    def __init__(self, model: OracleModel, T: float, seed: int = 0, walker_id: int = 0, rng_kind: int = RNG_PHILOX,
                 target_worm_length_fraction: float = 2.0, num_worms_attenuation_factor: float = 0.01,
                 init_num_worms: float = 5.0):
        self.model = model
        self.L = lib()
        self.h = self.L.oracle_walker_create(model.handle, T, rng_kind, seed, walker_id, target_worm_length_fraction,
                                             num_worms_attenuation_factor, init_num_worms)
        self._stream = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_walker_destroy(self.h)
            self.h = None

    def set_injected_stream(self, stream):
        if stream is None:
            self._stream = None
            self.L.oracle_set_injected_stream(self.h, None, 0)
        else:
            self._stream = np.ascontiguousarray(stream, dtype=np.uint64)
            self.L.oracle_set_injected_stream(self.h, self._stream.ctypes.data_as(u64p), len(self._stream))

    def sweep_capped(self, n_sweeps: int, thermalized: bool, cap: int) -> bool:
        """Screening aid: True if the walker needed more than `cap` visits (and was abandoned)."""
        return bool(self.L.oracle_sweep_capped(self.h, n_sweeps, int(thermalized), cap))

    def set_temperature(self, T: float):
        self.L.oracle_set_temperature(self.h, float(T))

    @property
    def rng_draws(self) -> int:
        return int(self.L.oracle_rng_draws(self.h))

    @property
    def stream_exhausted(self) -> bool:
        return bool(self.L.oracle_stream_exhausted(self.h))

    @property
    def flags(self) -> int:
        return int(self.L.oracle_flags(self.h))

    def init(self, init_opstring_cutoff: int = -1, diagonal_warmup_sweeps: int = 5):
        self.L.oracle_init(self.h, init_opstring_cutoff, diagonal_warmup_sweeps)

    def sweep(self, n_sweeps: int = 1, thermalized: bool = False, measure: bool = False):
        self.L.oracle_sweep(self.h, n_sweeps, int(thermalized), int(measure))

    def diagonal_update(self):
        self.L.oracle_diagonal_update(self.h)

    def make_vertex_list(self):
        self.L.oracle_make_vertex_list(self.h)

    def worm_update(self, thermalized: bool = False):
        self.L.oracle_worm_update(self.h, int(thermalized))

    def worm_traverse(self, l0: int, p0: int, wormfunc0: int) -> int:
        return int(self.L.oracle_worm_traverse(self.h, l0, p0, wormfunc0))

    def measure(self) -> np.ndarray:
        out = np.zeros(self.L.oracle_n_obs(self.h))
        self.L.oracle_measure(self.h, out.ctypes.data_as(f64p))
        return out

    def fetch_accumulators(self, reset: bool = True):
        sums = np.zeros(self.L.oracle_n_obs(self.h))
        counts = np.zeros(2, dtype=np.int64)
        self.L.oracle_fetch_accumulators(self.h, sums.ctypes.data_as(f64p), counts.ctypes.data_as(i64p), int(reset))
        return sums, counts

    def fetch_counters(self, reset: bool = False):
        out = np.zeros(4, dtype=np.uint64)
        self.L.oracle_fetch_counters(self.h, out.ctypes.data_as(u64p), int(reset))
        return dict(visits=int(out[0]), sweeps=int(out[1]), sum_n=int(out[2]), sum_M=int(out[3]))

    def get_state(self) -> dict:
        M = int(self.L.oracle_opstring_length(self.h))
        ops = np.zeros(max(M, 1), dtype=np.uint64)
        state = np.zeros(self.model.n_sites, dtype=np.uint8)
        st = WalkerState()
        st.operators = ops.ctypes.data_as(u64p)
        st.operators_len = M
        st.state = state.ctypes.data_as(u8p)
        self.L.oracle_get_state(self.h, C.byref(st))
        return dict(num_operators=int(st.num_operators), avg_worm_length=float(st.avg_worm_length),
                    num_worms=float(st.num_worms), operators=ops[:M].copy(), state=state, rng_draws=int(st.rng_draws),
                    T=float(st.T))

    def set_state(self, s: dict):
        ops = np.ascontiguousarray(s["operators"], dtype=np.uint64)
        state = np.ascontiguousarray(s["state"], dtype=np.uint8)
        st = WalkerState()
        st.num_operators = int(s["num_operators"])
        st.avg_worm_length = float(s.get("avg_worm_length", 1.0))
        st.num_worms = float(s.get("num_worms", 5.0))
        st.operators = ops.ctypes.data_as(u64p)
        st.operators_len = len(ops)
        st.state = state.ctypes.data_as(u8p)
        st.rng_draws = int(s.get("rng_draws", 0))
        st.T = float(s["T"])
        self.L.oracle_set_state(self.h, C.byref(st))

    def get_vertex_list(self):
        M = int(self.L.oracle_opstring_length(self.h))
        v = np.zeros((M, 4, 2), dtype=np.int64)
        vf = np.zeros((self.model.n_sites, 2), dtype=np.int64)
        vl = np.zeros((self.model.n_sites, 2), dtype=np.int64)
        self.L.oracle_get_vertex_list(self.h, v.ctypes.data_as(i64p), vf.ctypes.data_as(i64p), vl.ctypes.data_as(i64p))
        return v, vf, vl
