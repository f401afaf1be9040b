"""ctypes binding of libsse_b200.so (include/sse_b200.h) — the thin call layer the host uses.

This is the Python twin of the Julia `ccall` wrapper in julia/SSEB200.jl: one
function per C export, no logic.  There is NO fallback: if the shared library is missing or a call
returns a non-zero status an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsse_b200.so")

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_sites", C.c_int32), ("site_dim", u8p),
        ("n_bonds", C.c_int32), ("bond_type", i32p), ("bond_sites", i32p),
        ("n_types", C.c_int32), ("type_dims", i32p), ("type_vertex_off", i32p), ("type_diag_off", i32p),
        ("n_vertices", C.c_int32), ("weights", f64p), ("signs", i8p), ("leg_states", u8p),
        ("diag_vertices", i32p),
        ("max_worm", C.c_int32), ("trans_offset", i32p), ("trans_count", i32p),
        ("n_outcomes", C.c_int32), ("out_cumprob", f64p), ("out_target", i32p), ("out_leg", i32p), ("out_worm", i32p),
        ("energy_offset", C.c_double), ("norm_site_count", C.c_int32),
        ("n_estimators", C.c_int32), ("est_max_dim", C.c_int32), ("est_values", f64p),
    ]


class WalkersOpts(C.Structure):
    _fields_ = [
        ("n_walkers", C.c_int32), ("T", f64p), ("m_capacity", C.c_int64), ("n_capacity", C.c_int64),
        ("device", C.c_int32), ("seed", C.c_uint64), ("walker_id_offset", C.c_uint64),
        ("target_worm_length_fraction", C.c_double), ("num_worms_attenuation_factor", C.c_double),
        ("init_num_worms", C.c_double),
    ]


class WalkerState(C.Structure):
    _fields_ = [
        ("num_operators", C.c_int64), ("avg_worm_length", C.c_double), ("num_worms", C.c_double),
        ("operators", u64p), ("operators_len", C.c_int64), ("state", u8p), ("rng_draws", C.c_uint64),
        ("T", C.c_double),
    ]


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


_PTR_FIELDS = {
    "site_dim": (np.uint8, u8p), "bond_type": (np.int32, i32p), "bond_sites": (np.int32, i32p),
    "type_dims": (np.int32, i32p), "type_vertex_off": (np.int32, i32p), "type_diag_off": (np.int32, i32p),
    "weights": (np.float64, f64p), "signs": (np.int8, i8p), "leg_states": (np.uint8, u8p),
    "diag_vertices": (np.int32, i32p), "trans_offset": (np.int32, i32p), "trans_count": (np.int32, i32p),
    "out_cumprob": (np.float64, f64p), "out_target": (np.int32, i32p), "out_leg": (np.int32, i32p),
    "out_worm": (np.int32, i32p),
}


def build_model_desc(flat: dict, norm_site_count: int, est_values: np.ndarray | None):
    """flat = SSEData.flatten(); est_values = [n_est, n_sites, max_dim] f64 or None.
    Returns (ModelDesc, keepalive list)."""
    d = ModelDesc()
    keep = []
    for k, (dt, pt) in _PTR_FIELDS.items():
        a = np.ascontiguousarray(flat[k], dtype=dt)
        keep.append(a)
        setattr(d, k, _ptr(a, pt))
    for k in ("n_sites", "n_bonds", "n_types", "n_vertices", "max_worm", "n_outcomes"):
        setattr(d, k, int(flat[k]))
    d.energy_offset = float(flat["energy_offset"])
    d.norm_site_count = int(norm_site_count)
    if est_values is None or len(est_values) == 0:
        d.n_estimators = 0
        d.est_max_dim = 1
        d.est_values = None
    else:
        ev = np.ascontiguousarray(est_values, dtype=np.float64)
        assert ev.ndim == 3 and ev.shape[1] == flat["n_sites"]
        keep.append(ev)
        d.n_estimators = ev.shape[0]
        d.est_max_dim = ev.shape[2]
        d.est_values = _ptr(ev, f64p)
    return d, keep


def model_desc_from_model(model):
    """AbstractModel -> (ModelDesc, keepalive, sse_data): flattens generate_sse_data + estimator tables."""
    sse_data = model.generate_sse_data()
    flat = sse_data.flatten()
    ests = model.get_opstring_estimators()
    max_dim = int(max(s.dim for s in sse_data.sites))
    ev = None
    if ests:
        ev = np.stack([e.value_table(model, len(sse_data.sites), max_dim) for e in ests])
    desc, keep = build_model_desc(flat, model.normalization_site_count(), ev)
    return desc, keep, sse_data


class SSEError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libsse_b200.so (built by `__graft_entry__.build()` / csrc/Makefile). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SSEError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    sig = {
        "sse_last_error": (C.c_char_p, []),
        "sse_abi_version": (C.c_int32, []),
        "sse_model_create": (C.c_int32, [C.POINTER(ModelDesc), C.POINTER(vp)]),
        "sse_model_destroy": (C.c_int32, [vp]),
        "sse_walkers_create": (C.c_int32, [vp, C.POINTER(WalkersOpts), C.POINTER(vp)]),
        "sse_walkers_destroy": (C.c_int32, [vp]),
        "sse_set_stream": (C.c_int32, [vp, vp]),
        "sse_grow_capacity": (C.c_int32, [vp, C.c_int64, C.c_int64]),
        "sse_n_observables": (C.c_int32, [vp]),
        "sse_device_bytes": (C.c_int64, [vp]),
        "sse_walker_bytes": (C.c_int64, [vp, C.c_int64, C.c_int64]),
        "sse_init": (C.c_int32, [vp, C.c_int64, C.c_int32]),
        "sse_sweep": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32]),
        "sse_sync": (C.c_int32, [vp]),
        "sse_measure": (C.c_int32, [vp, f64p]),
        "sse_fetch_accumulators": (C.c_int32, [vp, f64p, i64p, C.c_int32]),
        "sse_accumulators_device_ptr": (C.c_int32, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "sse_fetch_counters": (C.c_int32, [vp, u64p, C.c_int32]),
        "sse_comm_unique_id": (C.c_int32, [C.c_char_p]),
        "sse_comm_init": (C.c_int32, [vp, C.c_char_p, C.c_int32, C.c_int32]),
        "sse_reduce_bins": (C.c_int32, [vp, i32p, C.c_int32, f64p, i64p, C.c_int32]),
        "sse_get_state": (C.c_int32, [vp, C.c_int32, C.POINTER(WalkerState)]),
        "sse_get_states": (C.c_int32, [vp, C.c_int32, C.c_int32, C.POINTER(WalkerState)]),
        "sse_set_state": (C.c_int32, [vp, C.c_int32, C.POINTER(WalkerState)]),
        "sse_set_states": (C.c_int32, [vp, C.c_int32, C.c_int32, C.POINTER(WalkerState)]),
        "sse_get_flags": (C.c_int32, [vp, u32p]),
        "sse_pt_log_weight_ratio": (C.c_int32, [vp, f64p, f64p]),
        "sse_set_temperature": (C.c_int32, [vp, f64p]),
        "sse_get_num_operators": (C.c_int32, [vp, i64p]),
        "sse_get_temperatures": (C.c_int32, [vp, f64p]),
        "sse_pt_set_ladder": (C.c_int32, [vp, i32p, C.c_int32]),
        "sse_pt_get_ladder": (C.c_int32, [vp, i32p]),
        "sse_pt_exchange": (C.c_int32, [vp, C.c_int32, C.c_uint64, C.c_uint64, i32p]),
        "sse_pt_uniforms": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_int32, f64p]),
        "sse_double_beta": (C.c_int32, [vp]),
        "sse_set_controller": (C.c_int32, [vp, C.c_double, C.c_double]),
        "sse_set_launch_shape": (C.c_int32, [vp, C.c_int32, C.c_int32]),
        "sse_advance": (C.c_int32, [vp, C.c_int32, C.c_uint64, C.c_int32, C.c_int32]),
        "sse_finish_sweeps": (C.c_int32, [vp, C.c_int32, C.c_int32]),
        "sse_continue_sweeps": (C.c_int32, [vp, C.c_int32, C.c_int32]),
        "sse_get_progress": (C.c_int32, [vp, u64p, u8p]),
        "sse_set_injected_stream": (C.c_int32, [vp, u64p, C.c_int64]),
        "sse_dbg_diagonal_update": (C.c_int32, [vp]),
        "sse_dbg_make_vertex_list": (C.c_int32, [vp]),
        "sse_dbg_worm_update": (C.c_int32, [vp, C.c_int32]),
        "sse_dbg_worm_traverse": (C.c_int32, [vp, C.c_int32, C.c_int64, C.c_int32, i64p]),
        "sse_dbg_get_vertex_list": (C.c_int32, [vp, C.c_int32, i64p, C.c_int64, i64p, i64p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError if the symbol is missing
        f.restype = res
        f.argtypes = args
    L._sse_signatures = sig
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "sse_last_error", "sse_abi_version", "sse_model_create", "sse_model_destroy", "sse_walkers_create",
    "sse_walkers_destroy", "sse_set_stream", "sse_grow_capacity", "sse_n_observables", "sse_device_bytes", "sse_walker_bytes", "sse_init", "sse_sweep",
    "sse_sync", "sse_measure", "sse_fetch_accumulators", "sse_accumulators_device_ptr", "sse_fetch_counters",
    "sse_comm_unique_id", "sse_comm_init", "sse_reduce_bins",
    "sse_get_state", "sse_get_states", "sse_set_state", "sse_set_states", "sse_get_flags", "sse_pt_log_weight_ratio", "sse_set_temperature",
    "sse_get_num_operators", "sse_get_temperatures", "sse_pt_set_ladder", "sse_pt_get_ladder", "sse_pt_exchange", "sse_pt_uniforms", "sse_double_beta", "sse_set_controller", "sse_set_launch_shape",
    "sse_advance", "sse_finish_sweeps", "sse_continue_sweeps", "sse_get_progress",
    "sse_set_injected_stream", "sse_dbg_diagonal_update", "sse_dbg_make_vertex_list",
    "sse_dbg_worm_update", "sse_dbg_worm_traverse", "sse_dbg_get_vertex_list",
]


def check(status: int):
    if status != 0:
        msg = lib().sse_last_error()
        raise SSEError(f"libsse_b200 status {status}: {msg.decode() if msg else '?'}")
