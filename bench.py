#!/usr/bin/env python
"""bench.py — throughput of the SSE sweep hot path (diagonal update -> vertex records -> worm update).

Metric (BASELINE.json / SURVEY.md §8d): worm operator-vertex visits per second = sum of the lengths
returned by worm_traverse! (src/sse.jl:302) over all walkers and sweeps / time of the WHOLE sweep.
Workload at every N (default): BASELINE.json configs[2], 2D square-lattice S=1/2 Heisenberg AFM L=64, beta=64,
as many walkers per GPU as its memory holds (weak scaling: walkers shard over ranks, no data-path collective).
`--L 32 --beta 32 --walkers 4096` is configs[1]; a short run of it is reported as `secondary` in the same line.

A "step" is one persistent launch (sse_advance) in which EVERY walker does `--visits-per-step` worm visits together with
all the diagonal updates, record builds and measurements of the sweeps it passes through; walkers are parked wherever
their budget ends and resume there in the next step, so every step is the same amount of work.

  python bench.py --gpus 1 --steps K --warmup W            (our arm)
  torchrun ... bench.py --gpus N ...                        (one rank per GPU, NCCL only for bin reduction)
  python bench.py --impl reference ...                      (CPU oracle on all host cores, same metric)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "operator-vertex visits/sec"
UNIT = "visits/s"

# profiles/r2_chase_lanes.txt (B200, profiles/tools/chase2.cu): dependent 16-byte ld.global.cg + 4-byte st.global per hop,
# one chain per lane, chains in flight -> hops/s.  The ceiling of a visit that costs nothing but its load and its store.
CHASE_POINTS = [(4736, 5.98e9), (9472, 1.05e10), (18944, 1.67e10), (37888, 1.99e10), (75776, 2.03e10)]


def chain_ceiling(chains):
    xs = [math.log(c) for c, _ in CHASE_POINTS]
    ys = [v for _, v in CHASE_POINTS]
    if chains <= CHASE_POINTS[0][0]:
        return ys[0] * chains / CHASE_POINTS[0][0]
    return float(np.interp(math.log(chains), xs, ys))


def capacities(L, beta):
    n_bonds = 2 * L * L
    n_est = 0.71 * beta * n_bonds  # <n> ~ beta * N_b * (|e_bond| + offset) for the energy_offset_factor = 0.25 tables
    return int(4.0 * n_est) + 16384, int(1.08 * n_est) + 4096


def model_params(args):
    import sse_b200 as S

    return dict(model=S.MagnetModel, lattice=dict(unitcell=S.UnitCells.square, size=(args.L, args.L)), J=1.0,
                s_half_deterministic=bool(args.deterministic), measure=["magnetization", "staggered_magnetization"])


def config_block(args):
    cfg = {(32, 32.0): "BASELINE.json configs[1]", (64, 64.0): "BASELINE.json configs[2]"}.get((args.L, args.beta), "custom size")
    return {
        "workload": f"2D square-lattice S=1/2 Heisenberg AFM L={args.L}, beta={args.beta}, independent walkers ({cfg})",
        "energy_offset_factor": 0.0 if args.deterministic else 0.25,
        "thermalisation": f"beta doubling x{args.beta_doublings} ({args.therm_per_level} sweeps per level) + {args.therm} sweeps at the target",
    }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val == "Active":
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def committed_profile():
    """Per-visit figures of sse::k_sweep from the committed ncu --set full capture (profiles/r2_ksweep_ncu.json):
    DRAM bytes and warp-instructions; None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ksweep_ncu.json")) as f:
            return json.load(f)
    except Exception:
        return None


def algorithmic_bytes(sum_M, sum_n, visits, measured_sweeps=0):
    """B_sweep = 8M (K1 read+write op codes) + 4M + 16n (K2) + 64V (K3) (+4M per measured sweep)."""
    return 8.0 * sum_M + 4.0 * sum_M + 16.0 * sum_n + 64.0 * visits + 4.0 * measured_sweeps


def host_cpu():
    """CPU model and core count of the box (SURVEY.md 8d asks for both next to the CPU figure)."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return f"{model}, nproc {os.cpu_count()}"


def cpu_baseline(args, cores=None, sweeps=None, native=True):
    """The CPU oracle (reference data layout, xoshiro256++ stream) on the host cores: one independent walker
    per thread, as Carlo runs one MC per MPI rank (docs/src/tutorial.md:49); brought to the target temperature the
    same way as the device walkers (beta doubling + thermalisation sweeps)."""
    import sse_b200  # noqa: F401
    import oracle as oracle_mod
    from oracle import OracleModel

    oracle_mod.build()
    is_native = bool(native and oracle_mod.use_native_build())
    model = model_params(args)
    om = OracleModel(model["model"](model))
    cores = cores or os.cpu_count() or 1
    sweeps = args.cpu_sweeps if sweeps is None else sweeps
    r = om.bench(1.0 / args.beta, cores, args.cpu_therm, sweeps, seed=args.seed, doublings=args.beta_doublings,
                 per_level=args.therm_per_level)
    return r, cores, sweeps, is_native


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s = args.cpu_sweeps_per_step
    r, cores, sweeps, native = cpu_baseline(args, sweeps=(args.steps + args.warmup) * s)
    value = r["visits"] / r["seconds"]
    sample = (f"{cores} independent walkers (one per host thread), thermalised like the device walkers, then "
              f"{sweeps} timed sweeps each = ({args.steps} steps + {args.warmup} warm-up) x {s} sweeps; mean n={r['mean_n']:.0f}, "
              f"M={r['mean_M']:.0f}; C++ oracle, reference data layout, xoshiro256++, -O3 {'-march=native' if native else '-march=x86-64-v3'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] * args.steps / (args.steps + args.warmup) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/f64", "data": "synthetic",
        "config": config_block(args),
        "detail": {"note": "reference CPU path = C++ oracle restating src/sse.jl (julia is not installed in this image); "
                           "per-walker data layout of the reference", "per_core": r["visits"] / r["thread_seconds"]},
        "sweeps_per_s": r["walker_sweeps"] / r["seconds"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "host": host_cpu()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def setup_walkers(args, dev_index, rank, n_walkers, stream):
    """Model, walkers, thermalisation (untimed).  Returns (walkers, dmodel, T, setup seconds)."""
    from sse_b200.walkers import DeviceModel, Walkers

    t0 = time.time()
    mp = model_params(args)
    dm = DeviceModel(model=mp["model"](mp))
    m_cap, n_cap = capacities(args.L, args.beta)
    if args.m_capacity:
        m_cap = args.m_capacity
    if args.n_capacity:
        n_cap = args.n_capacity
    W = n_walkers
    if W <= 0:  # as many as the GPU's memory holds
        import torch

        free, _total = torch.cuda.mem_get_info(dev_index)
        per = dm.walker_bytes(m_cap, n_cap)
        W = int((free - (2 << 30)) * 0.98 / per)
        W -= W % 148
        if args.max_walkers:
            W = min(W, args.max_walkers)
    T = 1.0 / args.beta
    wk = Walkers(dm, np.full(W, T), m_capacity=m_cap, n_capacity=n_cap, seed=args.seed, walker_id_offset=rank * W, device=dev_index)
    wk.set_stream(stream)
    if args.worm_warps or args.stream_warps:
        wk.set_launch_shape(args.worm_warps, args.stream_warps)
    if args.beta_doublings > 0:
        wk.thermalize_by_beta_doubling(args.beta_doublings, sweeps_per_level=args.therm_per_level)
    else:
        wk.init()
    if args.therm:
        wk.sweep(args.therm, thermalized=False)
    return wk, dm, T, W, time.time() - t0


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sweep backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from sse_b200.walkers import Walkers

    stream = torch.cuda.Stream(device=dev)
    wk, dm, T, W, t_setup = setup_walkers(args, local_rank, rank, args.walkers, stream.cuda_stream)
    B = int(args.visits_per_step)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):  # the first one also takes the walkers out of step
        wk.advance(B, thermalized=True)

    # ---- timed region 1: kernel-only, state resident in HBM ----
    wk.fetch_counters(reset=True)
    sampler = ClockSampler(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(stream):
        events[0].record(stream)
        for k in range(args.steps):
            wk.advance(B, thermalized=True, measure=False, sync=False)
            events[k + 1].record(stream)
    barrier()
    wk.sync()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = events[0].elapsed_time(events[-1])
    launch_ms = [events[k].elapsed_time(events[k + 1]) for k in range(args.steps)]
    cnt = wk.fetch_counters(reset=True)

    # ---- timed region 2: end to end through the public API with host buffers ----
    n_obs = wk.n_obs
    T_host = torch.full((W,), T, dtype=torch.float64).pin_memory()
    T_np = T_host.numpy()
    if world > 1:  # the library's own NCCL communicator (sse_comm_init): the id travels over torch.distributed
        box = [Walkers.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        wk.comm_init(box[0], rank, world)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))  # bounded so that a large --steps still fits the driver's slot
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        wk.set_temperature(T_np)                                     # H2D: this step's parameters
        wk.advance(B, thermalized=True, measure=True, sync=False)    # sweeps + on-device estimators
        sums, counts = wk.reduce_bins(None, 1, reset=True)           # the bin: summed over walkers on the device, over
                                                                     # ranks by NCCL inside the library, then D2H
    barrier()
    e2e_s = time.perf_counter() - t0
    cnt2 = wk.fetch_counters(reset=True)
    energy = float(sums[0, 4] / sums[0, 0]) if counts[0, 0] > 0 else None

    # ---- timed region 3: the call pattern of a Carlo job (julia/SSEB200.jl): sweep! = one launch of one sweep + sync,
    #      measure! = sse_measure with all observables copied to the host; and its batched form (sweeps_per_call) ----
    carlo = None
    if args.carlo_steps > 0:
        wk.finish_sweeps(thermalized=True)
        wk.fetch_counters(reset=True)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.carlo_steps):
            wk.sweep(1, thermalized=True, measure=False)
            wk.measure()
        barrier()
        dt1 = time.perf_counter() - t0
        c1 = wk.fetch_counters(reset=True)
        S = args.carlo_batch
        t0 = time.perf_counter()
        for k in range(max(1, args.carlo_steps // 2)):
            wk.sweep(S, thermalized=True, measure=True)
            wk.fetch_accumulators(reset=True)
        barrier()
        dt2 = time.perf_counter() - t0
        c2 = wk.fetch_counters(reset=True)
        carlo = {"per_sweep_launch": {"value": c1["visits"] / dt1, "unit": UNIT, "api": "sse_sweep(1) + sse_sync + sse_measure per Carlo step",
                                      "steps": args.carlo_steps, "ms_per_sweep": 1e3 * dt1 / args.carlo_steps},
                 "batched": {"value": c2["visits"] / dt2, "unit": UNIT, "sweeps_per_call": S,
                             "api": "sse_sweep(sweeps_per_call, measure=1) + sse_fetch_accumulators per Carlo step"}}

    # ---- aggregate over ranks: max time, summed work ----
    red = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([cnt["visits"], cnt["sweeps"], cnt["sum_n"], cnt["sum_M"], cnt2["visits"], cnt2["sweeps"], W],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_total, e2e_ms = red.tolist()
    visits, sweeps, sum_n, sum_M, visits2, sweeps2, W_all = tot.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        # roofline of the dominant (only) kernel on rank 0: algorithmic bytes per launch / mean launch time
        b_launch = algorithmic_bytes(cnt["sum_M"], cnt["sum_n"], cnt["visits"]) / args.steps
        avg_launch_ms = float(np.mean(launch_ms))
        achieved = b_launch / (avg_launch_ms * 1e-3) / 1e9
        rank_visits_per_s = cnt["visits"] / (ms_total * 1e-3)
        prof = committed_profile()
        traffic = issue_frac = None
        if prof and prof.get("L") == args.L:
            traffic = prof["dram_bytes_per_visit"] * cnt["visits"] / args.steps
            issue_frac = prof["warp_instructions_per_visit"] * rank_visits_per_s / (148 * 4 * 1.965e9)
        occupancy = cnt["lane_iters"] / max(1, 32 * cnt["warp_iters"])
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        cfg = config_block(args)
        cfg.update({
            "walkers_per_gpu": W, "visits_per_walker_per_step": B, "n_sites": args.L * args.L,
            "mean_n": sum_n / max(1, sweeps), "mean_M": sum_M / max(1, sweeps), "visits_per_sweep": visits / max(1, sweeps),
            "l2": "inputs larger than L2: per-GPU walker state %.1f GB >> 126 MB" % (wk.device_bytes() / 1e9),
            "bytes_per_walker": wk.device_bytes() / W,
            "parallelism": f"walkers sharded over {world} rank(s), no collective inside a sweep",
        })
        line = {
            "metric": METRIC, "value": visits / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": cfg,
            "sweeps_per_s": sweeps / (ms_total * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": (prof or {}).get("source") if traffic else None, "peak_source": peak_src,
                         "kernel": "sse::k_sweep<false>", "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": b_launch,
                         "formula": "12*M + 16*n + 64*V per walker-sweep (SURVEY.md 8d)",
                         # the two other ceilings of this latency-bound path
                         "chain_ceiling": {"value": chain_ceiling(W), "unit": UNIT, "frac": rank_visits_per_s / chain_ceiling(W),
                                           "what": "hops/s of W dependent 16-byte-load + 4-byte-store chains, one per lane (profiles/r2_chase_lanes.txt)"},
                         "issue_frac": issue_frac},
            "e2e": {"value": visits2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * W,
                    "d2h_bytes_per_step": 8 * (n_obs + 2), "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "sse_set_temperature + sse_advance(measure=1) + sse_reduce_bins (device sum over walkers, NCCL all-reduce "
                           "over ranks inside the library) per step",
                    "energy_per_site": energy},
            "carlo_call_pattern": carlo,
            "gpu_launches": args.steps,
            "kernel_stats": {
                "worm_lane_occupancy": occupancy,
                "worm_loop_ns_per_iteration": cnt["cycles_worm"] / max(1, cnt["warp_iters"]) / clk * 1e3,
                "stream_warps_busy_per_sm": (cnt["cycles_build"] + cnt["cycles_finish"]) / (ms_total * 1e-3 * clk * 1e6) / min(W, 148),
                "build_ms_per_walker_sweep": cnt["cycles_build"] / max(1, cnt["sweeps"]) / clk * 1e-3,
                "finish_ms_per_walker_sweep": cnt["cycles_finish"] / max(1, cnt["sweeps"]) / clk * 1e-3,
            },
            "clocks": clocks,
            "setup_s": t_setup,
        }
    del wk
    if rank == 0:
        if world == 1 and args.secondary and (args.L, args.beta) == (64, 64.0):
            line["secondary"] = secondary_config1(args, local_rank, stream)
        if world == 1 and not args.no_cpu:
            r, cores, sw, native = cpu_baseline(args)
            line["cpu_baseline"] = {
                "value": r["visits"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cores} walkers (one per host thread), thermalised like the device walkers, {sw} timed sweeps each, "
                          f"mean n={r['mean_n']:.0f}; C++ oracle, reference data layout, xoshiro256++, "
                          f"{'-march=native' if native else '-march=x86-64-v3'}",
                "per_core": r["visits"] / r["thread_seconds"], "host": host_cpu(),
            }
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def secondary_config1(args, dev_index, stream):
    """BASELINE.json configs[1] (L = beta = 32, 4096 walkers on one B200), short run of the same kernel."""
    import copy
    import torch

    a = copy.copy(args)
    a.L, a.beta, a.beta_doublings, a.therm, a.m_capacity, a.n_capacity = 32, 32.0, 5, 20, 0, 0
    wk, dm, T, W, t_setup = setup_walkers(a, dev_index, 0, 4096, stream.cuda_stream)
    B = 400000
    wk.advance(B, thermalized=True)
    wk.fetch_counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(3):
            wk.advance(B, thermalized=True, sync=False)
        e1.record(stream)
    torch.cuda.synchronize()
    wk.sync()
    c = wk.fetch_counters(reset=True)
    ms = e0.elapsed_time(e1)
    return {"config": config_block(a), "walkers_per_gpu": W, "value": c["visits"] / (ms * 1e-3), "unit": UNIT, "steps": 3,
            "ms_per_step": ms / 3, "sweeps_per_s": c["sweeps"] / (ms * 1e-3), "mean_n": c["sum_n"] / max(1, c["sweeps"]),
            "chain_ceiling": chain_ceiling(W), "setup_s": t_setup}


class StdoutGuard:
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to fd 1 when
    NCCL_DEBUG is set), so fd 1 is pointed at stderr while the benchmark runs and the line goes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


GUARD = None


def emit(line):
    if GUARD is not None:
        GUARD.emit(line)
    else:
        print(line)


def main():
    global GUARD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--beta", type=float, default=64.0)
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU; 0 = as many as the GPU's memory holds")
    ap.add_argument("--max-walkers", type=int, default=0)
    ap.add_argument("--visits-per-step", type=float, default=3.0e6,
                    help="worm visits every walker does per step (one sse_advance launch); ~3.5 sweeps at L = beta = 64")
    ap.add_argument("--worm-warps", type=int, default=0)
    ap.add_argument("--stream-warps", type=int, default=0)
    ap.add_argument("--therm", type=int, default=12, help="sweeps at the target temperature after the doubling levels")
    ap.add_argument("--beta-doublings", type=int, default=-1,
                    help="untimed setup: start 2^k times hotter and double beta k times (sse_double_beta); default log2(beta)")
    ap.add_argument("--therm-per-level", type=int, default=8,
                    help="sweeps per beta-doubling level (run with the controller attenuation 0.1, see Walkers.thermalize_by_beta_doubling)")
    ap.add_argument("--deterministic", action="store_true",
                    help="energy_offset_factor=0 tables (the reference's intended but unreachable S=1/2 branch)")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--m-capacity", type=int, default=0)
    ap.add_argument("--n-capacity", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the end-to-end leg (at most --steps)")
    ap.add_argument("--carlo-steps", type=int, default=2, help="steps of the Carlo call-pattern leg (0 = skip)")
    ap.add_argument("--carlo-batch", type=int, default=4)
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    ap.add_argument("--cpu-therm", type=int, default=12)
    ap.add_argument("--cpu-sweeps", type=int, default=0, help="timed sweeps per host thread of the cpu_baseline leg (0 = ~15 s worth)")
    ap.add_argument("--cpu-sweeps-per-step", type=int, default=0, help="--impl reference: sweeps per step and thread (0 = ~3 s worth)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.beta_doublings < 0:
        args.beta_doublings = max(0, int(round(math.log2(args.beta))))
    # one CPU walker-sweep costs ~ n * 2.1 visits at ~1e7 visits/s
    sweep_s = 0.71 * args.beta * 2 * args.L * args.L * 2.1 / 1.0e7
    if args.cpu_sweeps <= 0:
        args.cpu_sweeps = max(8, int(15.0 / sweep_s))
    if args.cpu_sweeps_per_step <= 0:
        args.cpu_sweeps_per_step = max(2, int(3.0 / sweep_s))
    GUARD = StdoutGuard()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
