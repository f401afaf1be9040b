#!/usr/bin/env python
"""bench.py — throughput of the SSE sweep hot path (diagonal update -> vertex records -> worm update).

Metric (BASELINE.json / SURVEY.md §8d): worm operator-vertex visits per second = sum of the lengths
returned by worm_traverse! (src/sse.jl:302) over all walkers and sweeps / time of the WHOLE sweep.
Workload at every N: BASELINE.json configs[1], 2D square-lattice S=1/2 Heisenberg AFM L=32, beta=32,
4096 walkers per GPU (weak scaling: walkers shard over ranks, no data-path collective).

A "step" is one persistent launch advancing every walker by --sweeps-per-step full sweeps.

  python bench.py --gpus 1 --steps K --warmup W            (our arm)
  torchrun ... bench.py --gpus N ...                        (one rank per GPU, NCCL only for bin reduction)
  python bench.py --impl reference ...                      (CPU oracle on all host cores, same metric)
"""
import argparse
import json
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


def build_params(args, n_walkers, walker_id_offset=0, device=-1):
    import sse_b200 as S

    L = args.L
    T = 1.0 / args.beta
    n_bonds = 2 * L * L
    # expected n ~ beta * N_b * (|e_bond| + offset) ~ 0.71 * beta * N_b for the eof=0.25 tables
    n_est = 0.75 * args.beta * n_bonds
    m_cap = int(args.m_capacity or 3.6 * n_est)
    n_cap = int(args.n_capacity or 1.7 * n_est)
    return dict(
        model=S.MagnetModel,
        lattice=dict(unitcell=S.UnitCells.square, size=(L, L)),
        J=1.0,
        s_half_deterministic=bool(args.deterministic),
        measure=["magnetization", "staggered_magnetization"],
        T=T,
        n_walkers=n_walkers,
        seed=args.seed,
        walker_id_offset=walker_id_offset,
        device=device,
        m_capacity=m_cap,
        n_capacity=n_cap,
    )


def workload_name(args):
    cfg = {(32, 32.0): "BASELINE.json configs[1]", (64, 64.0): "BASELINE.json configs[2]"}.get((args.L, args.beta), "custom size")
    return f"2D square-lattice S=1/2 Heisenberg AFM L={args.L}, beta={args.beta}, {args.walkers} walkers per GPU ({cfg})"


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


def measured_traffic_per_walker_sweep():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per walker-sweep from the committed ncu --set full
    capture of this kernel on this workload (profiles/r1_d_dram_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_d_dram_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_walker_sweep"]), d["source"]
    except Exception:
        return None, None


def algorithmic_bytes(sum_M, sum_n, visits, measured_sweeps=0):
    """B_sweep = 8M (K1 read+write op codes) + 4M + 16n (K2) + 64V (K3) (+4M per measured sweep)."""
    return 8.0 * sum_M + 4.0 * sum_M + 16.0 * sum_n + 64.0 * visits + 4.0 * measured_sweeps


def cpu_baseline(args, cores=None, therm=None, sweeps=None):
    """The CPU oracle (reference data layout, xoshiro256++ stream) on the host cores: one independent walker
    per thread, as Carlo runs one MC per MPI rank (docs/src/tutorial.md:49)."""
    import sse_b200  # noqa: F401
    from oracle import OracleModel
    import oracle as oracle_mod

    oracle_mod.build()
    p = build_params(args, 1)
    model = p["model"](p)
    om = OracleModel(model)
    cores = cores or os.cpu_count() or 1
    therm = args.cpu_therm if therm is None else therm
    sweeps = args.cpu_sweeps if sweeps is None else sweeps
    r = om.bench(p["T"], cores, therm, sweeps, seed=args.seed)
    return r, cores, therm, sweeps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s = args.cpu_sweeps_per_step
    r, cores, therm, sweeps = cpu_baseline(args, therm=args.cpu_therm + args.warmup * s, sweeps=args.steps * s)
    value = r["visits"] / r["seconds"]
    sample = (f"{cores} independent walkers (one per host thread), {therm} thermalisation sweeps untimed, then "
              f"{sweeps} timed sweeps each = {args.steps} steps x {s} sweeps; mean n={r['mean_n']:.0f}, M={r['mean_M']:.0f}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64/f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "reference CPU path = C++ oracle restating src/sse.jl "
                   "(julia is not installed in this image); per-walker data layout of the reference",
                   "energy_offset_factor": 0.0 if args.deterministic else 0.25},
        "sweeps_per_s": r["walker_sweeps"] / r["seconds"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class _DevArr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr, False), version=2)


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

    from sse_b200.mc import MC

    W = args.walkers
    S = args.sweeps_per_step
    params = build_params(args, W, walker_id_offset=rank * W, device=local_rank)
    t_setup = time.time()
    mc = MC(params)
    wk = mc.walkers
    stream = torch.cuda.Stream(device=dev)
    wk.set_stream(stream.cuda_stream)
    if args.walkers_per_warp != 1:
        wk.set_walkers_per_warp(args.walkers_per_warp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # thermalise (untimed): init! + un-thermalised sweeps (string growth, worm-count controller)
    if args.beta_doublings > 0:
        # large systems: grow the cold walkers from hot ones (sse_double_beta), then thermalise at the target
        wk.thermalize_by_beta_doubling(args.beta_doublings, sweeps_per_level=args.therm_per_level)
    else:
        wk.init()
    done = 0
    while done < args.therm:
        k = min(50, args.therm - done)
        wk.sweep(k, thermalized=False, measure=False)
        done += k
    for _ in range(args.warmup):
        wk.sweep(S, thermalized=True, measure=False)
    t_setup = time.time() - t_setup

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
            wk.sweep(S, thermalized=True, measure=False, sync=False)
            events[k + 1].record(stream)
    barrier()
    wk.sync()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = events[0].elapsed_time(events[-1])
    launch_ms = [events[k].elapsed_time(events[k + 1]) for k in range(args.steps)]
    cnt = wk.fetch_counters(reset=True)

    # ---- timed region 2: end to end through the public API with host buffers ----
    n_obs = wk.n_obs
    T_host = torch.full((W,), params["T"], dtype=torch.float64).pin_memory()
    T_np = T_host.numpy()
    sptr, cptr = wk.accumulators_device_ptr()
    acc_t = torch.as_tensor(_DevArr(sptr, (W, n_obs), "<f8"), device=dev)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        wk.set_temperature(T_np)                                  # H2D: this step's parameters
        wk.sweep(S, thermalized=True, measure=True, sync=False)   # sweeps + on-device estimators
        if world > 1:                                             # NCCL: reduce the binned observables only
            with torch.cuda.stream(stream):
                bin_sum = acc_t.sum(dim=0)
                dist.all_reduce(bin_sum)
        sums, counts = wk.fetch_accumulators(reset=True)          # D2H: the bin
    barrier()
    e2e_s = time.perf_counter() - t0
    cnt2 = wk.fetch_counters(reset=True)
    energy = float(sums[:, 4].sum() / sums[:, 0].sum())

    # ---- aggregate over ranks: max time, summed work ----
    red = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([cnt["visits"], cnt["sweeps"], cnt["sum_n"], cnt["sum_M"], cnt2["visits"], cnt2["sweeps"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_total, e2e_ms = red.tolist()
    visits, sweeps, sum_n, sum_M, visits2, sweeps2 = tot.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        # roofline of the dominant (only) kernel on rank 0: algorithmic bytes per launch / mean launch time
        b_launch = algorithmic_bytes(cnt["sum_M"], cnt["sum_n"], cnt["visits"]) / args.steps
        avg_launch_ms = float(np.mean(launch_ms))
        achieved = b_launch / (avg_launch_ms * 1e-3) / 1e9
        tpws, tsrc = measured_traffic_per_walker_sweep()
        default_workload = (args.L == 32 and args.beta == 32.0 and not args.deterministic)
        traffic = tpws * cnt["sweeps"] / args.steps if (tpws and default_workload) else None
        line = {
            "metric": METRIC, "value": visits / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "walkers_per_gpu": W, "walkers_per_warp": args.walkers_per_warp,
                "sweeps_per_step": S,
                "thermalisation_sweeps": args.therm, "beta_doublings": args.beta_doublings, "energy_offset_factor": 0.0 if args.deterministic else 0.25,
                "mean_n": sum_n / sweeps, "mean_M": sum_M / sweeps, "visits_per_sweep": visits / sweeps,
                "l2": "inputs larger than L2: per-GPU walker state %.1f GB >> 126 MB" % (wk.device_bytes() / 1e9),
                "parallelism": f"walkers sharded over {world} rank(s), no collective inside a sweep",
            },
            "sweeps_per_s": sweeps / (ms_total * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": tsrc if traffic else None, "peak_source": peak_src,
                         "kernel": "sse::k_walkers<false>" if args.walkers_per_warp == 1 else "sse::k_walkers_multi<false,%d>" % args.walkers_per_warp,
                         "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": b_launch,
                         "formula": "12*M + 16*n + 64*V per walker-sweep (SURVEY.md 8d)"},
            "e2e": {"value": visits2 / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * W,
                    "d2h_bytes_per_step": 8 * W * n_obs + 16 * W, "ms_per_step": e2e_ms / args.steps,
                    "api": "sse_set_temperature + sse_sweep(measure=1) + sse_fetch_accumulators per step",
                    "energy_per_site": energy},
            "gpu_launches": args.steps,
            "phase_cycle_share": {k: cnt[k] / max(1, cnt["cycles_diag_build"] + cnt["cycles_worm"] + cnt["cycles_commit_measure"])
                                  for k in ("cycles_diag_build", "cycles_worm", "cycles_commit_measure")},
            "worm_cycles_per_visit": cnt["cycles_worm"] / max(1, cnt["visits"]),
            "clocks": clocks,
            "setup_s": t_setup,
        }
        if world == 1 and not args.no_cpu:
            r, cores, therm, sw = cpu_baseline(args)
            line["cpu_baseline"] = {
                "value": r["visits"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cores} walkers (one per host thread), {therm} thermalisation + {sw} timed sweeps each, "
                          f"mean n={r['mean_n']:.0f}; C++ oracle, reference data layout, xoshiro256++",
                "per_core": r["visits"] / r["thread_seconds"],
            }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=32)
    ap.add_argument("--beta", type=float, default=32.0)
    ap.add_argument("--walkers", type=int, default=4096, help="walkers per GPU")
    ap.add_argument("--walkers-per-warp", type=int, default=1, choices=[1, 2, 4],
                    help="launch shape (sse_set_walkers_per_warp): 2 or 4 interleave the worm updates of a warp's walkers; "
                         "meant for --walkers well beyond 4144 (e.g. --walkers 8192 --walkers-per-warp 2)")
    ap.add_argument("--sweeps-per-step", type=int, default=64,
                    help="sweeps per launch (one Carlo bin; the reference tutorial uses binsize 100). Longer launches "
                         "average the per-walker worm-length imbalance: busy fraction 81 %% at 32, 86 %% at 100")
    ap.add_argument("--therm", type=int, default=300)
    ap.add_argument("--beta-doublings", type=int, default=0,
                    help="untimed setup: start 2^k times hotter and double beta k times (sse_double_beta) before the "
                         "--therm sweeps at the target; for L=64, beta=64 use 6")
    ap.add_argument("--therm-per-level", type=int, default=10,
                    help="sweeps per beta-doubling level (run with the controller attenuation 0.1, see Walkers.thermalize_by_beta_doubling)")
    ap.add_argument("--deterministic", action="store_true",
                    help="energy_offset_factor=0 tables (the reference's intended but unreachable S=1/2 branch)")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--m-capacity", type=int, default=0)
    ap.add_argument("--n-capacity", type=int, default=0)
    ap.add_argument("--cpu-therm", type=int, default=300)
    ap.add_argument("--cpu-sweeps", type=int, default=400)
    ap.add_argument("--cpu-sweeps-per-step", type=int, default=80)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
