#!/usr/bin/env python
"""BASELINE.json configs[4]: bilayer Heisenberg model in the dimer basis (ClusterModel, the reference's cluster job
test/test_jobs.jl:139-167 scaled to L), temperature sweep with the walkers of every temperature spread over all ranks.

  python profiles/tools/temperature_sweep.py --L 48 --beta-max 48 --n-T 8 --replicas 256           (one GPU)
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/tools/temperature_sweep.py ...   (one rank per GPU)

Every rank owns `replicas` walkers per temperature (global stream ids rank*W + i, no data-path collective).  A bin is
`--binsize` sweeps with the estimators accumulated on the device; bins are summed per temperature over the rank's walkers
and over all ranks INSIDE the library (sse_reduce_bins: device sum + NCCL all-reduce).  With --exchange the replicas of a
rank form temperature ladders and neighbour swaps are decided on the device between bins (sse_pt_exchange).
Rank 0 prints one JSON line: energy per temperature (jackknife over bins), throughput, and — with --oracle-check — the same
energies from the CPU oracle (reference-layout restatement) with their z-scores.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def jackknife_ratio(num, den):
    """mean and error of <num>/<den> from bins (Carlo's evaluables do the same for Energy = SignEnergy / Sign)."""
    num, den = np.asarray(num, float), np.asarray(den, float)
    n = len(num)
    full = num.sum() / den.sum()
    if n < 2:
        return full, float("nan")
    jk = (num.sum() - num) / (den.sum() - den)
    return float(n * full - (n - 1) * jk.mean()), float(np.sqrt((n - 1) / n * ((jk - jk.mean()) ** 2).sum()))


def run(args):
    import sse_b200 as S
    from sse_b200.walkers import DeviceModel, Walkers

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        try:
            import torch

            torch.cuda.set_device(local_rank)
        except Exception:
            pass

    from helpers import dimer_bilayer

    model = dimer_bilayer(args.L)
    dm = DeviceModel(model=model)
    Ts = np.geomspace(1.0 / args.beta_max, args.T_max, args.n_T)
    nT, R = len(Ts), args.replicas
    W = nT * R
    Tw = np.tile(Ts, R)  # walker i: temperature i % nT, replica i // nT (a replica's walkers are a ladder)
    group = (np.arange(W) % nT).astype(np.int32)
    n_bonds = 2 * args.L * args.L
    n_est = 2.9 * args.beta_max * n_bonds  # ~2.7 operators per bond and unit of beta (oracle, L = 6, 8)
    wk = Walkers(dm, Tw, m_capacity=int(4 * n_est) + 8192, n_capacity=int(1.25 * n_est) + 2048, seed=args.seed,
                 walker_id_offset=rank * W, device=local_rank)
    if world > 1:
        box = [Walkers.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        wk.comm_init(box[0], rank, world)
    t0 = time.time()
    wk.thermalize_by_beta_doubling(args.doublings, sweeps_per_level=args.per_level)
    wk.sweep(args.therm, thermalized=False)
    t_setup = time.time() - t0
    ladders = None
    if args.exchange:
        from sse_b200.tempering import DeviceReplicaExchange

        ladders = DeviceReplicaExchange(wk, seed=args.seed + 1)  # one ladder over all walkers sorted by temperature
    wk.fetch_counters(reset=True)
    wk.fetch_accumulators(reset=True)
    bins_s, bins_c = [], []
    t0 = time.time()
    for b in range(args.bins):
        wk.sweep(args.binsize, thermalized=True, measure=True)
        if ladders is not None:  # temperatures may have moved: a bin belongs to the temperature RANK, R walkers per rank
            g = (ladders.rank_of_walker() // R).astype(np.int32)
        else:
            g = group
        s, c = wk.reduce_bins(g, nT, reset=True)
        bins_s.append(s)
        bins_c.append(c)
        if ladders is not None:
            ladders.step()
            ladders.step()
    seconds = time.time() - t0
    cnt = wk.fetch_counters(reset=True)
    tot = np.array([cnt["visits"], cnt["sweeps"]], dtype=np.float64)
    if world > 1:
        import torch

        t = torch.tensor(tot, device=torch.device("cuda", local_rank))
        dist.all_reduce(t)
        tot = t.cpu().numpy()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    bins_s, bins_c = np.array(bins_s), np.array(bins_c)  # [bins, nT, n_obs], [bins, nT, 2]
    energy = [jackknife_ratio(bins_s[:, k, 4], bins_s[:, k, 0]) for k in range(nT)]
    nops = [jackknife_ratio(bins_s[:, k, 1], bins_c[:, k, 0]) for k in range(nT)]
    line = {"workload": f"fully frustrated bilayer, dimer basis (ClusterModel), L={args.L}, beta up to {args.beta_max} (BASELINE.json configs[4])",
            "n_gpus": world, "walkers_per_gpu": W, "temperatures": Ts.tolist(), "replicas_per_temperature": R * world,
            "bins": args.bins, "binsize": args.binsize, "exchange": bool(args.exchange),
            "energy": [e[0] for e in energy], "energy_error": [e[1] for e in energy],
            "operator_count": [x[0] for x in nops],
            "visits_per_s": tot[0] / seconds, "sweeps_per_s": tot[1] / seconds, "seconds": seconds, "setup_s": t_setup,
            "pt_accept": (ladders.accepted / max(1, ladders.proposed)) if ladders is not None else None}
    if args.oracle_check:
        from concurrent.futures import ThreadPoolExecutor

        from oracle import OracleModel, OracleWalker

        om = OracleModel(model)

        def one(k):
            T = float(Ts[k])
            ow = OracleWalker(om, T * 2 ** args.doublings, seed=args.seed + 77, walker_id=k, num_worms_attenuation_factor=0.1)
            ow.init()
            for _ in range(args.doublings):
                ow.sweep(args.per_level)
                st = ow.get_state()
                st["operators"] = np.concatenate([st["operators"], st["operators"]])
                st["num_operators"] *= 2
                st["T"] /= 2.0
                st["avg_worm_length"] *= 2.0
                ow.set_state(st)
            fw = OracleWalker(om, T, seed=args.seed + 77, walker_id=k)
            fw.set_state(ow.get_state())
            fw.sweep(args.therm)
            num, den = [], []
            for _ in range(args.oracle_bins):
                fw.sweep(args.oracle_binsize, thermalized=True, measure=True)
                s, c = fw.fetch_accumulators()
                num.append(s[4])
                den.append(s[0])
            return jackknife_ratio(num, den)

        ks = list(range(0, nT, max(1, nT // args.oracle_points)))
        with ThreadPoolExecutor(len(ks)) as ex:
            ref = list(ex.map(one, ks))
        z = [(energy[k][0] - r[0]) / np.hypot(energy[k][1], r[1]) for k, r in zip(ks, ref)]
        line["oracle"] = {"temperature_index": ks, "energy": [r[0] for r in ref], "energy_error": [r[1] for r in ref], "z": z,
                          "max_abs_z": float(np.max(np.abs(z)))}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=48)
    ap.add_argument("--beta-max", type=float, default=48.0)
    ap.add_argument("--T-max", type=float, default=2.0)
    ap.add_argument("--n-T", type=int, default=8)
    ap.add_argument("--replicas", type=int, default=256, help="walkers per temperature and GPU")
    ap.add_argument("--doublings", type=int, default=4)
    ap.add_argument("--per-level", type=int, default=6)
    ap.add_argument("--therm", type=int, default=20)
    ap.add_argument("--bins", type=int, default=10)
    ap.add_argument("--binsize", type=int, default=4)
    ap.add_argument("--exchange", action="store_true")
    ap.add_argument("--oracle-check", action="store_true")
    ap.add_argument("--oracle-points", type=int, default=4)
    ap.add_argument("--oracle-bins", type=int, default=10)
    ap.add_argument("--oracle-binsize", type=int, default=4)
    ap.add_argument("--seed", type=int, default=4848)
    run(ap.parse_args())


if __name__ == "__main__":
    main()
