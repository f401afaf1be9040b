"""Run a task on the CPU oracle or the GPU walkers and evaluate it like Carlo would (bins + jackknife)."""
import numpy as np

from sse_b200.carlo import Evaluator
from sse_b200.mc import MC
from sse_b200.walkers import OBS_FIXED, OBS_PER_EST


def obs_names(model):
    names = list(OBS_FIXED)
    for e in model.get_opstring_estimators():
        names += [f"Sign{e.prefix}{o}" for o in OBS_PER_EST]
    return names


def evaluate(model, bins: dict):
    ev = Evaluator(bins)
    MC.register_evaluables(ev, {}, model)
    return ev.results


def run_oracle_task(om, model, T, sweeps, therm, binsize, seed=1, walker_id=0):
    from oracle import OracleWalker

    w = OracleWalker(om, T, seed=seed, walker_id=walker_id)
    w.init()
    w.sweep(therm, thermalized=False)
    names = obs_names(model)
    bins = {n: [] for n in names}
    for _ in range(sweeps // binsize):
        w.sweep(binsize, thermalized=True, measure=True)
        sums, counts = w.fetch_accumulators(reset=True)
        for i, n in enumerate(names):
            c = counts[1] if n == "WormLengthFraction" else counts[0]
            if c > 0:
                bins[n].append(sums[i] / c)
    return evaluate(model, {k: np.array(v) for k, v in bins.items() if len(v)})


def run_gpu_tasks(dm, model, Ts, sweeps, therm, binsize, seed=1, m_capacity=None, replicas=1, doublings=0):
    """All temperatures at once: walker i*replicas+r runs T[i]; bins of the replicas are pooled.
    doublings > 0: the walkers start 2**doublings times hotter and are grown by beta doubling before the `therm` sweeps at
    their temperature (no cold-start transient with its 1e8-visit worms, so no seed needs screening)."""
    from sse_b200.mc import default_capacity
    from sse_b200.walkers import Walkers

    Ts = np.asarray(Ts, dtype=np.float64)
    Tw = np.repeat(Ts, replicas)
    m_def, n_def = default_capacity(dm.sse_data, float(Ts.min()))
    if doublings > 0:
        m_def *= 2  # a doubled string is twice as long as the hotter walker's (slots cost 0.25 B)
    gw = Walkers(dm, Tw, m_capacity=m_capacity or m_def, n_capacity=n_def if not m_capacity else None, seed=seed)
    if doublings > 0:
        gw.thermalize_by_beta_doubling(doublings, sweeps_per_level=max(10, therm // 10))
    else:
        gw.init()
    done = 0
    while done < therm:
        k = min(500, therm - done)
        gw.sweep(k, thermalized=False)
        done += k
    names = obs_names(model)
    bins = {n: [] for n in names}
    for _ in range(sweeps // binsize):
        gw.sweep(binsize, thermalized=True, measure=True)
        sums, counts = gw.fetch_accumulators(reset=True)
        for i, n in enumerate(names):
            c = counts[:, 1] if n == "WormLengthFraction" else counts[:, 0]
            bins[n].append(sums[:, i] / np.maximum(c, 1))
    out = []
    for it in range(len(Ts)):
        sl = slice(it * replicas, (it + 1) * replicas)
        b = {k: np.array(v)[:, sl].reshape(-1) for k, v in bins.items()}
        out.append(evaluate(model, b))
    return out
