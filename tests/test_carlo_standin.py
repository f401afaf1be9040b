"""CPU tests of the Carlo stand-in: binning and jackknife evaluation (carlo.py), evaluables (sse.jl:111-134)."""
import numpy as np

from sse_b200.carlo import Evaluator, MCContext


def test_mccontext_binning_and_thermalisation():
    ctx = MCContext(dict(thermalization=3, binsize=4))
    assert not ctx.is_thermalized()
    ctx.sweeps = 4
    assert ctx.is_thermalized()
    for i in range(10):
        ctx.measure("X", np.array([i, 2.0 * i]))
    b = ctx.bin_array("X")
    assert b.shape == (2, 2)
    np.testing.assert_allclose(b[0], [1.5, 3.0])
    np.testing.assert_allclose(b[1], [5.5, 11.0])


def test_jackknife_ratio_matches_analytic_error():
    rng = np.random.default_rng(0)
    n = 400
    s = 1.0 + 0.05 * rng.standard_normal(n)
    se = -0.7 * s + 0.01 * rng.standard_normal(n)
    ev = Evaluator({"Sign": s, "SignEnergy": se})
    ev.evaluate("Energy", lambda a, b: a / b, ("SignEnergy", "Sign"))
    mean, err = ev["Energy"]
    assert abs(mean + 0.7) < 4 * err
    # ratio of correlated quantities: the jackknife error reflects only the uncorrelated part (0.01 / sqrt(n))
    assert 0.5 * 0.01 / np.sqrt(n) < err < 2.0 * 0.01 / np.sqrt(n)
    m, e = ev["Sign"]
    assert abs(e - s.std(ddof=1) / np.sqrt(n)) < 1e-12


def test_register_evaluables_names():
    import sse_b200 as S
    from sse_b200.mc import MC

    model = S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(2, 2)), J=1.0,
                               measure=["magnetization", "staggered_magnetization"]))
    rng = np.random.default_rng(1)
    names = ["Sign", "SignEnergy", "SignOperatorCount", "SignOperatorCount2"] + [
        f"Sign{p}{o}" for p in ("", "Stag") for o in ("Mag", "AbsMag", "Mag2", "Mag4", "MagChi")]
    bins = {k: 1.0 + 0.01 * rng.standard_normal(50) for k in names}
    ev = Evaluator(bins)
    MC.register_evaluables(ev, {}, model)
    for k in ("Energy", "SpecificHeat", "Mag", "AbsMag", "Mag2", "Mag4", "MagChi", "BinderRatio", "StagMag",
              "StagBinderRatio", "StagMagChi"):
        assert k in ev.results and np.isfinite(ev.results[k][0])


def test_replica_exchange_decisions_satisfy_detailed_balance():
    """The swap rule built on the reference's log weight ratio (sse.jl:395): for two walkers with operator counts
    (n_a, n_b) the acceptance probabilities of a swap and of its reverse differ by exactly the SSE weight ratio
    (T_b/T_a)^(n_b - n_a)... i.e. p(swap)/p(reverse) = exp(lw)."""
    from sse_b200.tempering import log_weight_ratio, swap_decisions

    Ta, Tb, na, nb = 0.5, 0.7, 120, 95
    lw = log_weight_ratio(na, Ta, Tb) + log_weight_ratio(nb, Tb, Ta)
    assert lw == -na * np.log(Tb / Ta) - nb * np.log(Ta / Tb)
    p_fwd = min(1.0, np.exp(lw))
    p_rev = min(1.0, np.exp(-lw))
    assert p_fwd / p_rev == np.exp(lw) or np.isclose(p_fwd / p_rev, np.exp(lw))
    # deterministic behaviour of the pairing: parity 0 proposes (0,1), (2,3); parity 1 proposes (1,2)
    T = np.array([0.1, 0.2, 0.3, 0.4])
    n = np.array([100, 100, 100, 100])  # equal counts: lw = 0 -> every proposal with u < 1 is accepted
    out = swap_decisions(n, T, np.argsort(T), 0, np.array([0.5, 0.5]))
    assert out.tolist() == [0.2, 0.1, 0.4, 0.3]
    out = swap_decisions(n, T, np.argsort(T), 1, np.array([0.5]))
    assert out.tolist() == [0.1, 0.3, 0.2, 0.4]
    # a swap that would move the colder temperature onto the longer string is always accepted, the reverse rarely
    n2 = np.array([50, 150])
    T2 = np.array([0.1, 0.2])  # walker 1 (hot) has MORE operators than walker 0 (cold): swap is favourable
    assert swap_decisions(n2, T2, [0, 1], 0, np.array([0.999])).tolist() == [0.2, 0.1]
    assert swap_decisions(n2[::-1], T2, [0, 1], 0, np.array([0.5])).tolist() == [0.1, 0.2]
