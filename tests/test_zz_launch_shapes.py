"""The persistent sweep kernel (sse::k_sweep): worm warps (one lane = one walker) + stream warps (one warp = one walker).
Whatever the launch shape — one walker per lane, several walkers per lane, one CTA or many — and wherever sse_advance
parks a walker (between sweeps, between worms, in the middle of a worm), every walker must stay on exactly the trajectory
of the oracle.  CPU: through the warp emulator (tests/emu); GPU: `-m gpu`."""
import numpy as np
import pytest

import test_gpu_parity as G
from helpers import MODEL_CLASSES, random_stream
from oracle import OracleWalker
from sse_b200.capi import SSEError
from sse_b200.walkers import Walkers
from test_emu_parity import emu, emu_built  # noqa: F401  (fixtures)

SHAPES = [(1, 1), (2, 3), (1, 8)]


def _body_injected_sweeps(shape):
    """sse_sweep under an injected stream (the dbg_* parity hooks run single phases, so this is the injected-stream
    coverage of the persistent kernel), walkers of different lengths."""
    model = MODEL_CLASSES["spin1_dz"]()
    dm, om = G._pair(model)
    W = 7
    Ts = np.linspace(0.25, 1.0, W)
    rng = np.random.default_rng(5)
    gw = Walkers(dm, Ts, m_capacity=4096, seed=3)
    gw.set_launch_shape(*shape)
    gw.init()
    gw.sweep(10)
    ows = []
    for i in range(W):
        ow = OracleWalker(om, float(Ts[i]), seed=3, walker_id=i)
        ow.init()
        ow.sweep(10)
        G._same_state(gw.get_state(i), ow.get_state(), f"start walker {i}")
        ows.append(ow)
    streams = np.stack([random_stream(rng, 400000) for _ in range(W)])
    gw.set_injected_stream(streams)
    gw.sweep(4, thermalized=False)
    gw.sweep(2, thermalized=True, measure=True)
    sums, counts = gw.fetch_accumulators()
    for i, ow in enumerate(ows):
        ow.set_injected_stream(streams[i])
        ow.sweep(4, thermalized=False)
        ow.sweep(2, thermalized=True, measure=True)
        assert not ow.stream_exhausted
        G._same_state(gw.get_state(i), ow.get_state(), f"walker {i}")
        osums, ocounts = ow.fetch_accumulators()
        assert np.array_equal(counts[i], ocounts)
        np.testing.assert_allclose(sums[i], osums, rtol=1e-12, atol=1e-300)
    # a stream that runs out inside the worm phase is reported, not recycled
    gw.set_injected_stream(streams[:, :300])
    with pytest.raises(SSEError):
        gw.sweep(3)


def _body_switching_keeps_the_trajectory():
    """Changing the launch shape between launches does not change the chain."""
    model = MODEL_CLASSES["heisenberg_eof"]()
    dm, om = G._pair(model)
    Ts = np.linspace(0.2, 0.9, 9)
    a = Walkers(dm, Ts, m_capacity=4096, seed=12)
    b = Walkers(dm, Ts, m_capacity=4096, seed=12)
    a.init()
    b.init()
    for shape in ((1, 1), (3, 2), (0, 0), (1, 15), (8, 8)):
        b.set_launch_shape(*shape)
        a.sweep(4)
        b.sweep(4)
    for i in range(len(Ts)):
        G._same_state(a.get_state(i), b.get_state(i), f"walker {i}")
    with pytest.raises(SSEError):
        b.set_launch_shape(12, 30)


def _body_many_walkers_per_lane(monkeypatch):
    """More walkers in a CTA than worm lanes: a lane serves its walkers in turn (one CTA, one worm warp, 70 walkers)."""
    monkeypatch.setenv("SSE_B200_CTAS", "1")
    model = MODEL_CLASSES["heisenberg_eof"]()
    dm, om = G._pair(model)
    W = 70
    Ts = np.linspace(0.3, 2.0, W)
    gw = Walkers(dm, Ts, m_capacity=4096, seed=17)
    gw.set_launch_shape(1, 3)
    gw.init()
    gw.sweep(6, thermalized=False)
    gw.sweep(3, thermalized=True, measure=True)
    sums, counts = gw.fetch_accumulators()
    for i in (0, 31, 32, 63, 64, 69):
        ow = OracleWalker(om, float(Ts[i]), seed=17, walker_id=i)
        ow.init()
        ow.sweep(6, thermalized=False)
        ow.sweep(3, thermalized=True, measure=True)
        G._same_state(gw.get_state(i), ow.get_state(), f"walker {i}")
        osums, ocounts = ow.fetch_accumulators()
        assert np.array_equal(counts[i], ocounts)
        np.testing.assert_allclose(sums[i], osums, rtol=1e-12, atol=1e-300)
    assert gw.fetch_counters()["sweeps"] == W * 9


def _body_advance_parks_anywhere(name, budgets):
    """sse_advance: a fixed number of worm visits per walker and launch; walkers are parked between sweeps, between two
    worms or inside a worm and resume there.  After finish_sweeps every walker sits on the oracle's trajectory after
    exactly as many sweeps as it reports."""
    model = MODEL_CLASSES[name]()
    dm, om = G._pair(model)
    W = 6
    Ts = np.linspace(0.2, 1.2, W)
    gw = Walkers(dm, Ts, m_capacity=8192, seed=23)
    gw.init()
    visits = 0
    for b in budgets:
        gw.advance(b, thermalized=False)
    done, in_flight = gw.progress()
    assert in_flight.any()  # with these budgets somebody is parked inside a sweep
    with pytest.raises(SSEError):
        gw.get_state(int(np.nonzero(in_flight)[0][0]))  # ... and says so instead of returning a half-updated configuration
    gw.finish_sweeps(thermalized=False)
    done2, in_flight2 = gw.progress()
    assert not in_flight2.any()
    assert np.array_equal(done2, done + in_flight)
    c = gw.fetch_counters()
    assert c["sweeps"] == int(done2.sum())
    for i in range(W):
        ow = OracleWalker(om, float(Ts[i]), seed=23, walker_id=i)
        ow.init()
        ow.sweep(int(done2[i]), thermalized=False)
        G._same_state(gw.get_state(i), ow.get_state(), f"{name} walker {i} after {done2[i]} sweeps")
        visits += ow.visits if hasattr(ow, "visits") else 0
    # a launch gives every walker the same number of visits: nobody did more than the budgets allow, and walkers with
    # work left used them up
    assert c["visits"] <= W * (sum(budgets) + max(budgets)) + 10**6
    # sweep() afterwards keeps walkers in step again
    gw.sweep(2, thermalized=True, measure=True)
    done3, _ = gw.progress()
    assert np.array_equal(done3, done2 + 2)
    # max_sweeps bounds an advance as well
    gw.advance(10**9, max_sweeps=1, thermalized=True)
    done4, fl4 = gw.progress()
    assert np.array_equal(done4, done3 + 1) and not fl4.any()


@pytest.mark.parametrize("shape", SHAPES)
def test_emu_injected_sweeps(emu, shape):
    _body_injected_sweeps(shape)


def test_emu_switching_keeps_the_trajectory(emu):
    _body_switching_keeps_the_trajectory()


def test_emu_many_walkers_per_lane(emu, monkeypatch):
    _body_many_walkers_per_lane(monkeypatch)


@pytest.mark.parametrize("name, budgets", [("heisenberg_eof", [7, 50, 3, 200, 1, 90]), ("spin1_dz", [40, 11, 300, 5])])
def test_emu_advance_parks_anywhere(emu, name, budgets):
    _body_advance_parks_anywhere(name, budgets)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_gpu_injected_sweeps(shape):
    _body_injected_sweeps(shape)


@pytest.mark.gpu
def test_gpu_switching_keeps_the_trajectory():
    _body_switching_keeps_the_trajectory()


@pytest.mark.gpu
def test_gpu_many_walkers_per_lane(monkeypatch):
    _body_many_walkers_per_lane(monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("name, budgets", [("heisenberg_eof", [7, 50, 3, 200, 1, 90]), ("spin1_dz", [40, 11, 300, 5]),
                                           ("dimer_bilayer", [1000, 17, 3000])])
def test_gpu_advance_parks_anywhere(name, budgets):
    _body_advance_parks_anywhere(name, budgets)


@pytest.fixture(params=[(1, 2), (4, 12)])
def shape_env(request, monkeypatch):
    monkeypatch.setenv("SSE_B200_WORM_WARPS", str(request.param[0]))
    monkeypatch.setenv("SSE_B200_STREAM_WARPS", str(request.param[1]))
    return request.param


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["heisenberg_eof", "spin1_dz", "dimer_bilayer"])
def test_gpu_shapes_sweep_parity_philox(shape_env, name):
    G.test_sweep_parity_philox(name)


@pytest.mark.gpu
def test_gpu_shapes_edge_cases(shape_env):
    G.test_edge_cases_empty_and_ragged_strings()


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1])
def test_gpu_shapes_large_lattice_memory_paths(shape_env, level, monkeypatch):
    G.test_large_lattice_memory_paths(level, monkeypatch)


def _body_reduce_bins(with_comm):
    """sse_reduce_bins: per-group sums of the accumulators in walker order (and, with a communicator, the NCCL all-reduce
    over its ranks: one rank here) equal the host-side sums of sse_fetch_accumulators, and reset the bin."""
    model = MODEL_CLASSES["heisenberg_eof"]()
    dm, om = G._pair(model)
    W = 13
    Ts = np.repeat([0.4, 0.8, 1.6, 3.2], 4)[:W]
    group = np.repeat([0, 1, 2, 3], 4)[:W].astype(np.int32)
    gw = Walkers(dm, Ts, m_capacity=4096, seed=41)
    if with_comm:
        gw.comm_init(Walkers.comm_unique_id(), 0, 1)
    gw.init()
    gw.sweep(10, thermalized=False)
    gw.sweep(6, thermalized=True, measure=True)
    sums, counts = gw.fetch_accumulators(reset=False)
    gs, gc = gw.reduce_bins(group, 4, reset=True)
    for g in range(4):
        sel = group == g
        expect = np.zeros(gw.n_obs)
        for i in np.nonzero(sel)[0]:  # the library adds in walker order
            expect = expect + sums[i]
        assert np.array_equal(gs[g], expect)
        assert np.array_equal(gc[g], counts[sel].sum(axis=0))
    s2, c2 = gw.fetch_accumulators()
    assert not s2.any() and not c2.any()
    one, onec = gw.reduce_bins()  # a single group, empty bin
    assert one.shape == (1, gw.n_obs) and not one.any() and not onec.any()
    with pytest.raises(SSEError):
        gw.reduce_bins(group, 3)


def test_emu_reduce_bins(emu):
    _body_reduce_bins(False)


@pytest.mark.gpu
@pytest.mark.parametrize("with_comm", [False, True])
def test_gpu_reduce_bins(with_comm):
    _body_reduce_bins(with_comm)


def _body_device_replica_exchange():
    """sse_pt_exchange: the device's swap decisions (log weight ratio of src/sse.jl:395 on the device, uniforms from the
    Philox stream (seed, step)) equal tempering.swap_decisions fed with the same uniforms; temperatures stay a
    permutation of the ladder; bins are attributed to temperatures through the ladder."""
    from helpers import heisenberg_square, isconsistent
    from sse_b200.tempering import DeviceReplicaExchange, pt_uniforms, swap_decisions

    model = heisenberg_square(4, False)
    dm, om = G._pair(model)
    ladder = np.linspace(0.3, 1.2, 11)
    perm = np.random.default_rng(3).permutation(len(ladder))
    gw = Walkers(dm, ladder[perm], m_capacity=4096, seed=8)
    gw.init()
    gw.sweep(40, thermalized=False)
    rx = DeviceReplicaExchange(gw, seed=77)
    assert np.array_equal(gw.pt_get_ladder(), np.argsort(ladder[perm], kind="stable"))
    for it in range(12):
        gw.sweep(3, thermalized=True, measure=True)
        with pytest.raises(RuntimeError):
            rx.step()                      # a bin is open
        rank = rx.rank_of_walker()
        sums, counts = gw.reduce_bins(rank, len(ladder))  # bins by temperature rank
        assert np.all(counts[:, 0] == 3)
        n, T = gw.num_operators(), gw.temperatures()
        order, parity = gw.pt_get_ladder(), rx.parity
        expect = swap_decisions(n, T, order, parity, pt_uniforms(77, it, len(ladder)))
        acc = rx.step()
        T_new = gw.temperatures()
        assert np.array_equal(T_new, expect) and np.array_equal(gw.T, T_new)
        assert acc == np.count_nonzero(T_new != T) // 2
        assert np.array_equal(np.sort(T_new), ladder)
        assert np.array_equal(T_new[gw.pt_get_ladder()], ladder)  # the ladder stays sorted by temperature
    assert 0 < rx.accepted <= rx.proposed
    for i in (0, 5, 10):
        st = gw.get_state(i)
        assert st["T"] == gw.T[i] and isconsistent(st["operators"], st["state"], om.sse_data)


def test_emu_device_replica_exchange(emu):
    _body_device_replica_exchange()


@pytest.mark.gpu
def test_gpu_device_replica_exchange():
    _body_device_replica_exchange()


def _body_grow_capacity():
    """The reference resizes its string in place (sse.jl:138-145); here the capacity is fixed, the overflow is loud, and
    sse_grow_capacity lets the same walkers continue on exactly the oracle's trajectory."""
    from helpers import heisenberg_square

    model = heisenberg_square(4, False)
    dm, om = G._pair(model)
    Ts = np.array([0.08, 0.5, 0.12])
    gw = Walkers(dm, Ts, m_capacity=600, n_capacity=400, seed=61)
    gw.init()
    done = 0
    with pytest.raises(SSEError, match="m_capacity"):
        for _ in range(60):
            gw.sweep(1)
            done += 1
    sd, _ = gw.progress()  # walkers that hit the limit stopped BEFORE their next sweep; the others went on
    with pytest.raises(SSEError):
        gw.grow_capacity(100, 400)
    gw.grow_capacity(4096, 1024)
    # every walker continues; compare each at its own count
    gw.sweep(25)
    sd2, fl = gw.progress()
    assert not fl.any() and (sd2 >= sd + 24).all()
    for i in range(len(Ts)):
        ow = OracleWalker(om, float(Ts[i]), seed=61, walker_id=i)
        ow.init()
        ow.sweep(int(sd2[i]))
        G._same_state(gw.get_state(i), ow.get_state(), f"walker {i} after growing, {sd2[i]} sweeps")
    assert gw.get_state(0)["operators"].shape[0] > 600
    # the same through sweep(auto_grow=True): the call is resumed, the batch stays in step
    gw = Walkers(dm, Ts, m_capacity=600, n_capacity=300, seed=61)
    gw.init()
    for _ in range(12):
        gw.sweep(5, auto_grow=True)
    sd3, fl = gw.progress()
    assert (sd3 == 60).all() and not fl.any() and gw.m_capacity > 600 and gw.n_capacity > 300
    for i in range(len(Ts)):
        ow = OracleWalker(om, float(Ts[i]), seed=61, walker_id=i)
        ow.init()
        ow.sweep(60)
        G._same_state(gw.get_state(i), ow.get_state(), f"auto_grow walker {i}")


def test_emu_grow_capacity(emu):
    _body_grow_capacity()


@pytest.mark.gpu
def test_gpu_grow_capacity():
    _body_grow_capacity()


@pytest.mark.gpu
def test_gpu_temperature_sweep_tool():
    """profiles/tools/temperature_sweep.py (BASELINE configs[4] at small L, one rank): bins reduced per temperature inside
    the library, device-side replica exchange between bins, energies within 4 sigma of the CPU oracle's."""
    import argparse
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "tools"))
    import temperature_sweep as ts

    args = argparse.Namespace(L=4, beta_max=6.0, T_max=2.0, n_T=4, replicas=24, doublings=2, per_level=10, therm=40, bins=12,
                              binsize=25, exchange=True, oracle_check=True, oracle_points=4, oracle_bins=20, oracle_binsize=100,
                              seed=99)
    line = ts.run(args)
    assert line["oracle"]["max_abs_z"] < 4.0, line["oracle"]
    assert 0.0 < line["pt_accept"] <= 1.0
    assert np.all(np.diff(line["operator_count"]) < 0)  # colder temperatures carry longer strings
