"""2 and 4 walkers per warp (sse::k_walkers_multi, sse_set_walkers_per_warp): the interleaved worm updates must leave
every walker on exactly the trajectory of the one-walker-per-warp kernel and of the oracle.  CPU: through the warp
emulator (tests/emu); GPU: `-m gpu`.  The file sorts last on purpose: it was written when
round 1 had two GPU-minutes left (first B200 run: profiles/r1_f_gpu_new_tests.txt, all passed)."""
import numpy as np
import pytest

import test_gpu_parity as G
from helpers import MODEL_CLASSES, random_stream
from oracle import OracleWalker
from sse_b200.capi import SSEError
from sse_b200.walkers import Walkers
from test_emu_parity import emu, emu_built  # noqa: F401  (fixtures)


def _body_injected_sweeps(k):
    """sse_sweep under an injected stream (the dbg_* parity hooks always run one walker per warp, so this is the
    injected-stream coverage of the interleaved kernel), walkers of different lengths, a ragged last warp."""
    model = MODEL_CLASSES["spin1_dz"]()
    dm, om = G._pair(model)
    W = 7
    Ts = np.linspace(0.25, 1.0, W)
    rng = np.random.default_rng(5)
    gw = Walkers(dm, Ts, m_capacity=4096, seed=3)
    gw.set_walkers_per_warp(k)
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
    # a stream that runs out inside the interleaved worm phase is reported, not recycled
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
    for k in (2, 4, 1, 4, 2):
        b.set_walkers_per_warp(k)
        a.sweep(4)
        b.sweep(4)
    for i in range(len(Ts)):
        G._same_state(a.get_state(i), b.get_state(i), f"walker {i}")
    with pytest.raises(SSEError):
        b.set_walkers_per_warp(3)


@pytest.mark.parametrize("k", [2, 4])
def test_emu_multi_injected_sweeps(emu, k):
    _body_injected_sweeps(k)


def test_emu_multi_switching_keeps_the_trajectory(emu):
    _body_switching_keeps_the_trajectory()


@pytest.fixture(params=[2, 4])
def chains_env(request, monkeypatch):
    monkeypatch.setenv("SSE_B200_CHAINS", str(request.param))
    return request.param


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["heisenberg_eof", "spin1_dz", "dimer_bilayer"])
def test_gpu_multi_sweep_parity_philox(chains_env, name):
    G.test_sweep_parity_philox(name)


@pytest.mark.gpu
def test_gpu_multi_edge_cases(chains_env):
    G.test_edge_cases_empty_and_ragged_strings()


@pytest.mark.gpu
@pytest.mark.parametrize("level", [0, 1])
def test_gpu_multi_large_lattice_memory_paths(chains_env, level, monkeypatch):
    G.test_large_lattice_memory_paths(level, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 4])
def test_gpu_multi_injected_sweeps(k):
    _body_injected_sweeps(k)


@pytest.mark.gpu
def test_gpu_multi_switching_keeps_the_trajectory():
    _body_switching_keeps_the_trajectory()


def _body_occupancy_variants(monkeypatch):
    """SSE_B200_MULTI_MINB selects another compiled occupancy of the interleaved kernel (tuning only): same trajectories;
    a combination that was not compiled is an error, not a silent fallback."""
    model = MODEL_CLASSES["heisenberg_eof"]()
    dm, om = G._pair(model)
    Ts = np.linspace(0.2, 0.9, 6)
    ref = Walkers(dm, Ts, m_capacity=4096, seed=21)
    ref.init()
    ref.sweep(6)
    for k, minb in ((2, 5), (4, 4), (4, 3)):
        monkeypatch.setenv("SSE_B200_MULTI_MINB", str(minb))
        w = Walkers(dm, Ts, m_capacity=4096, seed=21)
        w.set_walkers_per_warp(k)
        w.init()
        w.sweep(6)
        for i in range(len(Ts)):
            G._same_state(ref.get_state(i), w.get_state(i), f"{k} per warp at {minb} CTAs/SM, walker {i}")
    monkeypatch.setenv("SSE_B200_MULTI_MINB", "6")
    w = Walkers(dm, Ts, m_capacity=4096, seed=21)
    w.set_walkers_per_warp(2)
    w.init()
    with pytest.raises(SSEError):
        w.sweep(1)


def test_emu_multi_occupancy_variants(emu, monkeypatch):
    _body_occupancy_variants(monkeypatch)


@pytest.mark.gpu
def test_gpu_multi_occupancy_variants(monkeypatch):
    _body_occupancy_variants(monkeypatch)
