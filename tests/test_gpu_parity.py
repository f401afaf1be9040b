"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the
C ABI (libsse_b200.so) and compares it bit-for-bit with the CPU oracle on the same operator string
and the same injected random stream (north star, level 1)."""
import numpy as np
import pytest

from helpers import MODEL_CLASSES, heisenberg_square, isconsistent, random_stream
from oracle import OracleModel, OracleWalker
from sse_b200.capi import model_desc_from_model
from sse_b200.util import opercode, vertex_code
from sse_b200.walkers import DeviceModel, Walkers

pytestmark = pytest.mark.gpu


def _pair(model):
    desc, keep, sse_data = model_desc_from_model(model)
    dm = DeviceModel(model=model, desc=desc, keep=keep, sse_data=sse_data)
    om = OracleModel(desc=desc, keep=keep, sse_data=sse_data)
    return dm, om


def _same_state(a, b, what=""):
    assert a["num_operators"] == b["num_operators"], what
    assert len(a["operators"]) == len(b["operators"]), what
    bad = np.nonzero(a["operators"] != b["operators"])[0]
    assert len(bad) == 0, f"{what}: operators differ at slots {bad[:8]}"
    assert np.array_equal(a["state"], b["state"]), what
    assert a["rng_draws"] == b["rng_draws"], what
    assert a["num_worms"] == b["num_worms"], what
    assert a["avg_worm_length"] == b["avg_worm_length"], what


def test_vertex_list_golden_vector_gpu():
    """test/test_vertex_list.jl:1-30 — the reference's only bit-exact golden vector on the hot path."""
    import sse_b200 as S
    from sse_b200.capi import build_model_desc
    from sse_b200.sse_data import SSEBond, SSEData
    from sse_b200.vertex_data import make_vertex_data

    splus, sz = S.spin_operators(2)
    H = np.kron(sz, sz) + 0.5 * (np.kron(splus, splus.T) + np.kron(splus.T, splus))
    vd = make_vertex_data((2, 2), H, energy_offset_factor=0.0)
    sd = SSEData([vd], [SSEBond(1, (1, 2)), SSEBond(1, (2, 3)), SSEBond(1, (1, 3))])
    flat = sd.flatten()
    flat["n_sites"] = 4  # site_count = 4 in the reference test; site 4 carries no bond
    flat["site_dim"] = np.array([2, 2, 2, 2], dtype=np.uint8)
    desc, keep = build_model_desc(flat, 4, None)
    dm = DeviceModel(desc=desc, keep=keep, sse_data=sd)
    v = vertex_code(False, 2)  # an off-diagonal vertex of the S=1/2 table
    ops = np.array([0, 0, opercode(1, v), 0, opercode(1, v), opercode(2, v), opercode(3, v)], dtype=np.uint64)
    w = Walkers(dm, [1.0], m_capacity=128)
    w.set_state(0, dict(num_operators=4, operators=ops, state=np.ones(4, dtype=np.uint8), T=1.0))
    w.dbg_make_vertex_list()
    vert, vf, vl = w.dbg_get_vertex_list(0, 7)
    expected = -np.ones((7, 4, 2), dtype=np.int64)
    expected[2] = [(3, 7), (3, 6), (1, 5), (2, 5)]
    expected[4] = [(3, 3), (4, 3), (1, 7), (1, 6)]
    expected[5] = [(4, 5), (4, 7), (2, 3), (2, 7)]
    expected[6] = [(3, 5), (4, 6), (1, 3), (2, 6)]
    assert np.array_equal(vert, expected)
    assert vl.tolist() == [[3, 7], [3, 6], [4, 7], [-1, -1]]
    assert vf.tolist() == [[1, 3], [2, 3], [2, 6], [-1, -1]]
    # building the records leaves the string untouched
    assert np.array_equal(w.get_state(0)["operators"], ops)


@pytest.mark.parametrize("name", list(MODEL_CLASSES))
def test_phase_parity_injected_stream(name):
    """diagonal_update, make_vertex_list!, worm_update each bit-exact vs the oracle under an injected stream."""
    model = MODEL_CLASSES[name]()
    dm, om = _pair(model)
    T = 0.3
    W = 8
    rng = np.random.default_rng(1234)
    # starting configurations: oracle walkers advanced by a few sweeps each (different lengths)
    starts = []
    for i in range(W):
        ow = OracleWalker(om, T, seed=99, walker_id=i)
        ow.init()
        ow.sweep(3 + 2 * i)
        starts.append(ow)
    gw = Walkers(dm, [T] * W, m_capacity=4096)
    for i, ow in enumerate(starts):
        gw.set_state(i, ow.get_state())

    # --- diagonal update ---
    L = 60000
    streams = np.stack([random_stream(rng, L) for _ in range(W)])
    gw.set_injected_stream(streams)
    for i, ow in enumerate(starts):
        ow.set_injected_stream(streams[i])
    gw.dbg_diagonal_update()
    for i, ow in enumerate(starts):
        ow.diagonal_update()
        assert not ow.stream_exhausted
        _same_state(gw.get_state(i), ow.get_state(), f"{name} diagonal_update walker {i}")

    # --- vertex list ---
    gw.dbg_make_vertex_list()
    for i, ow in enumerate(starts):
        ow.make_vertex_list()
        M = len(ow.get_state()["operators"])
        gv, gf, gl = gw.dbg_get_vertex_list(i, M)
        ov, of, ol = ow.get_vertex_list()
        assert np.array_equal(gv, ov), f"{name} vertex list walker {i}"
        assert np.array_equal(gf, of) and np.array_equal(gl, ol)

    # --- worm update (not thermalised: controller runs) ---
    gw.dbg_worm_update(False)
    for i, ow in enumerate(starts):
        ow.worm_update(False)
        assert not ow.stream_exhausted
        _same_state(gw.get_state(i), ow.get_state(), f"{name} worm_update walker {i}")
        st = ow.get_state()
        assert isconsistent(st["operators"], st["state"], om.sse_data)


@pytest.mark.parametrize("name", list(MODEL_CLASSES))
def test_sweep_parity_philox(name, W=33, therm=40, meas=25):
    """Whole runs (init! + thermalisation with string growth + measured sweeps) agree exactly with the oracle
    when both draw from the Philox stream of the same (seed, walker id)."""
    model = MODEL_CLASSES[name]()
    dm, om = _pair(model)
    Ts = np.linspace(0.15, 1.5, W)
    gw = Walkers(dm, Ts, m_capacity=8192, seed=4242, walker_id_offset=7)
    gw.init()
    gw.sweep(therm, thermalized=False)
    gw.sweep(meas, thermalized=True, measure=True)
    sums, counts = gw.fetch_accumulators()
    cnt = gw.fetch_counters()
    visits = 0
    for i in (0, 1, W // 2, W - 1):
        ow = OracleWalker(om, float(Ts[i]), seed=4242, walker_id=7 + i)
        ow.init()
        ow.sweep(therm, thermalized=False)
        ow.sweep(meas, thermalized=True, measure=True)
        st = ow.get_state()
        _same_state(gw.get_state(i), st, f"{name} walker {i}")
        assert isconsistent(st["operators"], st["state"], om.sse_data)
        osums, ocounts = ow.fetch_accumulators()
        assert np.array_equal(counts[i], ocounts)
        np.testing.assert_allclose(sums[i], osums, rtol=1e-12, atol=1e-300)
        # instantaneous measure! agrees too
    assert cnt["sweeps"] == W * (therm + meas)
    assert (gw.get_flags() & 7).sum() == 0


def test_worm_traverse_reference_cases():
    """test/test_sse.jl:30-60: worm_traverse!((1,1,1), ...) on 1- and 2-operator strings stays consistent,
    and matches the oracle under the same stream."""
    model = heisenberg_square(4, True)
    dm, om = _pair(model)
    sd = om.sse_data
    vd = sd.get_vertex_data(1)
    strings = [
        [opercode(1, int(vd.diagonal_vertices[2]))],
        [opercode(1, vertex_code(True, 1)), opercode(1, vertex_code(True, 1))],
    ]
    rng = np.random.default_rng(7)
    for ops in strings:
        ops = np.array(ops, dtype=np.uint64)
        # a state consistent with the diagonal operators on bond 1 = sites (1, 2)
        ls = vd.get_leg_state((int(ops[0]) & ((1 << 25) - 1)) >> 1)
        state = np.ones(16, dtype=np.uint8)
        b = sd.bonds[0]
        state[b.sites[0] - 1], state[b.sites[1] - 1] = ls[0], ls[1]
        start = dict(num_operators=len(ops), operators=ops, state=state, T=0.1)
        stream = random_stream(rng, 4096)
        gw = Walkers(dm, [0.1], m_capacity=128)
        gw.set_state(0, start)
        gw.set_injected_stream(stream[None, :])
        gw.dbg_make_vertex_list()
        glen = gw.dbg_worm_traverse(1, 1, 1)[0]
        ow = OracleWalker(om, 0.1)
        ow.set_state(start)
        ow.set_injected_stream(stream)
        ow.make_vertex_list()
        olen = ow.worm_traverse(1, 1, 1)
        assert glen == olen
        gs, os_ = gw.get_state(0), ow.get_state()
        assert np.array_equal(gs["operators"], os_["operators"])
        assert gs["rng_draws"] == os_["rng_draws"]
        # state0 from v_first as in the reference test, then isconsistent
        v, vf, vl = ow.get_vertex_list()
        state0 = np.ones(16, dtype=np.int64)
        for s in range(16):
            if vf[s, 0] > 0:
                op = int(os_["operators"][vf[s, 1] - 1])
                state0[s] = sd.get_vertex_data(op >> 26).get_leg_state((op & ((1 << 25) - 1)) >> 1)[vf[s, 0] - 1]
        assert isconsistent(gs["operators"], state0, sd)


def test_measure_matches_oracle():
    """Carlo.measure! on the device == oracle for all 8 magnetization estimators (testjob_magnet_square)."""
    import sse_b200 as S
    from sse_b200.estimators import all_magnetization_estimators

    model = S.MagnetModel(dict(lattice=dict(unitcell=S.UnitCells.square, size=(2, 4)), J=1.23, hz=-0.2,
                               measure=all_magnetization_estimators(2)))
    dm, om = _pair(model)
    W = 5
    Ts = np.linspace(0.2, 2.0, W)
    gw = Walkers(dm, Ts, m_capacity=4096, seed=3)
    gw.init()
    gw.sweep(30, thermalized=False)
    gw.sweep(3, thermalized=True)
    g = gw.measure()
    assert g.shape == (W, 6 + 5 * 8)
    for i in range(W):
        ow = OracleWalker(om, float(Ts[i]), seed=3, walker_id=i)
        ow.init()
        ow.sweep(30, thermalized=False)
        ow.sweep(3, thermalized=True)
        np.testing.assert_allclose(g[i], ow.measure(), rtol=1e-12, atol=1e-300)


def test_checkpoint_roundtrip_and_pt_hooks():
    model = MODEL_CLASSES["spin1_dz"]()
    dm, om = _pair(model)
    gw = Walkers(dm, [0.5, 0.7], m_capacity=4096, seed=11)
    gw.init()
    gw.sweep(20)
    s0, s1 = gw.get_state(0), gw.get_state(1)
    n = gw.num_operators()
    assert n[0] == s0["num_operators"] and n[1] == s1["num_operators"]
    lw = gw.pt_log_weight_ratio([0.7, 0.5])
    assert lw[0] == -s0["num_operators"] * np.log(0.7 / 0.5)
    # restore into a fresh batch (swapped walkers) and continue: identical trajectories
    gw2 = Walkers(dm, [0.5, 0.7], m_capacity=4096, seed=11)
    gw2.set_state(0, s0)
    gw2.set_state(1, s1)
    gw.sweep(5)
    gw2.sweep(5)
    _same_state(gw.get_state(0), gw2.get_state(0))
    _same_state(gw.get_state(1), gw2.get_state(1))
    # the batched calls (sse_get_states / sse_set_states): one round of copies, same result
    gw3 = Walkers(dm, [0.5, 0.7], m_capacity=4096, seed=11)
    gw3.set_states([s0, s1])
    gw3.sweep(5)
    for a, b in zip(gw.get_states(), gw3.get_states()):
        _same_state(a, b)
    # a batch with one invalid state is rejected as a whole: walker 0 keeps what it had
    from sse_b200.capi import SSEError

    before = gw3.get_states()
    bad = dict(s1, operators=s1["operators"].copy())
    bad["operators"][np.flatnonzero(bad["operators"])[0]] |= np.uint64(1) << np.uint64(60)  # bond out of range
    with pytest.raises(SSEError):
        gw3.set_states([s0, bad])
    for a, b in zip(before, gw3.get_states()):
        _same_state(a, b)
    gw.set_temperature([0.9, 0.9])
    assert gw.get_state(0)["T"] == 0.9


def test_overflow_is_loud():
    from sse_b200.capi import SSEError

    model = heisenberg_square(4, True)
    dm, _ = _pair(model)
    gw = Walkers(dm, [0.05], m_capacity=128)
    with pytest.raises(SSEError):
        gw.init()
        gw.sweep(50)


def test_edge_cases_empty_and_ragged_strings():
    """Empty strings (n = 0: worm_traverse! returns 0 without drawing, every site is redrawn), strings whose
    length is not a multiple of the 32-slot chunk, and walkers of very different lengths in one batch."""
    model = MODEL_CLASSES["mixed_honeycomb"]()
    dm, om = _pair(model)
    Ts = np.array([50.0, 50.0, 5.0, 0.5, 0.07])  # n ~ 0 at T = 50
    W = len(Ts)
    gw = Walkers(dm, Ts, m_capacity=8192, seed=99)
    gw.init(init_opstring_cutoff=37, diagonal_warmup_sweeps=0)  # odd length, no warm-up: starts with n = 0
    ows = []
    for i in range(W):
        ow = OracleWalker(om, float(Ts[i]), seed=99, walker_id=i)
        ow.init(37, 0)
        ows.append(ow)
    for step in range(6):
        gw.sweep(1, thermalized=step >= 3, measure=step >= 3)
        for i, ow in enumerate(ows):
            ow.sweep(1, thermalized=step >= 3, measure=step >= 3)
            _same_state(gw.get_state(i), ow.get_state(), f"step {step} walker {i}")
    n = gw.num_operators()
    assert n[0] < 8 and n[4] > 100
    sums, counts = gw.fetch_accumulators()
    for i, ow in enumerate(ows):
        osums, ocounts = ow.fetch_accumulators()
        assert np.array_equal(counts[i], ocounts)
        np.testing.assert_allclose(sums[i], osums, rtol=1e-12, atol=1e-300)


def test_api_errors_are_loud():
    from sse_b200.capi import SSEError

    model = heisenberg_square(4, True)
    dm, om = _pair(model)
    gw = Walkers(dm, [0.5, 0.5], m_capacity=2048, seed=1)
    gw.init()
    gw.sweep(5)
    st = gw.get_state(0)
    bad = dict(st)
    bad["num_operators"] = st["num_operators"] + 1
    with pytest.raises(SSEError):
        gw.set_state(0, bad)
    bad = dict(st)
    ops = st["operators"].copy()
    ops[np.nonzero(ops)[0][0]] |= np.uint64(1) << np.uint64(60)  # bond index out of range
    bad["operators"] = ops
    with pytest.raises(SSEError):
        gw.set_state(0, bad)
    bad = dict(st)
    s2 = st["state"].copy()
    s2[0] = 3
    bad["state"] = s2
    with pytest.raises(SSEError):
        gw.set_state(0, bad)
    with pytest.raises(SSEError):
        gw.dbg_worm_update(False)  # no vertex list built
    with pytest.raises(SSEError):
        gw.set_temperature([0.5, -1.0])
    # an injected stream that runs out is reported, not silently recycled
    gw.set_injected_stream(np.zeros((2, 16), dtype=np.uint64))
    with pytest.raises(SSEError):
        gw.dbg_diagonal_update()
    with pytest.raises(SSEError):
        Walkers(dm, [0.5], m_capacity=2048, n_capacity=1 << 23)


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


def test_mc_carlo_interface_and_checkpoint():
    """The Carlo-facing mirror (mc.MC + carlo.run): Carlo's sweep!/measure! loop and the device-resident loop are
    the same Markov chain with the same per-bin sums; write_checkpoint/read_checkpoint resume it exactly."""
    import sse_b200 as S
    from sse_b200.carlo import MCContext, run
    from sse_b200.mc import MC, evaluate_group

    params = dict(model=S.MagnetModel, lattice=dict(unitcell=S.UnitCells.chain, size=(16,)), J=1.0, T=0.1,
                  n_walkers=8, measure=["magnetization", "staggered_magnetization"], seed=5, sweeps=40,
                  thermalization=30, binsize=10)  # BASELINE config 0: spin-1/2 chain L=16, T=0.1
    mc_a, mc_b = MC(params), MC(params)
    ctx_b = run(mc_b, params, fused=True)
    # Carlo's own loop: sweep!; sweeps += 1; measure! once thermalised (one launch per call)
    ctx_a = MCContext(params)
    mc_a.init(ctx_a, params)
    for s in range(params["thermalization"]):
        mc_a.sweep_many(ctx_a, 1, thermalized=False, measure=False)
    for s in range(params["sweeps"]):
        mc_a.sweep_many(ctx_a, 1, thermalized=True, measure=False)
        mc_a.measure(ctx_a)
    for name in ("Sign", "OperatorCount", "SignEnergy", "SignMag2", "SignStagMag2", "SignStagMagChi"):
        a, b = ctx_a.bin_array(name), ctx_b.bin_array(name)
        assert a.shape == b.shape == (4, 8)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-300)
    res = evaluate_group(ctx_b, mc_b, range(8))
    assert -0.47 < res["Energy"][0] < -0.42  # e0(L=16 chain) ~ -0.446 per site
    assert set(res) >= {"Energy", "SpecificHeat", "Mag", "StagMag2", "StagBinderRatio", "StagMagChi"}
    # checkpoint round trip through the reference's five fields
    ck = mc_b.write_checkpoint()
    mc_c = MC(params)
    mc_c.read_checkpoint(ck)
    ctx_c = MCContext(params)
    mc_b.sweep_many(ctx_b, 7, thermalized=True, measure=False)
    mc_c.sweep_many(ctx_c, 7, thermalized=True, measure=False)
    for i in (0, 7):
        _same_state(mc_b.walkers.get_state(i), mc_c.walkers.get_state(i), f"checkpoint walker {i}")
    lw = mc_b.parallel_tempering_log_weight_ratio("T", 0.2)
    assert lw.shape == (8,) and np.all(lw < 0)
    with pytest.raises(ValueError):
        mc_b.parallel_tempering_change_parameter("J", 1.0)


@pytest.mark.parametrize("level", [0, 1])
def test_large_lattice_memory_paths(level, monkeypatch):
    """Lattices too large for the per-warp shared-memory arrays keep vlast (level 1) or state/mark/vlast (level 0)
    in global memory.  Force those paths on a small model and require the same bit-exact trajectory."""
    monkeypatch.setenv("SSE_B200_SMEM_LEVEL", str(level))
    model = MODEL_CLASSES["spin1_dz"]()
    dm, om = _pair(model)
    Ts = np.array([0.2, 0.6, 1.5])
    gw = Walkers(dm, Ts, m_capacity=8192, seed=77)
    gw.init()
    gw.sweep(30, thermalized=False)
    gw.sweep(10, thermalized=True, measure=True)
    sums, counts = gw.fetch_accumulators()
    for i in range(len(Ts)):
        ow = OracleWalker(om, float(Ts[i]), seed=77, walker_id=i)
        ow.init()
        ow.sweep(30, thermalized=False)
        ow.sweep(10, thermalized=True, measure=True)
        _same_state(gw.get_state(i), ow.get_state(), f"level {level} walker {i}")
        osums, ocounts = ow.fetch_accumulators()
        np.testing.assert_allclose(sums[i], osums, rtol=1e-12, atol=1e-300)
    # the vertex list read-back works without the shared-memory vlast copy as well
    gw.dbg_make_vertex_list()
    ow.make_vertex_list()
    M = len(ow.get_state()["operators"])
    gv, gf, gl = gw.dbg_get_vertex_list(len(Ts) - 1, M)
    ov, of, ol = ow.get_vertex_list()
    assert np.array_equal(gv, ov) and np.array_equal(gf, of) and np.array_equal(gl, ol)


def test_replica_exchange_on_device_walkers():
    """Parallel-tempering hooks (sse.jl:390-405) driven by tempering.ReplicaExchange: temperatures stay a permutation
    of the ladder, the device sees the new temperatures, swaps do happen, and configurations stay consistent."""
    from sse_b200.tempering import ReplicaExchange

    model = heisenberg_square(4, False)
    dm, om = _pair(model)
    ladder = np.linspace(0.3, 1.2, 12)
    gw = Walkers(dm, ladder, m_capacity=4096, seed=8)
    gw.init()
    gw.sweep(50, thermalized=False)
    rx = ReplicaExchange(gw, seed=1)
    for _ in range(20):
        gw.sweep(3, thermalized=True)
        with pytest.raises(RuntimeError):
            rx.step()            # WormLengthFraction samples of these sweeps are still in the bin
        gw.fetch_accumulators()  # flush: a bin must not span a change of temperature
        T_new = rx.step()
        assert np.allclose(np.sort(T_new), ladder)
    assert rx.proposed > 0 and 0 < rx.accepted <= rx.proposed
    for i in (0, 5, 11):
        st = gw.get_state(i)
        assert st["T"] == gw.T[i]
        assert isconsistent(st["operators"], st["state"], om.sse_data)
    # colder temperature labels end up on longer operator strings on average
    n = gw.num_operators()
    order = np.argsort(gw.T)
    assert n[order[:3]].mean() > n[order[-3:]].mean()
