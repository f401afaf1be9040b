"""CPU checks: the C-ABI library loads and exports every symbol include/sse_b200.h declares (no compute
calls without a GPU); walker sharding + bin reduction over a 2-rank gloo group."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    from sse_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "sse_b200.h")).read()
    declared = set(re.findall(r"\b(sse_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"sse_model_desc", "sse_walkers_opts", "sse_walker_state"}
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = capi.lib()  # raises if the .so or any symbol is missing
    for name in declared:
        assert hasattr(lib, name)
    assert lib.sse_abi_version() == 2


def test_product_path_has_no_oracle_dependency():
    """The product package must never import, include or link the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "stochasticseriesexpansion.jl_b200")
    bad = re.compile(r"^\s*(import|from)\s+oracle\b|libsse_oracle|#include\s+\".*oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f"{f} depends on the oracle"


def test_product_path_never_loads_the_emulator():
    """tests/emu is test infrastructure: no Python file of the package may name the emulator library, and the only
    hooks in csrc are the two preprocessor guards the emulator build overrides."""
    pkg = os.path.join(ROOT, "stochasticseriesexpansion.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                assert "emu" not in open(path).read().lower().replace("enumerate", ""), f
            if f.endswith((".cu", ".cuh")):
                txt = open(path).read()
                assert "cuda_emu" not in txt.replace("tests/emu/cuda_emu.h", "") and "#include \"../../tests" not in txt, f
    from sse_b200 import capi

    assert capi.LIB_PATH.endswith(os.path.join("csrc", "libsse_b200.so"))


def test_shard_walkers_partition():
    from sse_b200.sharding import shard_walkers

    for n, world in [(4096, 8), (10, 4), (7, 8), (1, 1)]:
        seen = []
        for r in range(world):
            off, cnt = shard_walkers(n, r, world)
            seen += list(range(off, off + cnt))
        assert seen == list(range(n))


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sse_b200  # noqa: F401
    from helpers import heisenberg_square
    from oracle import OracleModel, OracleWalker
    from sse_b200.sharding import reduce_bins, shard_walkers

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_total = 6
    Ts = np.array([0.5, 0.5, 0.5, 1.0, 1.0, 1.0])
    groups = np.array([0, 0, 0, 1, 1, 1])
    off, cnt = shard_walkers(n_total, rank, world)
    om = OracleModel(heisenberg_square(2, True))
    sums, counts = [], []
    for g in range(off, off + cnt):  # stand-in for the device walkers: same stream ids as a 1-rank run
        w = OracleWalker(om, float(Ts[g]), seed=5, walker_id=g)
        w.init()
        w.sweep(20, thermalized=True, measure=True)
        s, c = w.fetch_accumulators()
        sums.append(s)
        counts.append(c)
    gs, gc = reduce_bins(np.array(sums), np.array(counts), groups[off:off + cnt], n_groups=2)
    q.put((rank, gs.numpy(), gc.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_bin_reduction_gloo_world2():
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from helpers import heisenberg_square
    from oracle import OracleModel, OracleWalker

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    om = OracleModel(heisenberg_square(2, True))
    Ts = [0.5, 0.5, 0.5, 1.0, 1.0, 1.0]
    ref_s = np.zeros((2, results[0][1].shape[1]))
    ref_c = np.zeros((2, 2))
    for g, T in enumerate(Ts):
        w = OracleWalker(om, T, seed=5, walker_id=g)
        w.init()
        w.sweep(20, thermalized=True, measure=True)
        sm, c = w.fetch_accumulators()
        ref_s[g // 3] += sm
        ref_c[g // 3] += c
    for rank, gs, gc in results:
        np.testing.assert_allclose(gs, ref_s, rtol=1e-13)
        np.testing.assert_array_equal(gc, ref_c)
