"""The product's kernel SOURCE (csrc/sse_kernels.cuh + sse_capi.cu), compiled with g++ on top of the test-only warp
emulator in tests/emu/, against the CPU oracle — the same bit-exact parity checks as tests/test_gpu_parity.py, run
without a GPU.  The emulator is test infrastructure: the package never loads it (capi.LIB_PATH is patched here, for the
duration of one test).  What this catches before a GPU slot is spent: wrong results, collectives executed under
divergence or with lanes missing from their mask (deadlock report), missing __syncwarp, reads of uninitialised shared /
device memory, misaligned vector accesses, out-of-bounds writes past a device allocation (red zones).  What it cannot
see: performance, and hazards that only a real memory system exposes — the `-m gpu` tests remain the parity gate."""
import os
import subprocess

import pytest

import test_gpu_parity as G
from sse_b200 import capi

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
EMU_LIB = os.path.join(EMU_DIR, "libsse_b200_emu.so")


@pytest.fixture(scope="module")
def emu_built():
    import platform

    if platform.machine() != "x86_64":
        pytest.skip("tests/emu needs x86-64")
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return EMU_LIB


@pytest.fixture()
def emu(emu_built, monkeypatch):
    """Point the ctypes layer at the emulator build for one test, then restore the real library path."""
    saved = capi._lib
    monkeypatch.setattr(capi, "LIB_PATH", emu_built)
    capi._lib = None
    capi.lib()
    yield
    capi._lib = saved


def test_emu_exports_the_whole_abi(emu):
    L = capi.lib()
    for name in capi.EXPORTED_SYMBOLS:
        assert hasattr(L, name)
    assert L.sse_abi_version() == 2


def test_emu_vertex_list_golden_vector(emu):
    G.test_vertex_list_golden_vector_gpu()


@pytest.mark.parametrize("name", list(G.MODEL_CLASSES))
def test_emu_phase_parity_injected_stream(emu, name):
    G.test_phase_parity_injected_stream(name)


@pytest.mark.parametrize("name", ["heisenberg_eof", "spin1_dz", "dimer_bilayer"])
def test_emu_sweep_parity_philox(emu, name):
    G.test_sweep_parity_philox(name, W=10, therm=25, meas=10)  # the emulated "GPU" has 3 SMs: 3-4 walkers per CTA


def test_emu_worm_traverse_reference_cases(emu):
    G.test_worm_traverse_reference_cases()


def test_emu_measure_matches_oracle(emu):
    G.test_measure_matches_oracle()


def test_emu_checkpoint_roundtrip_and_pt_hooks(emu):
    G.test_checkpoint_roundtrip_and_pt_hooks()


def test_emu_overflow_is_loud(emu):
    G.test_overflow_is_loud()


def test_emu_edge_cases_empty_and_ragged_strings(emu):
    G.test_edge_cases_empty_and_ragged_strings()


def test_emu_api_errors_are_loud(emu):
    G.test_api_errors_are_loud()


@pytest.mark.parametrize("level", [0, 1])
def test_emu_large_lattice_memory_paths(emu, level, monkeypatch):
    G.test_large_lattice_memory_paths(level, monkeypatch)


# ---- host logic above the C ABI (mc.MC / carlo stand-in / tempering) driven by the emulated kernels ---------------
def test_emu_mc_carlo_interface_and_checkpoint(emu):
    G.test_mc_carlo_interface_and_checkpoint()


def test_emu_replica_exchange_on_device_walkers(emu):
    G.test_replica_exchange_on_device_walkers()


def test_emu_smoke_entry(emu):
    """__graft_entry__.smoke()'s own logic (64 walkers x 40 sweeps + estimators vs the oracle), on the emulated kernels."""
    import __graft_entry__ as g

    g.smoke()


def test_emu_reference_dump_replay_machinery(emu):
    """The replay of julia/dump_golden.jl dumps through the device path (tests/test_reference_dumps.py), self-checked on
    an oracle-written dump of the same schema, here with the emulated kernels as the device."""
    import test_reference_dumps as R

    R.test_replay_machinery_selfcheck_gpu()
