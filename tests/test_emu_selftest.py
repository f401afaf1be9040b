"""The warp emulator checks itself: collective semantics, and every failure mode it exists to catch aborts loudly."""
import os
import platform
import subprocess

import pytest

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def selftest():
    if platform.machine() != "x86_64":
        pytest.skip("tests/emu needs x86-64")
    subprocess.check_call(["make", "-s", "-C", EMU_DIR, "emu_selftest"])
    return os.path.join(EMU_DIR, "emu_selftest")


def test_collectives_and_barriers(selftest):
    r = subprocess.run([selftest, "ok"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


@pytest.mark.parametrize("case, message", [
    ("missing_lane", "which has exited"),
    ("mask_mismatch", "deadlock: no thread of block 0 can make progress"),
    ("oob", "out-of-bounds write ABOVE"),
    ("misaligned", "misaligned 16-byte global access"),
])
def test_failure_modes_abort(selftest, case, message):
    r = subprocess.run([selftest, case], capture_output=True, text=True)
    assert r.returncode != 0, "the emulator did not detect: " + case
    assert "NOT DETECTED" not in r.stdout
    assert message in r.stderr, r.stderr
