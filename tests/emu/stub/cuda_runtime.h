/* Stub for the test-only warp emulator: everything lives in tests/emu/cuda_emu.h (force-included). */
