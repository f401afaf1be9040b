// selftest.cpp — checks the warp emulator itself (tests/emu/cuda_emu.h): collective results, and that the failure
// modes it exists to catch really abort: a collective with a lane missing from the rendezvous, mismatching masks,
// an out-of-bounds store past a device allocation, a misaligned vector access.  Usage: emu_selftest <case>.
#include "cuda_emu.h"

namespace sse {
extern __shared__ __align__(16) uint8_t smem[];
}

__global__ void k_ok(uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t odd = __ballot_sync(0xffffffffu, lane & 1u);
    const uint32_t from5 = __shfl_sync(0xffffffffu, lane * 3u, 5);
    const uint32_t up = __shfl_up_sync(0xffffffffu, lane, 2);
    const uint32_t x = __shfl_xor_sync(0xffffffffu, lane, 16);
    // half-warp masks progress independently of each other
    const uint32_t half = lane < 16 ? 0x0000ffffu : 0xffff0000u;
    uint32_t hb = 0;
    if (lane < 16) hb = __ballot_sync(half, lane == 3);
    else { __syncwarp(half); hb = __ballot_sync(half, lane == 20); }
    // shared memory + __syncthreads across the CTA's warps
    reinterpret_cast<uint32_t *>(sse::smem)[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const uint32_t nb = reinterpret_cast<uint32_t *>(sse::smem)[(threadIdx.x + 32) % blockDim.x];
    uint32_t *o = out + (blockIdx.x * blockDim.x + threadIdx.x) * 6;
    o[0] = odd; o[1] = from5; o[2] = up; o[3] = x; o[4] = hb; o[5] = nb + warp * 0;
}

__global__ void k_missing_lane(uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane != 7) out[0] = __ballot_sync(0xffffffffu, 1);  // lane 7 never arrives
}

__global__ void k_mask_mismatch(uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31u;
    out[lane] = lane < 16 ? __ballot_sync(0xffffffffu, 1) : __ballot_sync(0xffff0000u, 1);
}

__global__ void k_oob(uint32_t *out) { out[64 + (threadIdx.x & 31u)] = 1; }  // allocation holds 64 words

__global__ void k_misaligned(uint32_t *out) { (void)sse::lane_ld128(reinterpret_cast<const uint4 *>(out + 1), 0); }

int main(int argc, char **argv) {
    const std::string c = argc > 1 ? argv[1] : "ok";
    uint32_t *d = nullptr;
    if (c == "ok") {
        const unsigned grid = 2, block = 64;
        cudaMalloc((void **)&d, grid * block * 6 * sizeof(uint32_t));
        emu::launch(k_ok, grid, block, block * 4, d);
        for (unsigned t = 0; t < grid * block; ++t) {
            const uint32_t lane = t & 31u, tid = t % block, *o = d + t * 6;
            const uint32_t exp_hb = lane < 16 ? (1u << 3) : (1u << 20);
            if (o[0] != 0xaaaaaaaau || o[1] != 15u || o[2] != (lane >= 2 ? lane - 2 : lane) || o[3] != (lane ^ 16u) ||
                o[4] != exp_hb || o[5] != (tid + 32) % block) {
                printf("MISMATCH at thread %u: %x %u %u %u %x %u\n", t, o[0], o[1], o[2], o[3], o[4], o[5]);
                return 1;
            }
        }
        cudaFree(d);
        printf("ok\n");
        return 0;
    }
    cudaMalloc((void **)&d, 64 * sizeof(uint32_t));
    if (c == "missing_lane") emu::launch(k_missing_lane, 1u, 32u, 0, d);
    else if (c == "mask_mismatch") emu::launch(k_mask_mismatch, 1u, 32u, 0, d);
    else if (c == "oob") emu::launch(k_oob, 1u, 32u, 0, d);
    else if (c == "misaligned") emu::launch(k_misaligned, 1u, 32u, 0, d);
    printf("NOT DETECTED\n");  // every case above must abort inside the emulator
    return 0;
}
