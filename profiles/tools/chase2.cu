// chase2.cu — pointer-chase throughput on B200 with ONE LANE = ONE dependent chain (divergent addresses), the access
// pattern of the lane-per-walker worm phase: every hop is a 16-byte ld.global.cg of a record in the chain's own region,
// optionally followed by a 4-byte st.global back into the record just loaded (the visit's op-code store).
// Sweeps the number of chains in flight, the footprint per chain and the warps per SM that carry them.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chase2 chase2.cu ;  run: ./chase2 [max_total_GB]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void init(uint4 *buf, size_t n, size_t total) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        uint32_t k = (uint32_t)(i & (n - 1));
        uint32_t nxt = (uint32_t)((1664525ull * k + 1013904223ull) & (n - 1));  // full-period LCG, n a power of two
        buf[i] = make_uint4(k, nxt, 0, 0);
    }
}
// rec32 = 1: records are 32 bytes (two 16-byte loads of one sector), as in the round-1 layout
__global__ void chase(uint4 *buf, size_t n, int C, int hops, int store, int rec32, unsigned long long *out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    uint4 *base = buf + (size_t)c * n;
    uint32_t k = (c * 7919u) & (uint32_t)(n - 1);
    if (rec32) k &= ~1u;
    uint32_t acc = 0;
    for (int h = 0; h < hops; ++h) {
        uint4 R, H;
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(R.x), "=r"(R.y), "=r"(R.z), "=r"(R.w) : "l"(base + k));
        if (rec32) {
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(H.x), "=r"(H.y), "=r"(H.z), "=r"(H.w) : "l"(base + k + 1));
            acc += H.x;
        }
        if (store) asm volatile("st.global.u32 [%0], %1;" ::"l"(base + k), "r"(R.x) : "memory");
        k = R.y;
        if (rec32) k &= ~1u;
    }
    if (k == 0xffffffffu || acc == 0x12345u) atomicAdd(out, 1ull);
}
int main(int argc, char **argv) {
    const double max_gb = argc > 1 ? atof(argv[1]) : 165.0;
    unsigned long long *out;
    cudaMalloc(&out, 16);
    cudaMemset(out, 0, 16);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs\n", prop.name, sms);
    const int chains_per_sm[] = {32, 64, 128, 256, 512, 1024};
    const int logn[] = {16, 18};  // records per chain: 1 MB, 4 MB
    for (int ln : logn)
        for (int cps : chains_per_sm)
            for (int rec32 = 0; rec32 < 2; ++rec32)
                for (int store = 0; store < 2; ++store) {
                    const size_t n = (size_t)1 << ln;
                    const int C = cps * sms;
                    const size_t bytes = n * (size_t)C * sizeof(uint4);
                    if (bytes > max_gb * 1e9) continue;
                    if (rec32 && store == 0) continue;
                    uint4 *buf;
                    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc fail %zu\n", bytes); cudaGetLastError(); continue; }
                    init<<<sms * 8, 1024>>>(buf, n, n * (size_t)C);
                    const int threads = cps < 1024 ? (cps < 128 ? cps : 128) : 1024;  // CTA size; grid covers C lanes
                    const int hops = 4000;
                    chase<<<(C + threads - 1) / threads, threads>>>(buf, n, C, 400, store, rec32, out);  // warm
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0);
                    cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    chase<<<(C + threads - 1) / threads, threads>>>(buf, n, C, hops, store, rec32, out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    printf("chains=%6d (%4d/SM) rec=%2dB footprint/chain=%5zu KB total=%6.1f GB store=%d : %7.1f ns/hop, %.2e hops/s  (%s)\n", C, cps,
                           rec32 ? 32 : 16, n * 16 / 1024, bytes / 1e9, store, ms * 1e6 / hops, (double)C * hops / (ms * 1e-3),
                           cudaGetErrorString(cudaGetLastError()));
                    fflush(stdout);
                    cudaFree(buf);
                }
    return 0;
}
