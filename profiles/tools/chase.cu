// chase.cu — pointer-chase latency floor on B200: one warp = one dependent chain of 16-byte ld.global.cg
// loads (uniform address per warp), W warps in flight, per-chain footprint n records of 16 B.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chase chase.cu ;  run: ./chase
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void init(uint4 *buf, size_t n, int W) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n * W) return;
    uint32_t k = (uint32_t)(i % n);
    uint32_t nxt = (uint32_t)((1664525ull * k + 1013904223ull) % n);  // full-period LCG when n is a power of two
    uint32_t nn = (uint32_t)((1664525ull * nxt + 1013904223ull) % n);
    uint32_t n3 = (uint32_t)((1664525ull * nn + 1013904223ull) % n);
    buf[i] = make_uint4(k, nxt, nn, n3);
}
__global__ void chase(uint4 *buf, size_t n, int W, int hops, int store, unsigned long long *out) {
    int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= W) return;
    uint4 *base = buf + (size_t)w * n;
    __shared__ uint32_t sm[256];
    uint32_t k = (w * 7919u) % (uint32_t)n, kprev = 0, vprev = 0;
    long long t0 = clock64();
    for (int h = 0; h < hops; ++h) {
        uint4 R;
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(R.x), "=r"(R.y), "=r"(R.z), "=r"(R.w) : "l"(base + k));
        if (store == 1) asm volatile("st.global.u32 [%0], %1;" ::"l"(base + k), "r"(R.x + 1) : "memory");
        if (store == 2) {  // delayed by one hop: the store of hop h-1 is issued after the load of hop h
            if (h) asm volatile("st.global.u32 [%0], %1;" ::"l"(base + kprev), "r"(vprev) : "memory");
            kprev = k; vprev = R.x + 1;
        }
        if (store == 3) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(base + k), "r"(R.x + 1) : "memory");
        if (store == 4) { if ((threadIdx.x & 31) == 0) asm volatile("st.global.u32 [%0], %1;" ::"l"(base + k), "r"(R.x + 1) : "memory"); }
        if (store == 5) { if ((threadIdx.x & 31) == 0) asm volatile("st.global.u32 [%0], %1;" ::"l"(base + k), "r"(R.x + 1) : "memory"); __syncwarp(); }
        if (store == 7 || store == 8) {  // store + L2 prefetch of the record two (7) / three (8) hops ahead
            asm volatile("st.global.u32 [%0], %1;" ::"l"(base + k), "r"(R.x + 1) : "memory");
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (store == 7 ? R.z : R.w)));
        }
        if (store == 6) { uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + 4 * (k & 255); asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(R.x + 1) : "memory"); }
        k = R.y;
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, (unsigned long long)(t1 - t0)); atomicAdd(out + 1, (unsigned long long)k); }
}
int main() {
    int Ws[] = {4096, 8192};
    size_t ns[] = {1 << 16};  // records per chain: 64 KB, 1 MB
    unsigned long long *out;
    cudaMalloc(&out, 16);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (size_t n : ns) for (int W : Ws) for (int store = 0; store < 9; ++store) {
        uint4 *buf; size_t bytes = n * W * sizeof(uint4);
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc fail\n"); continue; }
        init<<<(unsigned)((n * W + 255) / 256), 256>>>(buf, n, W);
        int hops = 20000;
        chase<<<(W + 3) / 4, 128>>>(buf, n, W, 2000, store, out);  // warm
        cudaMemset(out, 0, 16);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        chase<<<(W + 3) / 4, 128>>>(buf, n, W, hops, store, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("W=%5d footprint/chain=%5zu KB total=%7.1f MB store=%d : %.1f ns/hop (wall), %.0f cycles/hop (clock64), %.2e hops/s\n",
               W, n * 16 / 1024, bytes / 1e6, store, ms * 1e6 / hops, (double)h[0] / W / hops, (double)W * hops / (ms * 1e-3));
        cudaFree(buf);
    }
    return 0;
}
