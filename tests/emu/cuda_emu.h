// cuda_emu.h — TEST INFRASTRUCTURE ONLY.  A warp-level emulator that lets g++ compile the product's CUDA
// translation unit (csrc/sse_capi.cu + sse_kernels.cuh, unchanged) into tests/emu/libsse_b200_emu.so, so the
// `-m "not gpu"` suite can run the real kernel source on the CPU and compare it bit for bit with the oracle.
// It is never loaded by the package (stochasticseriesexpansion.jl_b200/capi.py only knows csrc/libsse_b200.so);
// tests/test_emu_parity.py loads it explicitly.  It is not a fallback and not a performance model.
//
// Execution model: every CUDA thread of a CTA is a fiber (hand-rolled x86-64 context switch); CTAs run one
// after the other.  A fiber runs until it reaches a warp collective (__ballot_sync / __shfl*_sync / __syncwarp)
// or __syncthreads, publishes its operand and yields; the LAST participant to arrive computes every
// participant's result and releases them.  This is one legal interleaving under CUDA's independent thread
// scheduling, so code that is correct on the device is correct here; code that relies on implicit lock-step
// (a missing __syncwarp) or executes a collective under divergence (lanes named in the mask that never arrive,
// mismatching masks) fails loudly: the scheduler detects the deadlock and aborts with a per-lane report.
// Extras: shared memory and cudaMalloc memory are filled with garbage (uninitialised reads show up as
// mismatches), cudaMalloc blocks carry red zones that are verified on cudaFree and after every launch.
#pragma once
#if !defined(__x86_64__)
#error "tests/emu needs x86-64 (hand-rolled fiber switch)"
#endif

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

// ---- CUDA keywords ------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __shared__
#define __grid_constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint4 {
    uint32_t x, y, z, w;
} __attribute__((aligned(16)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct uint2 {
    uint32_t x, y;
} __attribute__((aligned(8)));
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};

namespace emu {

constexpr int MAX_SMEM = 232448;  // 227 KB, the sm_100 per-CTA limit
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t REDZONE = 4096;

enum Op { OP_NONE = 0, OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_SYNCWARP };

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    emu_dim3 tid;
    int lane = 0, warp = 0;
    bool done = false;
    // collective in flight
    int op = OP_NONE;
    uint32_t cmask = 0;
    uint64_t val = 0;
    int arg = 0, width = 32;
    bool arrived = false, released = false;
    uint64_t result = 0;
    // __syncthreads
    uint64_t bar_gen = 0;
    bool at_bar = false;
};

struct Cta {
    std::vector<Fiber> f;
    void *sched_sp = nullptr;
    uint64_t progress = 0;
    uint64_t bar_gen = 0;
    int bar_count = 0, n_exited = 0;
    void (*body)(void *) = nullptr;
    void *body_arg = nullptr;
};

extern Cta g_cta;
extern Fiber *g_cur;
extern emu_dim3 g_blockIdx, g_blockDim, g_gridDim;
extern uint64_t g_clock;
extern "C" void emu_switch(void **save_sp, void *load_sp);

[[noreturn]] void die(const char *fmt, ...);
void yield();
void poll_yield();  // a polling loop gives the other fibers a turn (aborts after too many fruitless polls)
uint64_t collective(int op, uint32_t mask, uint64_t val, int arg, int width);
void cta_barrier();
void run_grid(unsigned grid, unsigned block, size_t smem_bytes, void (*body)(void *), void *arg);
void check_redzones(const char *when);

}  // namespace emu

namespace sse {
extern uint8_t smem[] __attribute__((aligned(16)));  // the kernel's `extern __shared__ uint8_t smem[]`
}

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ---- intrinsics ---------------------------------------------------------------------------------------
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (s & 31u)) >> 32);
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
static inline int __double2loint(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return (int)(uint32_t)b;
}
static inline int __double2hiint(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return (int)(uint32_t)(b >> 32);
}
static inline double __hiloint2double(int hi, int lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &b, 8);
    return v;
}
static inline double __longlong_as_double(long long x) {
    double v;
    memcpy(&v, &x, 8);
    return v;
}
static inline long long __double_as_longlong(double x) {
    long long v;
    memcpy(&v, &x, 8);
    return v;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
template <class T>
static inline T __ldcg(const T *p) { return *p; }
template <class T>
static inline T __ldcs(const T *p) { return *p; }
template <class T>
static inline void __stcg(T *p, T v) { *p = v; }
template <class T>
static inline void __stcs(T *p, T v) { *p = v; }
static inline long long clock64() { return (long long)(emu::g_clock += 7); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o + v;
    return o;
}
static inline int atomicAdd(int *p, int v) {
    int o = *p;
    *p = o + v;
    return o;
}
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o | v;
    return o;
}
static inline uint32_t atomicSub(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o - v;
    return o;
}
static inline uint32_t atomicCAS(uint32_t *p, uint32_t cmp, uint32_t v) {
    uint32_t o = *p;
    if (o == cmp) *p = v;
    return o;
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) {
    uint32_t o = *p;
    *p = o | v;
    return o;
}
static inline size_t __cvta_generic_to_shared(const void *p) {
    const uint8_t *q = (const uint8_t *)p;
    if (q < sse::smem || q >= sse::smem + emu::MAX_SMEM) emu::die("__cvta_generic_to_shared: %p is not in shared memory", p);
    return (size_t)(q - sse::smem);
}

// warp collectives (any member mask; `width` as in CUDA)
static inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    return (uint32_t)emu::collective(emu::OP_BALLOT, mask, pred ? 1u : 0u, 0, 32);
}
static inline uint32_t __activemask() { return 0xffffffffu; }
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::collective(emu::OP_SYNCWARP, mask, 0, 0, 32); }
static inline void __syncthreads() { emu::cta_barrier(); }
#define EMU_SHFL(T)                                                                                              \
    static inline T __shfl_sync(uint32_t m, T v, int src, int width = 32) {                                      \
        return (T)emu::collective(emu::OP_SHFL, m, (uint64_t)v, src, width);                                     \
    }                                                                                                            \
    static inline T __shfl_up_sync(uint32_t m, T v, unsigned d, int width = 32) {                                \
        return (T)emu::collective(emu::OP_SHFL_UP, m, (uint64_t)v, (int)d, width);                               \
    }                                                                                                            \
    static inline T __shfl_down_sync(uint32_t m, T v, unsigned d, int width = 32) {                              \
        return (T)emu::collective(emu::OP_SHFL_DOWN, m, (uint64_t)v, (int)d, width);                             \
    }                                                                                                            \
    static inline T __shfl_xor_sync(uint32_t m, T v, int d, int width = 32) {                                    \
        return (T)emu::collective(emu::OP_SHFL_XOR, m, (uint64_t)v, d, width);                                   \
    }
EMU_SHFL(int)
EMU_SHFL(uint32_t)
EMU_SHFL(unsigned long long)
EMU_SHFL(long long)
#undef EMU_SHFL

// ---- the product's inline-PTX helpers (sse_kernels.cuh guards its own with SSE_PTX_HELPERS_PROVIDED) ----
#define SSE_PTX_HELPERS_PROVIDED 1
namespace sse {
static inline uint32_t lanemask_lt() { return (1u << emu::g_cur->lane) - 1u; }
static inline uint8_t *emu_smem_at(uint32_t a, uint32_t bytes) {
    if ((size_t)a + bytes > (size_t)emu::MAX_SMEM || (a % bytes) != 0) emu::die("shared access at %u (+%u) misaligned or out of range", a, bytes);
    return smem + a;
}
static inline void emu_check_global(const void *p, size_t align) {
    if (((uintptr_t)p % align) != 0) emu::die("misaligned %zu-byte global access at %p", align, p);
}
static inline uint4 lds128(uint32_t a) { return *reinterpret_cast<const uint4 *>(emu_smem_at(a, 16)); }
// Per-lane accesses of the worm phase (every lane touches its own walker), the volatile status words of the
// scheduler, and the back-off of a polling loop (a yield: fibers only switch at collectives otherwise).
static inline unsigned long long policy_evict_first() { return 1; }
static inline unsigned long long policy_evict_last() { return 2; }
static inline uint4 lane_ld128(const uint4 *p, unsigned long long) {
    emu_check_global(p, 16);
    return *p;
}
static inline void st128_hint(uint4 *p, uint4 v, unsigned long long) {
    emu_check_global(p, 16);
    *p = v;
}
static inline void st16_hint(void *p, uint32_t v, unsigned long long) {
    emu_check_global(p, 2);
    *reinterpret_cast<uint16_t *>(p) = (uint16_t)v;
}
static inline void st8_hint(void *p, uint32_t v, unsigned long long) { *reinterpret_cast<uint8_t *>(p) = (uint8_t)v; }
static inline uint2 lane_ld64(const uint2 *p) {
    emu_check_global(p, 8);
    return *p;
}
static inline void lane_st32(void *p, uint32_t v, unsigned long long) {
    emu_check_global(p, 4);
    *reinterpret_cast<uint32_t *>(p) = v;
}
static inline uint32_t ld_volatile_shared(const uint32_t *p) {
    __cvta_generic_to_shared(p);
    return *reinterpret_cast<const volatile uint32_t *>(p);
}
static inline void st_volatile_shared(uint32_t *p, uint32_t v) {
    __cvta_generic_to_shared(p);
    *reinterpret_cast<volatile uint32_t *>(p) = v;
}
static inline void backoff(unsigned) { emu::poll_yield(); }
// asynchronous copies: performed at once (one legal timing); pred = false zero-fills
static inline void cp_async4(uint32_t dst_s, const void *src, bool pred, unsigned long long) {
    uint32_t v = 0;
    if (pred) {
        emu_check_global(src, 4);
        v = *reinterpret_cast<const uint32_t *>(src);
    }
    *reinterpret_cast<uint32_t *>(emu_smem_at(dst_s, 4)) = v;
}
static inline void cp_async16(uint32_t dst_s, const void *src, bool pred) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pred) {
        emu_check_global(src, 16);
        v = *reinterpret_cast<const uint4 *>(src);
    }
    *reinterpret_cast<uint4 *>(emu_smem_at(dst_s, 16)) = v;
}
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}
template <int N>
static inline void regs_shrink() {}
template <int N>
static inline void regs_grow() {}
static inline uint32_t lds32(uint32_t a) { return *reinterpret_cast<const uint32_t *>(emu_smem_at(a, 4)); }
}  // namespace sse

// ---- CUDA runtime stand-ins (host "device memory" = malloc with red zones) ------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorMemoryAllocation = 2;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
constexpr unsigned cudaStreamNonBlocking = 1;
static inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                            cudaStream_t) {
    for (size_t r = 0; r < height; ++r) memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {
    *v = 3;  // a small "GPU": several walkers share a CTA even in small tests
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { return cudaMemset(d, v, n); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = (cudaStream_t)0x1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int v) {
    return v <= emu::MAX_SMEM ? cudaSuccess : 1;
}

// kernel launch: sse_capi.cu routes its launches through SSE_LAUNCH_KERNEL
namespace emu {
template <class... A>
struct Thunk {
    void (*k)(A...);
    std::tuple<A...> args;
    static void run(void *self) {
        Thunk *t = static_cast<Thunk *>(self);
        std::apply(t->k, t->args);
    }
};
template <class... A, class... B>
void launch(void (*k)(A...), unsigned grid, unsigned block, size_t smem_bytes, B &&...b) {
    Thunk<A...> t{k, std::tuple<A...>(std::forward<B>(b)...)};
    run_grid(grid, block, smem_bytes, &Thunk<A...>::run, &t);
}
}  // namespace emu
#define SSE_LAUNCH_KERNEL(kern, grid, block, smem_bytes, stream, ...) \
    emu::launch(kern, (unsigned)(grid), (unsigned)(block), (size_t)(smem_bytes), __VA_ARGS__)
