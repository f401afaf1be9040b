/* sse_rng.h — the random-stream contract of the B200 SSE sweep backend.
 *
 * The reference draws from Carlo's `ctx.rng` (Julia `Random.Xoshiro`, call sites
 * /root/reference/src/sse.jl:48,152,166,178,222,242,243,251,282); that stream is not pinned by any
 * reference test, so bit-exactness is defined on "draw k" instead:
 *
 *   - every walker owns ONE sequential stream of raw 64-bit draws x_0, x_1, ...; the sweep consumes
 *     them in exactly the reference's order;
 *   - U()  = (x >> 11) * 2^-53              uniform double in [0,1)        (replaces rand(rng))
 *   - I(k) = 1 + mulhi64(x, k)              uniform integer in 1..k        (replaces rand(rng, 1:k))
 *   - production stream: Philox4x32-10 block j = philox(counter = (j_lo, j_hi, walker_lo, walker_hi),
 *     key = (seed_lo, seed_hi)) yields draws x_{2j} = words (0,1) and x_{2j+1} = words (2,3);
 *     debug/parity stream: an explicit uint64 array ("injected stream").
 *
 * Both the CUDA kernels and the CPU oracle include this file, so the two sides cannot drift.
 * Also here: a portable tanh built from IEEE +,-,*,/ only (no FMA contraction: compile with
 * nvcc -fmad=false / gcc -ffp-contract=off) so the worm-count controller (sse.jl:204-217) is
 * bit-identical on host and device.
 */
#ifndef SSE_RNG_H
#define SSE_RNG_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SSE_HD __host__ __device__ __forceinline__
#else
#define SSE_HD static inline
#endif

SSE_HD uint32_t sse_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

SSE_HD uint64_t sse_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

/* Philox4x32-10 (Salmon et al., SC'11): block j of walker `walker` under key `seed`;
 * counter = (j_lo, j_hi, walker_lo, walker_hi), key = (seed_lo, seed_hi). */
SSE_HD void sse_philox_block(uint64_t seed, uint64_t walker, uint64_t j, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)j, c1 = (uint32_t)(j >> 32), c2 = (uint32_t)walker, c3 = (uint32_t)(walker >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = sse_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = sse_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw k of the production stream: one Philox block yields two draws,
 * x_{2j} = words (0,1) and x_{2j+1} = words (2,3) of block j. */
SSE_HD uint64_t sse_philox_draw(uint64_t seed, uint64_t walker, uint64_t k) {
    uint32_t b[4];
    sse_philox_block(seed, walker, k >> 1, b);
    return (k & 1) ? ((uint64_t)b[2] | ((uint64_t)b[3] << 32)) : ((uint64_t)b[0] | ((uint64_t)b[1] << 32));
}

/* U(): uniform double in [0,1) from a raw draw. */
SSE_HD double sse_u01(uint64_t x) { return (double)(x >> 11) * 0x1.0p-53; }

/* I(k)-1: uniform integer in 0..k-1 from a raw draw (the reference's 1-based value minus one). */
SSE_HD uint64_t sse_uint_below(uint64_t x, uint64_t k) { return sse_mulhi64(x, k); }

/* I(k)-1 for k < 2^32, the same value as sse_uint_below with two 32x32 multiplications:
 * x*k = xh*k*2^32 + xl*k, so its upper 64 bits are (xh*k + (xl*k >> 32)) >> 32 (no carry can be lost: xh*k <= (2^32-1)^2). */
SSE_HD uint32_t sse_uint_below32(uint64_t x, uint32_t k) {
    const uint64_t lo = (uint64_t)(uint32_t)x * (uint64_t)k, hi = (x >> 32) * (uint64_t)k;
    return (uint32_t)((hi + (lo >> 32)) >> 32);
}

/* exp(y) for y in [-40, 0], plain IEEE arithmetic, identical on host and device. */
SSE_HD double sse_exp_neg(double y) {
    const double LOG2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01,
                 LN2_LO = 1.90821492927058770002e-10;
    double kf = floor(y * LOG2E + 0.5);
    double r = (y - kf * LN2_HI) - kf * LN2_LO;
    /* Taylor/Horner to r^14, |r| <= 0.35 -> truncation error < 1e-18 */
    double p = 1.0 / 87178291200.0;
    p = p * r + 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    return ldexp(p, (int)kf);
}

/* tanh(x), absolute error ~1e-16, bit-identical on host and device (see header comment). */
SSE_HD double sse_tanh(double x) {
    double ax = fabs(x);
    if (!(ax < 20.0)) return x < 0 ? -1.0 : 1.0;
    double t = sse_exp_neg(-2.0 * ax);
    double v = (1.0 - t) / (1.0 + t);
    return x < 0 ? -v : v;
}

#endif /* SSE_RNG_H */
