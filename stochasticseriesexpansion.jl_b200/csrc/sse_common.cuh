// sse_common.cuh — data layout and small helpers shared by the device code of the B200 SSE sweep backend (sm_100a).
//
// Execution model (round 2): ONE persistent CTA per SM (sse::k_sweep, sse_sweep.cuh) owns a fixed subset of the walkers.
// Its warps have two roles:
//   * worm warps    — ONE LANE advances ONE walker's worm update (src/sse.jl:193-303): 32 independent dependent-load
//                     chains per warp with divergent addresses, per-lane Philox stream, per-lane record registers
//                     (sse_worm.cuh).  The chase is bound by DRAM latency and the random-sector rate, not by issue
//                     slots, so a few worm warps per SM carry every chain the memory system can serve.
//   * stream warps  — ONE WARP advances ONE walker through the streaming phases: the end of worm_update (controller +
//                     state rebuild), Carlo.measure!, diagonal_update fused with make_vertex_list! (sse_stream.cuh).
// Walkers move between the two roles through a per-CTA status table in shared memory; nothing synchronises walkers
// with each other, and a walker can be parked anywhere in its worm phase when its visit budget for the launch runs out
// (a monster worm only delays its own walker).
//
// Per-walker data in HBM:
//   words[W][Mw_cap]  uint2  {bits, rank}: occupancy bitmap of 32 slots of the padded operator string (bit p%32 of word
//                            p/32 set <=> slot p holds a non-identity operator) and the number of non-identity slots
//                            before the word.  Replaces the reference's 8 B/slot `operators` (src/sse.jl:12) for
//                            identity slots: 0.25 B per slot.  A worm start (sse.jl:241-247) is ONE 8-byte load.
//   rec  [W][R_cap]   uint4  16-byte vertex record of the k-th non-identity operator: {op code, 4 x 24-bit leg links
//                            (k' << 2 | leg')}.  Replaces the 8 B op code + 64 B/slot `vertices` (vertex_list.jl:1-13).
//                            The record array is a RING: generation g occupies [G, G + n) mod R_cap, and the diagonal
//                            update writes generation g+1 right behind it ([G + n, ...)), reading the old op codes just
//                            ahead of its own write head, so no second buffer exists.
//   state[W][N] u8, vfirst[W][N] u32 (link of the first leg on each site's world line; NONE32 = no operator),
//   vlast[W][N] u32 (scratch of the record build), ctl[W] = every scalar of the walker (WalkerCtl, 128 B).
// Device op code (u32): bit0 = 1, bit1 = diagonal, bits 2..13 = global vertex id, bits 14..31 = bond.
//
// Every phase reproduces the reference's draw ORDER and its Float64 expressions (compile with
// -fmad=false), so results are bit-identical to the CPU oracle under the same random stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sse_b200.h"
#include "../../include/sse_rng.h"

namespace sse {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t NONE32 = 0xffffffffu;
constexpr uint32_t NONE24 = 0x00ffffffu;
constexpr int VBITS = 12;                              // global vertex id bits in the device op code
constexpr uint32_t VMASK = ((1u << VBITS) - 1u) << 2;  // bits 2..13
constexpr int BOND_SHIFT = 2 + VBITS;                  // 14
constexpr int RNG_WORDS = 66;                          // 33 Philox blocks x 2 draws (see phase_diag_build)
constexpr int PHASE_WARPS = 4;                         // warps per CTA of sse::k_phase (one warp = one walker)
// CTA shape of sse::k_sweep.  Default: 16 warps (launch bounds 512 x 1 = 128 registers per thread), roles contiguous.
// SSE_SWEEP_SPLIT_REGS=1 builds the setmaxnreg variant: 24 warps = 8 worm warps + 16 stream warps launched at 80 registers;
// the worm warpgroups shrink to SSE_WORM_REGS, the stream warpgroups grow to SSE_STREAM_REGS (8*32*48 + 16*32*96 = 768*80).
// Measured on B200 (profiles/r2_l_l2policy_split.txt): the split loses 15-20 %, because at 48 registers the chase loop
// spills onto its critical path and the extra stream warps add memory contention that lengthens every worm visit.
#ifndef SSE_SWEEP_SPLIT_REGS
#define SSE_SWEEP_SPLIT_REGS 0
#endif
#if SSE_SWEEP_SPLIT_REGS
constexpr bool SPLIT_REGS = true;
constexpr int WORM_GROUP_WARPS = 8, STREAM_GROUP_WARPS = 16;
constexpr int SWEEP_MAX_WARPS = WORM_GROUP_WARPS + STREAM_GROUP_WARPS;
#ifndef SSE_WORM_REGS
#define SSE_WORM_REGS 48
#define SSE_STREAM_REGS 96
#endif
#else
#ifndef SSE_SWEEP_MAX_WARPS
#define SSE_SWEEP_MAX_WARPS 16
#endif
constexpr bool SPLIT_REGS = false;
constexpr int SWEEP_MAX_WARPS = SSE_SWEEP_MAX_WARPS;   // 16: launch bounds 512 x 1 = 128 registers per thread
constexpr int WORM_GROUP_WARPS = 8, STREAM_GROUP_WARPS = SWEEP_MAX_WARPS - 1;
#define SSE_WORM_REGS 128
#define SSE_STREAM_REGS 128
#endif
constexpr int ROT_MARGIN = 224;                        // ring slack between the write head and unread old records (>= 32 * (OP_AHEAD + 2))

__host__ __device__ __forceinline__ uint32_t op_pack(uint32_t bond, uint32_t gv, uint32_t diag) {
    return 1u | (diag << 1) | (gv << 2) | (bond << BOND_SHIFT);
}
__host__ __device__ __forceinline__ uint32_t op_gv(uint32_t op) { return (op >> 2) & ((1u << VBITS) - 1u); }
__host__ __device__ __forceinline__ uint32_t op_bond(uint32_t op) { return op >> BOND_SHIFT; }

// Shared-memory image of the vertex tables (built once on the host, copied per CTA).
struct TabLayout {
    int bytes;
    int off_t1;       // uint4  [nv*max_worm*4] first outcome fused with the transition header:
                      //        {cumprob0 lo, cumprob0 hi, packed step0, dim_out << 24 | offset of outcome 1 << 6 | remaining count}
    int off_outc;     // uint4  [n_outcomes] {cumprob lo, cumprob hi, packed step, dim_out << 24}
    int off_weights;  // double [nv]
    int off_vinfo;    // u32    [nv]  leg states packed, 8 bits per leg
    int off_diagv;    // u16    [n_diag]  global vertex id + 1, 0 = invalid
    int off_vneg;     // u8     [nv]  1 if the vertex sign is negative
    int off_estrows;  // double [n_est][est_nrows][est_max_dim]  the distinct rows of est_values (fused measurement), or -1
    int est_nrows;
};
// packed step (t1[].z / outc[].z): bits 1..13 = vertex bits of the op code (diag << 1 | gv << 2), bits 16..17 = exit leg,
// bits 24..31 = exit worm;  .w: bits 24..31 = dim of the exit leg's site, (t1 only) bits 6..23 = offset of the 2nd outcome,
// bits 0..5 = number of further outcomes.
struct SmTab {
    const uint4 *t1;
    const uint4 *outc;
    const double *weights;
    const uint32_t *vinfo;
    const uint16_t *diagv;
    const uint8_t *vneg;
    const double *estrows;  // nullptr if the estimator tables do not compress (see sse_model_create)
    uint32_t t1_s, outc_s;  // shared-space addresses of t1 / outc
};

struct DevModel {
    int n_sites, n_bonds, nv, max_worm, n_est, est_max_dim, norm_sites;
    double energy_offset;
    const uint4 *bond_info;   // [n_bonds] {site_a | dim_a << 24, site_b | dim_b << 24, diag table base,
                              //            row of est_values of (estimator 0, site a) | (0, b) << 8 | (1, a) << 16 | (1, b) << 24}
    const uint8_t *site_dim;  // [n_sites]
    const double *est_values; // [n_est][n_sites][est_max_dim]
    const uint8_t *tab_blob;  // TabLayout image
    TabLayout tl;
};

// Every scalar of one walker: the reference's `MC` fields (src/sse.jl:6-24), the stream position, and the progress of
// the sweep in flight (so a launch can stop anywhere in the worm phase and the next one resumes there).
struct __align__(16) WalkerCtl {
    double T;
    double num_worms;
    double avg_wl;
    double last_wlf;
    unsigned long long draws;         // stream position
    unsigned long long sweep_visits;  // sum of the lengths of this sweep's finished worms
    unsigned long long worm_len;      // parked worm: length so far (worm_traverse!'s `worm_length`)
    unsigned long long budget_left;   // worm visits this walker may still do in the current launch series
    unsigned long long sweeps_done;   // completed sweeps since sse_init / sse_set_state
    int M, n;
    uint32_t G;                       // ring position of record 0 of the current generation
    uint32_t flags;
    uint32_t phase;                   // 0 = between sweeps, 1 = worm phase of a sweep in progress
    uint32_t worms_left;              // worms of this sweep not yet finished (incl. a parked one)
    uint32_t inworm;                  // 1 = parked in the middle of a worm: pos / wf / pos0 / w0 / worm_len are valid
    uint32_t pos, wf, pos0, w0;       // parked worm: current leg link, current worm, start leg link, start worm
    int32_t sweeps_left;              // sweeps still to do in the current sse_sweep call
    uint32_t pad_[2];
};
static_assert(sizeof(WalkerCtl) == 128, "WalkerCtl is one 128-byte line");

struct DevWalkers {
    int W;
    int64_t M_cap;               // slots (multiple of 32)
    int64_t Mw_cap;              // words = M_cap / 32
    int64_t n_cap;               // max non-identity operators
    int64_t R_cap;               // ring size in records (> n_cap)
    uint2 *words;
    uint4 *rec;
    uint8_t *state;
    uint8_t *mark;               // [W][N] scratch, only when a stream warp's arrays do not fit in shared memory
    uint32_t *vfirst, *vlast;
    WalkerCtl *ctl;
    double *acc;                 // [W][n_obs]
    long long *acc_cnt;          // [W][2]
    unsigned long long *counters;// [SSE_N_COUNTERS]
    long long *dbg_len;          // [W] worm length of the last sse_dbg_worm_traverse
    double *obs_out;             // [W][n_obs] scratch for sse_measure
    const unsigned long long *inj;
    long long inj_len;
    unsigned long long seed, wid_off;
    double twlf, atten;
    int n_obs;
};

constexpr uint32_t FATAL_FLAGS = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;

// Context of one walker inside a streaming phase (uniform across the warp).
struct Ctx {
    uint2 *words;
    uint4 *rec;
    uint8_t *state;              // generic pointer: shared memory or the global array
    uint8_t *mark;
    unsigned long long *rng;     // per-warp shared scratch, RNG_WORDS entries
    uint32_t *queue;             // per-warp shared scratch: 3 x 64 words, operators waiting for the record build
    uint32_t opring_s, biring_s; // shared-space addresses of the warp's prefetch rings (op codes, bond-table rows)
    uint32_t *vfirst, *vlast;
    const unsigned long long *inj;
    long long inj_len;
    unsigned long long seed, wid, draws;
    double T, num_worms, avg_wl, last_wlf;
    int M, n;
    uint32_t G, Rcap, flags, lane;
};

// The inline-PTX helpers below are the only non-C++ code of the device side; the test-only warp emulator
// (tests/emu/cuda_emu.h) provides host versions and defines SSE_PTX_HELPERS_PROVIDED.
#ifndef SSE_PTX_HELPERS_PROVIDED
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// L2 eviction policies.  The record ring is touched in two ways: the worm phase reads and rewrites random sectors that
// nobody needs again soon (evict_first), the record build writes records whose forward links are patched a few hundred
// microseconds later (evict_last: a patch that finds its sector in L2 is merged there; one that does not costs a DRAM
// read-modify-write — measured: as much DRAM traffic as the whole worm phase).
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// per-lane accesses of the worm phase: every lane touches its own walker
__device__ __forceinline__ uint4 lane_ld128(const uint4 *p, unsigned long long pol) {
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint2 lane_ld64(const uint2 *p) {
    uint2 v;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lane_st32(void *p, uint32_t v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
// stores of the record build
__device__ __forceinline__ void st128_hint(uint4 *p, uint4 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st16_hint(void *p, uint32_t v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(p), "h"((unsigned short)v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st8_hint(void *p, uint32_t v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}
__device__ __forceinline__ void st_volatile_shared(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void backoff(unsigned ns) { __nanosleep(ns); }
// Asynchronous global -> shared copies (LDGSTS) for the prefetch queues of the streaming pass: the data never sits in a
// register while in flight, so no register move or scoreboard wait can stall on it, and the groups complete in order.
// pred = false writes zeros without reading.
__device__ __forceinline__ void cp_async4(uint32_t dst_s, const void *src, bool pred, unsigned long long pol) {
    const int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2, %3;" ::"r"(dst_s), "l"(src), "r"(sz), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void *src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_s), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// register re-balancing between the warpgroups of a CTA (every warp of a warpgroup must execute the same one)
template <int N>
__device__ __forceinline__ void regs_shrink() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void regs_grow() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
#endif

// ring position of logical record k of the generation that starts at G
__device__ __forceinline__ uint32_t ring(uint32_t G, uint32_t Rcap, uint32_t k) {
    const uint32_t i = G + k;
    return i >= Rcap ? i - Rcap : i;
}

// leg link j (24 bits) of a record: bits [24j, 24j+24) of the 96-bit little-endian field (y, z, w)
__device__ __forceinline__ uint32_t rec_link(const uint4 &r, uint32_t j) {
    const uint32_t lo = (j < 2) ? r.y : ((j == 2) ? r.z : r.w);
    const uint32_t hi = (j < 2) ? r.z : r.w;
    return __funnelshift_r(lo, hi, (24u * j) & 31u) & NONE24;
}
__host__ __device__ __forceinline__ uint4 rec_pack(uint32_t op, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3) {
    uint4 r;
    r.x = op;
    r.y = l0 | (l1 << 24);
    r.z = (l1 >> 8) | (l2 << 16);
    r.w = (l2 >> 16) | (l3 << 8);
    return r;
}
// overwrite the leg link named by `target` (k << 2 | leg) of the generation at G with `value`: the 3-byte field starts
// at byte 4 + 3*leg of the record, so it is one aligned 16-bit store and one byte store
__device__ __forceinline__ void rec_patch(uint4 *rec, uint32_t G, uint32_t Rcap, uint32_t target, uint32_t value, unsigned long long pol) {
    uint8_t *b = reinterpret_cast<uint8_t *>(rec + ring(G, Rcap, target >> 2)) + 4u + 3u * (target & 3u);
    if (target & 1u) {  // odd offset: byte, then aligned half-word
        st8_hint(b, value & 0xffu, pol);
        st16_hint(b + 1, value >> 8, pol);
    } else {            // even offset: aligned half-word, then byte
        st16_hint(b, value & 0xffffu, pol);
        st8_hint(b + 2, value >> 16, pol);
    }
}

__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(FULL, lo, src);
    hi = __shfl_sync(FULL, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_f64(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(FULL, lo, d);
    hi = __shfl_up_sync(FULL, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        int lo = __double2loint(v), hi = __double2hiint(v);
        lo = __shfl_xor_sync(FULL, lo, d);
        hi = __shfl_xor_sync(FULL, hi, d);
        v += __hiloint2double(hi, lo);
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

__device__ __forceinline__ SmTab stage_tables(const DevModel &dm, uint8_t *smem) {
    const int n16 = dm.tl.bytes >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(dm.tab_blob);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    SmTab st;
    st.t1 = reinterpret_cast<const uint4 *>(smem + dm.tl.off_t1);
    st.outc = reinterpret_cast<const uint4 *>(smem + dm.tl.off_outc);
    st.weights = reinterpret_cast<const double *>(smem + dm.tl.off_weights);
    st.vinfo = reinterpret_cast<const uint32_t *>(smem + dm.tl.off_vinfo);
    st.diagv = reinterpret_cast<const uint16_t *>(smem + dm.tl.off_diagv);
    st.vneg = reinterpret_cast<const uint8_t *>(smem + dm.tl.off_vneg);
    st.estrows = dm.tl.off_estrows >= 0 ? reinterpret_cast<const double *>(smem + dm.tl.off_estrows) : nullptr;
    st.t1_s = (uint32_t)__cvta_generic_to_shared(st.t1);
    st.outc_s = (uint32_t)__cvta_generic_to_shared(st.outc);
    return st;
}

constexpr int OP_AHEAD = 3;              // the streaming pass requests op codes this many chunks ahead of their use
constexpr int OP_RING = OP_AHEAD + 1;    // chunks of op codes in flight (power of two)
static_assert((OP_RING & (OP_RING - 1)) == 0, "OP_RING must be a power of two");
// draws + build queue (3 x 64 words) + op-code ring (OP_RING x 32 words) + bond-row ring (2 x 32 x 16 B)
constexpr int STREAM_FIXED_BYTES = (((RNG_WORDS * 8) + 15) & ~15) + 3 * 64 * 4 + OP_RING * 32 * 4 + 2 * 32 * 16;
// bytes of shared scratch of one streaming warp: random draws + (level >= 1) state[N], mark[N] + (level 2) vlast[N]
__host__ __device__ inline int stream_scratch_bytes(int n_sites, int level) {
    int b = STREAM_FIXED_BYTES;
    if (level >= 1) b += 2 * ((n_sites + 15) & ~15);
    if (level >= 2) b += 4 * ((n_sites + 3) & ~3);
    return b;
}

// Open / close the streaming context of walker w.  state[] is staged in the warp's shared scratch at level 1.
__device__ __forceinline__ Ctx ctx_open(const DevModel &dm, const DevWalkers &dw, int w, uint8_t *scratch, int level, uint32_t lane) {
    const int N = dm.n_sites;
    Ctx c;
    c.lane = lane;
    c.rng = reinterpret_cast<unsigned long long *>(scratch);
    c.queue = reinterpret_cast<uint32_t *>(scratch + (((RNG_WORDS * 8) + 15) & ~15));
    c.opring_s = (uint32_t)__cvta_generic_to_shared(c.queue + 3 * 64);
    c.biring_s = c.opring_s + OP_RING * 32 * 4;
    uint8_t *gstate = dw.state + (size_t)w * N;
    if (level) {
        c.state = scratch + STREAM_FIXED_BYTES;
        c.mark = c.state + ((N + 15) & ~15);
    } else {
        c.state = gstate;
        c.mark = dw.mark + (size_t)w * N;
    }
    c.words = dw.words + (size_t)w * dw.Mw_cap;
    c.rec = dw.rec + (size_t)w * dw.R_cap;
    c.vfirst = dw.vfirst + (size_t)w * N;
    c.vlast = level >= 2 ? reinterpret_cast<uint32_t *>(c.mark + ((N + 15) & ~15)) : dw.vlast + (size_t)w * N;
    c.inj = dw.inj ? dw.inj + (size_t)w * dw.inj_len : nullptr;
    c.inj_len = dw.inj_len;
    c.seed = dw.seed;
    c.wid = dw.wid_off + (unsigned long long)w;
    const WalkerCtl *ctl = dw.ctl + w;
    c.draws = __ldcg(&ctl->draws);
    c.T = __ldcg(&ctl->T);
    c.num_worms = __ldcg(&ctl->num_worms);
    c.avg_wl = __ldcg(&ctl->avg_wl);
    c.last_wlf = __ldcg(&ctl->last_wlf);
    c.M = __ldcg(&ctl->M);
    c.n = __ldcg(&ctl->n);
    c.G = __ldcg(&ctl->G);
    c.flags = __ldcg(&ctl->flags);
    c.Rcap = (uint32_t)dw.R_cap;
    for (int s = lane; s < N; s += 32) {
        if (level) c.state[s] = __ldcg(gstate + s);
        c.mark[s] = 0;
    }
    __syncwarp();
    return c;
}
__device__ __forceinline__ void ctx_close(const DevModel &dm, const DevWalkers &dw, int w, int level, const Ctx &c) {
    const int N = dm.n_sites;
    __syncwarp();
    if (level) {
        uint8_t *gstate = dw.state + (size_t)w * N;
        for (int s = c.lane; s < N; s += 32) gstate[s] = c.state[s];
    }
    if (c.lane == 0) {
        WalkerCtl *ctl = dw.ctl + w;
        ctl->draws = c.draws;
        ctl->T = c.T;
        ctl->num_worms = c.num_worms;
        ctl->avg_wl = c.avg_wl;
        ctl->last_wlf = c.last_wlf;
        ctl->M = c.M;
        ctl->n = c.n;
        ctl->G = c.G;
        ctl->flags = c.flags;
        if (c.flags & FATAL_FLAGS) atomicOr(reinterpret_cast<unsigned long long *>(dw.counters + SSE_CNT_ANY_FATAL), 1ull);
    }
    __syncwarp();
}

}  // namespace sse
