// sse_kernels.cuh — device side of the B200 SSE sweep backend (sm_100a).
//
// One WARP advances one walker through  diagonal update + vertex-record build (K1K2)  ->  worm update (K3)
// ->  commit + estimators (K4C)  for many sweeps inside one persistent launch; walkers never synchronise
// with each other.  Per-walker data lives in HBM in SoA form:
//
//   ops  [W][M_cap]  u32   padded operator string.  "committed" mode: 0 = identity, else the op code
//                          (bit0 = 1, bit1 = diagonal, bits 2..13 = global vertex id, bits 14.. = bond);
//                          "indexed" mode (between K1K2 and K4C): non-identity slots hold k+1, the index of
//                          their vertex record, so a worm start needs ONE dependent load.
//   rec  [W][n_cap]  32 B  vertex records = one DRAM sector.  First 16 bytes: the four leg links as u32
//                          (k' << 2 | leg').  Second 16 bytes: the op code + 4 x 24-bit two-hop prefetch hints
//                          (where the worm most likely is two visits later if it leaves through that leg; used
//                          only for prefetch.global.L2).  Both halves arrive with one sector per worm visit,
//                          the visit's only store (the op code) goes back into the same sector.
//   state[W][N] u8, vfirst/vlast [W][N] u32 (link of the first / last leg on each site's world line).
//
// Shared memory per CTA: the vertex tables (staged once) + per warp: the walker's state[N], a mark[N] byte
// array used to detect same-site collisions inside a 32-slot chunk, and a 66-word random-draw scratch.
// Warp primitives do the scans: ballot/popc prefix sums give each slot its random-stream offset (2/1/0
// draws by pre-update slot type) and its compact record index; one Philox block per lane feeds a whole
// chunk; shuffles resolve same-site ordering only in the rare chunks where two operators share a site.
//
// Every phase reproduces the reference's draw ORDER and its Float64 expressions
// (no FMA contraction: compile with -fmad=false), so results are bit-identical to the CPU oracle under
// the same random stream.  Reference lines are cited at each phase.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sse_b200.h"
#include "../../include/sse_rng.h"

namespace sse {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t NONE32 = 0xffffffffu;
constexpr uint32_t NONE24 = 0x00ffffffu;
constexpr int VBITS = 12;                 // global vertex id bits in the device op code
constexpr uint32_t VMASK = ((1u << VBITS) - 1u) << 2;  // bits 2..13
constexpr int BOND_SHIFT = 2 + VBITS;     // 14
constexpr int WARPS_PER_CTA = 4;
constexpr int RNG_WORDS = 66;             // 33 Philox blocks x 2 draws (see phase_diag_build)

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
};
// packed step: see worm_traverse_loop
struct SmTab {
    const uint4 *t1;
    const uint4 *outc;
    const double *weights;
    const uint32_t *vinfo;
    const uint16_t *diagv;
    const uint8_t *vneg;
};

struct DevModel {
    int n_sites, n_bonds, nv, max_worm, n_est, est_max_dim, norm_sites;
    double energy_offset;
    const uint4 *bond_info;   // [n_bonds] {site_a | dim_a << 24, site_b | dim_b << 24, diag table base, type}
    const uint8_t *site_dim;  // [n_sites]
    const double *est_values; // [n_est][n_sites][est_max_dim]
    const uint8_t *tab_blob;  // TabLayout image
    uint32_t pred_exit;       // 2 bits per entrance leg: the most likely exit leg (prefetch hints only)
    uint32_t variant;         // tuning switches (env SSE_B200_VARIANT): 2 = no hint prefetch, 4 = no hint pass
    TabLayout tl;
};

struct DevWalkers {
    int W;
    int64_t M_cap, n_cap;
    uint32_t *ops;
    uint4 *rec;
    uint8_t *state;
    uint8_t *mark;               // [W][N] scratch, only used when the per-warp arrays do not fit in shared memory
    uint32_t *vfirst, *vlast;
    double *T;
    int *M, *n;
    double *num_worms, *avg_wl, *last_wlf;
    unsigned long long *draws;
    uint32_t *flags;
    double *acc;                 // [W][n_obs]
    long long *acc_cnt;          // [W][2]
    unsigned long long *counters;// [8] visits, walker-sweeps, sum n, sum M, cycles in K1K2 / K3 / K4C, spare
    long long *dbg_len;          // [W] worm length of the last sse_dbg_worm_traverse
    double *obs_out;             // [W][n_obs] scratch for sse_measure
    const unsigned long long *inj;
    long long inj_len;
    unsigned long long seed, wid_off;
    double twlf, atten;
    int n_obs;
    int smem_state;              // per-warp arrays in shared memory: 0 none, 1 state+mark, 2 state+mark+vlast
};

enum Mode : int {
    MODE_SWEEP = 0, MODE_INIT, MODE_DIAG, MODE_MAKE_VL, MODE_WORM_UPDATE, MODE_WORM_TRAVERSE, MODE_COMMIT, MODE_MEASURE
};

struct LaunchArgs {
    int mode, n_sweeps, thermalized, measure, warmup;
    int l0, w0;       // MODE_WORM_TRAVERSE (0-based leg, 1-based worm)
    long long p0;     // 0-based slot
    int indexed;      // MODE_MEASURE: string currently in indexed mode?
};

// per-walker context held in registers (uniform across the warp)
struct Ctx {
    uint32_t *ops;
    uint4 *rec;
    uint8_t *state;              // generic pointer: shared memory or the global array
    uint8_t *mark;
    unsigned long long *rng;     // per-warp shared scratch, RNG_WORDS entries
    uint32_t *vfirst, *vlast;
    const unsigned long long *inj;
    long long inj_len;
    unsigned long long seed, wid, draws;
    double T, num_worms, avg_wl, last_wlf;
    int M, n;
    uint32_t flags, lane;
    unsigned long long visits;
};

// The inline-PTX helpers below are the only non-C++ code of this file; the test-only warp emulator
// (tests/emu/cuda_emu.h) provides host versions and defines SSE_PTX_HELPERS_PROVIDED.
#ifndef SSE_PTX_HELPERS_PROVIDED
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif

template <bool INJ>
__device__ __forceinline__ uint64_t draw(const Ctx &c, unsigned long long k) {
    if (INJ) return (long long)k < c.inj_len ? (uint64_t)__ldg(c.inj + k) : 0ull;
    return sse_philox_draw(c.seed, c.wid, k);
}

// Fill the warp's scratch with the draws [2*j0, 2*j0 + 64): lane L computes Philox block j0 + L (two draws).
template <bool INJ>
__device__ __forceinline__ void fill_draws(const Ctx &c, unsigned long long j0) {
    if (!INJ) {
        uint32_t b[4];
        sse_philox_block(c.seed, c.wid, j0 + c.lane, b);
        reinterpret_cast<uint4 *>(c.rng)[c.lane] = make_uint4(b[0], b[1], b[2], b[3]);
    }
}
// Draw with absolute index k from the scratch filled by fill_draws(j0) (or from the injected stream).
template <bool INJ>
__device__ __forceinline__ uint64_t scratch_draw(const Ctx &c, unsigned long long j0, unsigned long long k) {
    if (INJ) return (long long)k < c.inj_len ? (uint64_t)__ldg(c.inj + k) : 0ull;
    return c.rng[k - 2ull * j0];
}

// leg link j of a record's first half (plain u32 words)
__device__ __forceinline__ uint32_t rec_sel(const uint4 &r, uint32_t j) {
    const uint32_t a = (j & 1u) ? r.y : r.x, b = (j & 1u) ? r.w : r.z;
    return (j & 2u) ? b : a;
}
// hint j (24 bits) of a record's second half: bits [24j, 24j+24) of the 96-bit little-endian field (y, z, w)
__device__ __forceinline__ uint32_t rec_link(const uint4 &r, uint32_t j) {
    uint32_t lo = (j < 2) ? r.y : ((j == 2) ? r.z : r.w);
    uint32_t hi = (j < 2) ? r.z : r.w;
    return __funnelshift_r(lo, hi, (24u * j) & 31u) & NONE24;
}
__device__ __forceinline__ uint4 rec_pack(uint32_t op, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3) {
    uint4 r;
    r.x = op;
    r.y = l0 | (l1 << 24);
    r.z = (l1 >> 8) | (l2 << 16);
    r.w = (l2 >> 16) | (l3 << 8);
    return r;
}
// overwrite link `leg` of record `k` (one aligned word)
__device__ __forceinline__ void rec_patch(uint4 *rec, uint32_t target_link, uint32_t value) {
    reinterpret_cast<uint32_t *>(rec + 2u * (target_link >> 2))[target_link & 3u] = value;
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

// ------------------------------------------------------------------------------------------------------
// K1K2: diagonal_update (src/sse.jl:137-191) fused with make_vertex_list! (src/vertex_list.jl:15-54).
// do_diag = false builds the records of the unchanged string (make_vertex_list! alone);
// build = false performs the diagonal update only (Carlo.init! warm-up sweeps, src/sse.jl:54-57).
// Input string: committed mode.  Output: indexed mode if build, else committed.
// ------------------------------------------------------------------------------------------------------
template <bool INJ>
__device__ void phase_diag_build(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool do_diag,
                                 bool build) {
    const uint32_t lane = c.lane, lt = lanemask_lt();
    const int N = dm.n_sites;
    if (do_diag && 2ll * (long long)c.n >= (long long)c.M) {  // n >= 0.5*M  (sse.jl:138)
        long long newM = (3ll * (long long)c.M) / 2 + 100;     // floor(1.5*M + 100) (sse.jl:143)
        if (newM > dw.M_cap) { c.flags |= SSE_FLAG_M_OVERFLOW; return; }
        c.M = (int)newM;  // slots beyond the old M are identity by invariant (sse.jl:144)
    }
    if (build) {
        for (int s = lane; s < N; s += 32) { c.vfirst[s] = NONE32; c.vlast[s] = NONE32; }
    }
    __syncwarp();
    const int M = c.M;
    const uint32_t Nb = (uint32_t)dm.n_bonds;
    const double p_make_bond_raw = (double)dm.n_bonds / c.T;   // sse.jl:147
    const double p_remove_bond_raw = c.T / (double)dm.n_bonds; // sse.jl:148
    int n = c.n;
    uint32_t kbase = 0;
    unsigned long long draws = c.draws;
    const int nchunks = (M + 31) >> 5;
    uint32_t op_next = ((int)lane < M) ? c.ops[lane] : 0u;

    for (int ch = 0; ch < nchunks; ++ch) {
        const int p = ch * 32 + (int)lane;
        const bool active = p < M;
        const uint32_t op = op_next;
        op_next = (p + 32 < M) ? c.ops[p + 32] : 0u;  // prefetch the next chunk
        const bool nonid = op != 0u;
        const bool is_id = active && !nonid;
        const bool is_dg = nonid && (op & 2u);
        const bool is_off = nonid && !(op & 2u);
        uint32_t bond = op_bond(op);
        const uint32_t gv = op_gv(op);
        uint32_t newop = op;
        double r = 0.0;
        uint32_t idm = 0, dgm = 0;

        if (do_diag) {
            // stream offsets: 2 draws per identity slot, 1 per diagonal operator, in slot order (Appendix A)
            idm = __ballot_sync(FULL, is_id);
            dgm = __ballot_sync(FULL, is_dg);
            const uint32_t D = 2u * __popc(idm) + __popc(dgm);
            if (D) {
                const unsigned long long my = draws + 2u * __popc(idm & lt) + __popc(dgm & lt);
                const unsigned long long j0 = draws >> 1;
                fill_draws<INJ>(c, j0);
                if (!INJ && (draws & 1ull) && D == 64u && lane == 0) {  // the one draw beyond 32 blocks
                    uint32_t b[4];
                    sse_philox_block(c.seed, c.wid, j0 + 32, b);
                    reinterpret_cast<uint4 *>(c.rng)[32] = make_uint4(b[0], b[1], b[2], b[3]);
                }
                __syncwarp();
                if (is_id) {
                    bond = (uint32_t)sse_uint_below(scratch_draw<INJ>(c, j0, my), Nb);  // rand(rng, 1:N_b) - 1 (sse.jl:152)
                    r = sse_u01(scratch_draw<INJ>(c, j0, my + 1));                      // sse.jl:166
                } else if (is_dg) {
                    r = sse_u01(scratch_draw<INJ>(c, j0, my));                          // sse.jl:178
                }
                draws += D;
            }
        }
        uint4 bi = make_uint4(0, 0, 0, 0);
        if (is_id || nonid) bi = __ldg(dm.bond_info + bond);
        const uint32_t sa = bi.x & NONE24, sb = bi.y & NONE24;

        if (do_diag) {
            // State seen by each identity slot = state at chunk start overridden by earlier off-diagonal
            // operators of this chunk (sse.jl:182-188).  Off-diagonal lanes tag their sites in mark[]; only if
            // an identity lane reads a tagged site, or two off-diagonal lanes share a site, the in-order
            // shuffle loop runs.
            const uint32_t offm = __ballot_sync(FULL, is_off);
            uint32_t s_a = 1, s_b = 1, ta = 0, tb = 0;
            if (offm) {
                const uint8_t tag = (uint8_t)(0x80u | lane);
                if (is_off) {
                    const uint32_t vi = st.vinfo[gv];
                    ta = (vi >> 16) & 0xffu;
                    tb = vi >> 24;
                    c.mark[sa] = tag;
                    c.mark[sb] = tag;
                }
                __syncwarp();
                bool hit = false;
                if (is_off) hit = (c.mark[sa] != tag) || (c.mark[sb] != tag);
                if (is_id) {
                    hit = ((c.mark[sa] | c.mark[sb]) & 0x80u) != 0;
                    s_a = c.state[sa];
                    s_b = c.state[sb];
                }
                const uint32_t anyhit = __ballot_sync(FULL, hit);
                __syncwarp();
                bool wa = is_off, wb = is_off;
                if (anyhit) {
                    for (uint32_t m = offm; m;) {
                        const int L = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t qa = __shfl_sync(FULL, sa, L), qb = __shfl_sync(FULL, sb, L);
                        const uint32_t qta = __shfl_sync(FULL, ta, L), qtb = __shfl_sync(FULL, tb, L);
                        if (is_id && (int)lane > L) {
                            if (sa == qa) s_a = qta;
                            if (sa == qb) s_a = qtb;
                            if (sb == qa) s_b = qta;
                            if (sb == qb) s_b = qtb;
                        }
                        if (is_off && (int)lane < L) {  // a later operator of the chunk overwrites this site
                            if (sa == qa || sa == qb) wa = false;
                            if (sb == qa || sb == qb) wb = false;
                        }
                    }
                }
                if (is_off) {
                    if (wa) c.state[sa] = (uint8_t)ta;
                    if (wb) c.state[sb] = (uint8_t)tb;
                    c.mark[sa] = 0;
                    c.mark[sb] = 0;
                }
            } else if (is_id) {
                s_a = c.state[sa];
                s_b = c.state[sb];
            }

            double w = 0.0;
            uint32_t gvnew = 0;
            if (is_id) {
                // join_idx (util.jl:15-23) -> diagonal vertex -> weight (sse.jl:156-162)
                const uint32_t cidx = bi.z + (s_a - 1u) + (bi.x >> 24) * (s_b - 1u);
                const uint32_t dv = st.diagv[cidx];
                if (dv) { gvnew = dv - 1u; w = st.weights[gvnew]; }
            } else if (is_dg) {
                w = st.weights[gv];
            }
            // Accept tests (sse.jl:164-166,176-178) depend on the running operator count n.  Within the chunk
            // n stays in [n - #diagonal, n + #identity]; both tests are monotone in n (IEEE division and
            // multiplication are monotone), so evaluating them at the two ends decides every lane whose draw is
            // not between the two thresholds.  Only if some lane is undecided (probability ~ 64/(M-n) per chunk)
            // the in-order recurrence is solved exactly by fixed-point iteration.
            const int n_lo = n - __popc(dgm), n_hi = n + __popc(idm);
            bool acc = false, amb = false;
            if (is_id) {
                const double pm_lo = p_make_bond_raw / (double)(M - n_lo);
                const double pm_hi = (M - n_hi > 0) ? p_make_bond_raw / (double)(M - n_hi) : __longlong_as_double(0x7ff0000000000000ll);
                acc = r < pm_lo * w;
                amb = !acc && (r < pm_hi * w);
            } else if (is_dg) {
                const double rw = r * w;
                acc = rw < (double)(M - n_hi + 1) * p_remove_bond_raw;
                amb = !acc && (rw < (double)(M - n_lo + 1) * p_remove_bond_raw);
            }
            uint32_t ins, rem;
            if (__ballot_sync(FULL, amb)) {
                // lane l only depends on lanes < l: after i rounds the first i lanes are final
                ins = 0;
                rem = 0;
                while (true) {
                    const int nl = n + __popc(ins & lt) - __popc(rem & lt);
                    bool a2 = false;
                    if (is_id) {
                        const double p_make_bond = p_make_bond_raw / (double)(M - nl);  // sse.jl:164
                        a2 = r < p_make_bond * w;                                        // sse.jl:166
                    } else if (is_dg) {
                        const double p_remove_bond = (double)(M - nl + 1) * p_remove_bond_raw;  // sse.jl:176-177
                        a2 = r * w < p_remove_bond;                                              // sse.jl:178
                    }
                    const uint32_t ins2 = __ballot_sync(FULL, is_id && a2), rem2 = __ballot_sync(FULL, is_dg && a2);
                    if (ins2 == ins && rem2 == rem) break;
                    ins = ins2;
                    rem = rem2;
                }
            } else {
                ins = __ballot_sync(FULL, is_id && acc);
                rem = __ballot_sync(FULL, is_dg && acc);
            }
            n += __popc(ins) - __popc(rem);
            if (is_id && ((ins >> lane) & 1u)) newop = op_pack(bond, gvnew, 1u);
            if (is_dg && ((rem >> lane) & 1u)) newop = 0u;
        }

        if (build) {
            const bool nn = newop != 0u;
            const uint32_t nm = __ballot_sync(FULL, nn);
            const uint32_t k = kbase + __popc(nm & lt);
            if ((long long)kbase + __popc(nm) > dw.n_cap) { c.flags |= SSE_FLAG_N_OVERFLOW; return; }
            // same-site collisions inside the chunk are rare: every operator tags its two sites, a lost tag
            // reveals a collision, and only then the nearest earlier / later operator on each site is searched
            uint32_t pa = NONE24, pb = NONE24, sua = NONE24, sub = NONE24;
            if (nn) {
                c.mark[sa] = (uint8_t)lane;
                c.mark[sb] = (uint8_t)lane;
            }
            __syncwarp();
            const bool lost_a = nn && c.mark[sa] != (uint8_t)lane, lost_b = nn && c.mark[sb] != (uint8_t)lane;
            if (__ballot_sync(FULL, lost_a || lost_b)) {
                // losers flag the contested sites; every operator on a flagged site takes part in the search
                __syncwarp();
                if (lost_a) c.mark[sa] = 0x7f;
                if (lost_b) c.mark[sb] = 0x7f;
                __syncwarp();
                const bool inv = nn && (c.mark[sa] == 0x7f || c.mark[sb] == 0x7f);
                for (uint32_t m = __ballot_sync(FULL, inv); m;) {
                    const int L = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t qa = __shfl_sync(FULL, sa, L), qb = __shfl_sync(FULL, sb, L);
                    const uint32_t qk = __shfl_sync(FULL, k, L) << 2;
                    if (nn && (int)lane > L) {
                        if (sa == qa) pa = qk | 2u;
                        if (sa == qb) pa = qk | 3u;
                        if (sb == qa) pb = qk | 2u;
                        if (sb == qb) pb = qk | 3u;
                    } else if (nn && (int)lane < L) {
                        if (sua == NONE24) { if (sa == qa) sua = qk; else if (sa == qb) sua = qk | 1u; }
                        if (sub == NONE24) { if (sb == qa) sub = qk; else if (sb == qb) sub = qk | 1u; }
                    }
                }
            }
            uint32_t ma = NONE32, mb = NONE32;
            if (nn) {
                if (pa == NONE24) ma = c.vlast[sa];
                if (pb == NONE24) mb = c.vlast[sb];
            }
            __syncwarp();
            if (nn) {
                const uint32_t me = k << 2;
                uint32_t bla = pa, blb = pb;
                if (pa == NONE24) {
                    if (ma != NONE32) { bla = ma; rec_patch(c.rec, ma, me); }  // vertices[s1,p1] = (s,p) (vertex_list.jl:36-38)
                    else c.vfirst[sa] = me;                                     // vertex_list.jl:40
                }
                if (pb == NONE24) {
                    if (mb != NONE32) { blb = mb; rec_patch(c.rec, mb, me | 1u); }
                    else c.vfirst[sb] = me | 1u;
                }
                if (sua == NONE24) c.vlast[sa] = me | 2u;  // vertex_list.jl:42
                if (sub == NONE24) c.vlast[sb] = me | 3u;
                c.rec[2u * k] = make_uint4(bla, blb, sua, sub);          // links still unknown stay NONE24 until patched
                c.rec[2u * k + 1u] = make_uint4(newop, 0u, 0u, 0u);      // hints are filled by phase_hints
                c.ops[p] = k + 1u;
            } else if (nonid) {
                c.ops[p] = 0u;  // removed diagonal operator (sse.jl:179)
            }
            kbase += __popc(nm);
        } else if (active && newop != op) {
            c.ops[p] = newop;
        }
        __syncwarp();
    }
    if (build) {
        // periodic closure (vertex_list.jl:46-51)
        for (int s = lane; s < N; s += 32) {
            const uint32_t f = c.vfirst[s];
            if (f != NONE32) {
                const uint32_t l = c.vlast[s];
                rec_patch(c.rec, f, l);
                rec_patch(c.rec, l, f);
            }
        }
        __syncwarp();
    }
    c.n = n;
    c.draws = draws;
}

// ------------------------------------------------------------------------------------------------------
// Prefetch hints (no counterpart in the reference; affects speed only, never results).  For every record k and
// exit leg j: the worm arrives at (k1, l1) = link_j(k); its most likely exit there is pred_exit[l1] (from the
// vertex tables, host-computed), so two visits later it most likely needs the record link_{pred_exit[l1]}(k1).
// Four independent random loads per lane are in flight, so this pass runs at memory-level parallelism 128.
// ------------------------------------------------------------------------------------------------------
__device__ void phase_hints(const DevModel &dm, Ctx &c) {
    const uint32_t n = (uint32_t)c.n, pe = dm.pred_exit;
    __syncwarp();
    for (uint32_t k = c.lane; k < n; k += 32) {
        const uint4 R = __ldcg(c.rec + 2u * k);
        const uint32_t l[4] = {R.x, R.y, R.z, R.w};
        uint32_t h[4];
        uint4 R1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) R1[j] = __ldcg(c.rec + 2u * (l[j] >> 2));
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = rec_sel(R1[j], (pe >> (2u * (l[j] & 3u))) & 3u);
        const uint4 H = rec_pack(0u, h[0], h[1], h[2], h[3]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(c.rec + 2u * k + 1u);  // word 0 is the op code: keep it
        dst[1] = H.y;
        dst[2] = H.z;
        dst[3] = H.w;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// worm_traverse! inner loop (src/sse.jl:262-303) with scatter (src/vertex_data.jl:106-125).
// All 32 lanes execute the chain uniformly; the lanes pre-compute the next 64 uniform draws in parallel
// (one Philox block each) into the warp's shared scratch.  Kept out of line with a minimal argument set so
// the chase loop owns its registers: per visit one 16 B global load, one 16 B shared load, one f64 compare,
// one 4 B store.
// ------------------------------------------------------------------------------------------------------
struct WormArgs {
    uint4 *rec;
    uint32_t t1_s, outc_s, rbuf_s;  // shared-space addresses
    const unsigned long long *inj;
    long long inj_len;
    unsigned long long seed, wid, draws;
    uint32_t maxw, lane, k0, l0, w0, fell, variant;
};

#ifndef SSE_PTX_HELPERS_PROVIDED
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
#endif

// uniform doubles for the draws [2*j0, 2*j0 + 64) -> rbuf[64]
template <bool INJ>
__device__ __forceinline__ void fill_u01(const WormArgs &a, unsigned long long j0) {
    uint64_t x0, x1;
    if (INJ) {
        const unsigned long long k = 2ull * (j0 + a.lane);
        x0 = (long long)k < a.inj_len ? (uint64_t)__ldg(a.inj + k) : 0ull;
        x1 = (long long)(k + 1) < a.inj_len ? (uint64_t)__ldg(a.inj + k + 1) : 0ull;
    } else {
        uint32_t b[4];
        sse_philox_block(a.seed, a.wid, j0 + a.lane, b);
        x0 = (uint64_t)b[0] | ((uint64_t)b[1] << 32);
        x1 = (uint64_t)b[2] | ((uint64_t)b[3] << 32);
    }
    sts_f64x2(a.rbuf_s + 16u * a.lane, sse_u01(x0), sse_u01(x1));
}

#ifndef SSE_PTX_HELPERS_PROVIDED
__device__ __forceinline__ uint4 ldg_cg128(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void stg_u32(void *p, uint32_t v) {
    asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

// packed step (t1[].z / outc[].z): bits 1..13 = vertex bits of the op code (diag << 1 | gv << 2),
// bits 16..17 = exit leg, bits 24..31 = exit worm;  .w: bits 24..31 = dim of the exit leg's site,
// (t1 only) bits 6..23 = offset of the 2nd outcome, bits 0..5 = number of further outcomes.
// loop-invariant part of a worm
struct WormConst {
    uint4 *rec;
    uint32_t t1_s, outc_s, maxw4, pos0, w0;
    bool pref;
};
// loop-carried part
struct WormVar {
    uint32_t pos, wf, len, fell;
    uint32_t patch, patch_val;  // the previous visit re-entered its own record: its first word is patch_val
};

// One visit (the body of the reference's `while true`, src/sse.jl:274-300).  Rc/Hc = the current record (already
// requested), Rn/Hn receive the next one.  The code is ordered so that the next record's load is ISSUED as early
// as the data dependences allow (record -> shared-memory transition entry -> compare -> link); the store, the
// stop tests, the prefetch and the bookkeeping execute in the shadow of that load.  Returns true when the worm closed.
__device__ __forceinline__ bool worm_visit(const WormConst &k, WormVar &v, const uint4 &Rc, const uint4 &Hc, uint4 &Rn,
                                           uint4 &Hn, const double r) {
    const uint32_t pos = v.pos;
    const uint32_t x = v.patch ? v.patch_val : Hc.x;  // Rc = the four links, Hc = {op code, hints}
    // transitions[leg_in, worm_in, vi] fused with its first outcome (vertex_data.jl:115-123)
    uint4 e = lds128(k.t1_s + 16u * (op_gv(x) * k.maxw4 + (((v.wf - 1u) << 2) | (pos & 3u))));
    if (!(r < __hiloint2double((int)e.y, (int)e.x))) {
        const uint32_t off = (e.w >> 6) & 0x3ffffu, cnt = e.w & 63u;
        bool hit = false;
        for (uint32_t j = 0; !hit && j < cnt; ++j) {  // first out with random < cumprob
            e = lds128(k.outc_s + 16u * (off + j));
            hit = r < __hiloint2double((int)e.y, (int)e.x);
        }
        if (!hit) v.fell = 1;  // vertex_data.jl:124; clamped to the last outcome
    }
    const uint32_t leg_out = (e.z >> 16) & 3u;
    const uint32_t posn = rec_sel(Rc, leg_out);  // (leg_in, p) = vertices[leg_out, p] (sse.jl:295)
    uint4 *const rn = k.rec + 2u * (posn >> 2);
    Rn = ldg_cg128(rn);
    Hn = ldg_cg128(rn + 1);
    // ---- everything below overlaps with the load ----
    const uint32_t newop = (x & ~(VMASK | 2u)) | (e.z & (VMASK | 2u));  // OperCode(bond, new_vertex) (sse.jl:285)
    stg_u32(k.rec + 2u * (pos >> 2) + 1u, newop);
    // two-hop hint of the leg we leave through.  (Tried and rejected on B200, see DESIGN.md: a real touch load
    // instead of the hint, three-hop hints, speculating on the first outcome, parking the stores in registers.)
    if (k.pref) prefetch_l2(k.rec + 2u * (rec_link(Hc, leg_out) >> 2));
    const uint32_t w_out = e.z >> 24, dim_out = e.w >> 24;
    const bool stop1 = (((pos & ~3u) | leg_out) == k.pos0) && (w_out + k.w0 == dim_out);  // sse.jl:288-290
    v.len += stop1 ? 0u : 1u;
    v.wf = w_out;
    v.patch = ((posn >> 2) == (pos >> 2)) ? 1u : 0u;  // the link re-enters this record: its load preceded the store
    v.patch_val = newop;
    v.pos = posn;
    const bool stop2 = (posn == k.pos0) && (w_out == k.w0);  // sse.jl:297-299
    return stop1 || stop2;
}

template <bool INJ>
__device__ __noinline__ uint32_t worm_traverse_loop(WormArgs &a) {
    WormConst k;
    k.rec = a.rec;
    k.t1_s = a.t1_s;
    k.outc_s = a.outc_s;
    k.maxw4 = a.maxw * 4u;
    k.pos0 = (a.k0 << 2) | a.l0;
    k.w0 = a.w0;
    k.pref = !(a.variant & 2u);
    const uint32_t rbuf_s = a.rbuf_s;
    WormVar v;
    v.pos = k.pos0;
    v.wf = k.w0;
    v.len = 1;
    v.fell = 0;
    v.patch = 0;
    v.patch_val = 0;
    unsigned long long j0 = a.draws >> 1;
    uint32_t ri = (uint32_t)(a.draws & 1ull);  // index into rbuf (draw 2*j0 + ri)
    uint4 R0 = ldg_cg128(k.rec + 2u * (v.pos >> 2)), H0 = ldg_cg128(k.rec + 2u * (v.pos >> 2) + 1u), R1, H1;
    __syncwarp();
    fill_u01<INJ>(a, j0);
    __syncwarp();
    while (true) {
        // two visits per iteration so the record registers ping-pong without copies
        if (ri >= 64) {
            j0 += 32;
            ri = 0;
            __syncwarp();
            fill_u01<INJ>(a, j0);
            __syncwarp();
        }
        if (worm_visit(k, v, R0, H0, R1, H1, lds_f64(rbuf_s + 8u * ri++))) break;  // rand(rng) (sse.jl:282)
        if (ri >= 64) {
            j0 += 32;
            ri = 0;
            __syncwarp();
            fill_u01<INJ>(a, j0);
            __syncwarp();
        }
        if (worm_visit(k, v, R1, H1, R0, H0, lds_f64(rbuf_s + 8u * ri++))) break;
    }
    a.draws = 2ull * j0 + ri;
    a.fell = v.fell;
    return v.len;
}

template <bool INJ>
__device__ __forceinline__ uint32_t worm_traverse(const SmTab &st, const DevModel &dm, Ctx &c, const uint32_t k0,
                                                  const uint32_t l0, const uint32_t w0) {
    WormArgs a;
    a.rec = c.rec;
    a.t1_s = (uint32_t)__cvta_generic_to_shared(st.t1);
    a.outc_s = (uint32_t)__cvta_generic_to_shared(st.outc);
    a.rbuf_s = (uint32_t)__cvta_generic_to_shared(c.rng);
    a.inj = c.inj;
    a.inj_len = c.inj_len;
    a.seed = c.seed;
    a.wid = c.wid;
    a.draws = c.draws;
    a.maxw = (uint32_t)dm.max_worm;
    a.lane = c.lane;
    a.k0 = k0;
    a.l0 = l0;
    a.w0 = w0;
    a.fell = 0;
    a.variant = dm.variant;
    const uint32_t len = worm_traverse_loop<INJ>(a);
    c.draws = a.draws;
    if (a.fell) c.flags |= SSE_FLAG_SCATTER_FALLTHROUGH;
    if (INJ && (long long)c.draws > c.inj_len) c.flags |= SSE_FLAG_STREAM_EXHAUSTED;
    return len;
}

// ------------------------------------------------------------------------------------------------------
// worm_update (src/sse.jl:193-231) incl. worm_traverse! outer (src/sse.jl:233-260).  Needs indexed mode.
// ------------------------------------------------------------------------------------------------------
// worm_traverse! outer, start selection (src/sse.jl:241-251): picks the start leg (k0 = record index, l0 = leg) and the
// worm type w0.  Returns false if the injected stream ran out.
template <bool INJ>
__device__ __forceinline__ bool worm_pick_start(const DevModel &dm, Ctx &c, uint32_t &k0, uint32_t &l0, uint32_t &w0) {
    const uint32_t lane = c.lane;
    bool found = false;
    while (!found) {
        // rejection loop (sse.jl:241-247): 32 tries evaluated at once, the first success in order wins
        if (INJ && (long long)c.draws >= c.inj_len) { c.flags |= SSE_FLAG_STREAM_EXHAUSTED; return false; }
        const uint32_t p0 = (uint32_t)sse_uint_below(draw<INJ>(c, c.draws + 2u * lane), (uint64_t)c.M);
        const uint32_t ll = (uint32_t)sse_uint_below(draw<INJ>(c, c.draws + 2u * lane + 1u), 4u);
        const uint32_t v = __ldcg(c.ops + p0);
        const uint32_t ok = __ballot_sync(FULL, v != 0u);
        if (ok) {
            const int t = __ffs(ok) - 1;
            k0 = __shfl_sync(FULL, v, t) - 1u;
            l0 = __shfl_sync(FULL, ll, t);
            c.draws += 2u * (unsigned)(t + 1);
            found = true;
        } else {
            c.draws += 64u;
        }
    }
    const uint4 R0 = __ldcg(c.rec + 2u * k0 + 1u);  // {op code, hints}
    const uint4 bi = __ldg(dm.bond_info + op_bond(R0.x));
    const uint32_t dim0 = (l0 & 1u) ? (bi.y >> 24) : (bi.x >> 24);  // site_of_leg (sse.jl:250)
    w0 = 1u + (uint32_t)sse_uint_below(draw<INJ>(c, c.draws), dim0 - 1u);  // sse.jl:251
    c.draws += 1;
    return true;
}

// worm_update after the worms (src/sse.jl:200-228): WormLengthFraction, the worm-count controller, and the state
// rebuild from the first leg on each site.  total = 1 + sum of the worm lengths (sse.jl:194-198).
template <bool INJ>
__device__ __forceinline__ void worm_finish(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool thermalized,
                                            int widx, double total) {
    const uint32_t lane = c.lane, lt = lanemask_lt();
    if (thermalized && c.n != 0) {  // sse.jl:200-202
        c.last_wlf = total / (double)c.n;
        if (lane == 0) {
            dw.acc[(size_t)widx * dw.n_obs + SSE_OBS_WORM_LENGTH_FRACTION] += c.last_wlf;
            dw.acc_cnt[2 * widx + 1] += 1;
        }
    }
    const double avg_worm_length = total / ceil(c.num_worms);  // sse.jl:204
    if (!thermalized) {                                        // sse.jl:205-217
        c.avg_wl += dw.atten * (avg_worm_length - c.avg_wl);
        const double target_worms = dw.twlf * (double)c.n / c.avg_wl;
        c.num_worms += dw.atten * (target_worms - c.num_worms + 100.0 * sse_tanh(target_worms - c.num_worms));
        if (dw.atten != 0) {
            const double lo = 1.0, hi = 1.0 + (double)c.n / 2.0;
            c.num_worms = c.num_worms < lo ? lo : (c.num_worms > hi ? hi : c.num_worms);
        }
    }
    // rebuild the state from the first leg on each site; untouched sites are redrawn IN SITE ORDER (sse.jl:219-228)
    const int N = dm.n_sites;
    __syncwarp();
    for (int b = 0; b < N; b += 32) {
        const int s = b + (int)lane;
        const bool act = s < N;
        const uint32_t f = act ? c.vfirst[s] : 0u;
        const bool empty = act && f == NONE32;
        const uint32_t em = __ballot_sync(FULL, empty);
        if (empty) {
            const uint32_t d = dm.site_dim[s];
            c.state[s] = (uint8_t)(1u + (uint32_t)sse_uint_below(draw<INJ>(c, c.draws + __popc(em & lt)), d));
        } else if (act) {
            const uint32_t op = __ldcg(reinterpret_cast<const uint32_t *>(c.rec + 2u * (f >> 2) + 1u));
            c.state[s] = (uint8_t)((st.vinfo[op_gv(op)] >> (8u * (f & 3u))) & 0xffu);
        }
        c.draws += __popc(em);
    }
    if (INJ && (long long)c.draws > c.inj_len) c.flags |= SSE_FLAG_STREAM_EXHAUSTED;
    __syncwarp();
}

template <bool INJ>
__device__ void phase_worm_update(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool thermalized,
                                  int widx) {
    const int nworms = (int)ceil(c.num_worms);
    double total = 1.0;  // sse.jl:194
    for (int wi = 0; wi < nworms; ++wi) {
        if (c.n == 0) continue;  // worm_traverse! returns 0 without drawing (sse.jl:234-236)
        uint32_t k0 = 0, l0 = 0, w0 = 0;
        if (!worm_pick_start<INJ>(dm, c, k0, l0, w0)) return;
        const uint32_t len = worm_traverse<INJ>(st, dm, c, k0, l0, w0);
        total += (double)len;
        c.visits += len;
        if (c.flags & SSE_FLAG_STREAM_EXHAUSTED) return;
    }
    worm_finish<INJ>(st, dm, dw, c, thermalized, widx, total);
}

// ------------------------------------------------------------------------------------------------------
// K4C: commit the worm phase's vertices into the string (indexed -> committed) and, if asked, Carlo.measure!
// (src/sse.jl:70-87): measure_sign (:305-314), the scalar observables, measure_opstring! (:321-376) with the
// table-driven MagnetizationEstimator init/measure/result (magnetization_estimator.jl:96-230).
// out[n_obs] (global) receives the observables.
// ------------------------------------------------------------------------------------------------------
__device__ void phase_commit_measure(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool indexed,
                                     bool do_measure, double *out) {
    const uint32_t lane = c.lane;
    const int M = c.M;
    const int nchunks = (M + 31) >> 5;
    uint32_t neg = 0;
    if (indexed || do_measure) {
        for (int ch = 0; ch < nchunks; ++ch) {
            const int p = ch * 32 + (int)lane;
            uint32_t op = p < M ? c.ops[p] : 0u;
            if (indexed && op != 0u) {
                op = __ldcg(reinterpret_cast<const uint32_t *>(c.rec + 2u * (op - 1u) + 1u));
                c.ops[p] = op;
            }
            if (do_measure) neg += __popc(__ballot_sync(FULL, op != 0u && st.vneg[op_gv(op)]));
        }
        __syncwarp();
    }
    if (!do_measure) return;
    const double sign = (neg & 1u) ? -1.0 : 1.0;  // sse.jl:313
    const double nops = (double)c.n;
    if (lane == 0) {
        out[SSE_OBS_SIGN] = sign;
        out[SSE_OBS_OPERATOR_COUNT] = nops;
        out[SSE_OBS_SIGN_OPERATOR_COUNT] = sign * nops;
        out[SSE_OBS_SIGN_OPERATOR_COUNT2] = sign * (nops * nops);
        out[SSE_OBS_SIGN_ENERGY] = -sign * (nops * c.T + dm.energy_offset) / (double)dm.norm_sites;
        out[SSE_OBS_WORM_LENGTH_FRACTION] = c.last_wlf;
    }
    const int N = dm.n_sites, md = dm.est_max_dim;
    for (int e = 0; e < dm.n_est; ++e) {
        const double *ev = dm.est_values + (size_t)e * N * md;
        // init (magnetization_estimator.jl:96-123)
        double part = 0.0;
        for (int s = lane; s < N; s += 32) part += __ldg(ev + (size_t)s * md + (c.state[s] - 1));
        double tmpmag = warp_sum_f64(part);
        double mag = 0, absmag = 0, mag2 = 0, mag4 = 0;  // per-lane partial sums
        if (lane == 0) { mag = tmpmag; absmag = fabs(tmpmag); mag2 = tmpmag * tmpmag; mag4 = mag2 * mag2; }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int p = ch * 32 + (int)lane;
            const uint32_t op = p < M ? c.ops[p] : 0u;
            const bool nonid = op != 0u;
            double delta = 0.0;
            if (nonid && !(op & 2u)) {  // off-diagonal: tmpmag += sum_l sign*(m(top_l) - m(bottom_l)) (:134-150)
                const uint4 bi = __ldg(dm.bond_info + op_bond(op));
                const uint32_t vi = st.vinfo[op_gv(op)];
                const double *ea = ev + (size_t)(bi.x & NONE24) * md, *eb = ev + (size_t)(bi.y & NONE24) * md;
                delta = (__ldg(ea + ((vi >> 16) & 0xffu) - 1) - __ldg(ea + (vi & 0xffu) - 1)) +
                        (__ldg(eb + (vi >> 24) - 1) - __ldg(eb + ((vi >> 8) & 0xffu) - 1));
            }
            const uint32_t offm = __ballot_sync(FULL, delta != 0.0);
            double scan = delta;  // inclusive prefix sum over the chunk, in slot order
            if (offm) {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double up = shfl_up_f64(scan, d);
                    if ((int)lane >= d) scan += up;
                }
            }
            if (nonid) {  // every non-identity operator is one sample (:152-158)
                const double v = tmpmag + scan, v2 = v * v;
                mag += v;
                absmag += fabs(v);
                mag2 += v2;
                mag4 += v2 * v2;
            }
            if (offm) tmpmag += shfl_f64(scan, 31);
        }
        mag = warp_sum_f64(mag);
        absmag = warp_sum_f64(absmag);
        mag2 = warp_sum_f64(mag2);
        mag4 = warp_sum_f64(mag4);
        if (lane == 0) {  // result (:205-230)
            const double ns = 1.0 + nops;
            const double norm = 1.0 / (double)dm.norm_sites;
            mag *= norm;
            absmag *= norm;
            mag2 *= norm * norm;
            mag4 *= (norm * norm) * (norm * norm);
            double *o = out + SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * e;
            o[0] = sign * mag / ns;
            o[1] = sign * absmag / ns;
            o[2] = sign * mag2 / ns;
            o[3] = sign * mag4 / ns;
            o[4] = sign * (1.0 / c.T / (ns + 1.0) / ns * (mag * mag + mag2) * (double)dm.norm_sites);
        }
    }
    __syncwarp();
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
    return st;
}

// bytes of per-warp shared scratch: random draws + (level >= 1) state[N], mark[N] + (level 2) vlast[N]
__host__ __device__ inline int warp_scratch_bytes(int n_sites, int level) {
    int b = RNG_WORDS * 8;
    if (level >= 1) b += 2 * ((n_sites + 15) & ~15);
    if (level >= 2) b += 4 * ((n_sites + 3) & ~3);
    return (b + 15) & ~15;
}

// The one kernel: every mode shares the phase code above.  One warp = one walker.
template <bool INJ>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 7) k_walkers(const DevModel dm, const DevWalkers dw, const LaunchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const SmTab st = stage_tables(dm, smem);
    const int warp = threadIdx.x >> 5;
    const int w = blockIdx.x * WARPS_PER_CTA + warp;
    if (w >= dw.W) return;
    const int N = dm.n_sites;
    Ctx c;
    c.lane = threadIdx.x & 31;
    uint8_t *scratch = smem + dm.tl.bytes + (size_t)warp * warp_scratch_bytes(N, dw.smem_state);
    c.rng = reinterpret_cast<unsigned long long *>(scratch);
    uint8_t *gstate = dw.state + (size_t)w * N;
    if (dw.smem_state) {
        c.state = scratch + RNG_WORDS * 8;
        c.mark = c.state + ((N + 15) & ~15);
    } else {
        c.state = gstate;
        c.mark = dw.mark + (size_t)w * N;
    }
    c.ops = dw.ops + (size_t)w * dw.M_cap;
    c.rec = dw.rec + 2 * (size_t)w * dw.n_cap;
    c.vfirst = dw.vfirst + (size_t)w * N;
    uint32_t *gvlast = dw.vlast + (size_t)w * N;
    c.vlast = dw.smem_state >= 2 ? reinterpret_cast<uint32_t *>(c.mark + ((N + 15) & ~15)) : gvlast;
    c.inj = INJ ? dw.inj + (size_t)w * dw.inj_len : nullptr;
    c.inj_len = dw.inj_len;
    c.seed = dw.seed;
    c.wid = dw.wid_off + (unsigned long long)w;
    c.draws = dw.draws[w];
    c.T = dw.T[w];
    c.num_worms = dw.num_worms[w];
    c.avg_wl = dw.avg_wl[w];
    c.last_wlf = dw.last_wlf[w];
    c.M = dw.M[w];
    c.n = dw.n[w];
    c.flags = dw.flags[w];
    c.visits = 0;
    const uint32_t fatal = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;
    if (c.flags & fatal) return;
    for (int s = c.lane; s < N; s += 32) {
        if (dw.smem_state) c.state[s] = gstate[s];
        c.mark[s] = 0;
    }
    __syncwarp();
    double *out = dw.obs_out + (size_t)w * dw.n_obs;
    unsigned long long sweeps = 0, sum_n = 0, sum_M = 0, cyc[3] = {0, 0, 0};

    switch (a.mode) {
        case MODE_SWEEP:
            for (int s = 0; s < a.n_sweeps && !(c.flags & fatal); ++s) {  // Carlo.sweep! (sse.jl:62-68)
                const long long t0 = clock64();
                phase_diag_build<INJ>(st, dm, dw, c, true, true);
                if (c.flags & fatal) break;
                if (!(dm.variant & 4u)) phase_hints(dm, c);
                const long long t1 = clock64();
                phase_worm_update<INJ>(st, dm, dw, c, a.thermalized != 0, w);
                if (c.flags & fatal) break;
                const long long t2 = clock64();
                phase_commit_measure(st, dm, dw, c, true, a.measure != 0, out);
                const long long t3 = clock64();
                cyc[0] += (unsigned long long)(t1 - t0);
                cyc[1] += (unsigned long long)(t2 - t1);
                cyc[2] += (unsigned long long)(t3 - t2);
                ++sweeps;
                sum_n += (unsigned long long)c.n;
                sum_M += (unsigned long long)c.M;
                if (a.measure) {
                    __syncwarp();
                    for (int i = c.lane; i < dw.n_obs; i += 32)
                        if (i != SSE_OBS_WORM_LENGTH_FRACTION) dw.acc[(size_t)w * dw.n_obs + i] += out[i];
                    if (c.lane == 0) dw.acc_cnt[2 * w] += 1;
                    __syncwarp();
                }
            }
            break;
        case MODE_INIT: {  // Carlo.init! (sse.jl:47-60): M and the zeroed string are set by the host
            for (int s = c.lane; s < N; s += 32)
                c.state[s] = (uint8_t)(1u + (uint32_t)sse_uint_below(draw<INJ>(c, c.draws + s), dm.site_dim[s]));
            c.draws += N;
            __syncwarp();
            for (int i = 0; i < a.warmup && !(c.flags & fatal); ++i) phase_diag_build<INJ>(st, dm, dw, c, true, false);
            break;
        }
        case MODE_DIAG:
            phase_diag_build<INJ>(st, dm, dw, c, true, false);
            break;
        case MODE_MAKE_VL:
            phase_diag_build<INJ>(st, dm, dw, c, false, true);
            break;
        case MODE_WORM_UPDATE:
            phase_worm_update<INJ>(st, dm, dw, c, a.thermalized != 0, w);
            break;
        case MODE_WORM_TRAVERSE: {
            const uint32_t v = c.ops[a.p0];
            long long len = -1;
            if (v != 0u) len = (long long)worm_traverse<INJ>(st, dm, c, v - 1u, (uint32_t)a.l0, (uint32_t)a.w0);
            if (c.lane == 0) dw.dbg_len[w] = len;
            break;
        }
        case MODE_COMMIT:
            phase_commit_measure(st, dm, dw, c, true, false, out);
            break;
        case MODE_MEASURE:
            phase_commit_measure(st, dm, dw, c, a.indexed != 0, true, out);
            break;
    }
    if (INJ && (long long)c.draws > c.inj_len) c.flags |= SSE_FLAG_STREAM_EXHAUSTED;
    __syncwarp();
    if (dw.smem_state)
        for (int s = c.lane; s < N; s += 32) gstate[s] = c.state[s];
    if (dw.smem_state >= 2 && (a.mode == MODE_MAKE_VL))
        for (int s = c.lane; s < N; s += 32) gvlast[s] = c.vlast[s];  // read back by sse_dbg_get_vertex_list
    if (c.lane == 0) {
        dw.draws[w] = c.draws;
        dw.num_worms[w] = c.num_worms;
        dw.avg_wl[w] = c.avg_wl;
        dw.last_wlf[w] = c.last_wlf;
        dw.M[w] = c.M;
        dw.n[w] = c.n;
        dw.flags[w] = c.flags;
        if (c.visits) atomicAdd(dw.counters + 0, c.visits);
        if (sweeps) {
            atomicAdd(dw.counters + 1, sweeps);
            atomicAdd(dw.counters + 2, sum_n);
            atomicAdd(dw.counters + 3, sum_M);
            atomicAdd(dw.counters + 4, cyc[0]);  // SM cycles spent per phase, summed over walkers
            atomicAdd(dw.counters + 5, cyc[1]);
            atomicAdd(dw.counters + 6, cyc[2]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// beta doubling (thermalisation aid; NOT part of the reference): for a periodic configuration (state, S_M) the
// doubled string S_M S_M with the same state is a valid configuration at inverse temperature 2*beta with 2n
// operators, so a cold walker can be grown from a cheap hot one in log2(beta) steps instead of thousands of
// full-size sweeps.  Needs committed mode.  One warp per walker; M, n and the controller's average worm length double,
// T halves.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_double_beta(const DevWalkers dw) {
    const int w = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= dw.W) return;
    const uint32_t fatal = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;
    const uint32_t flags = dw.flags[w];
    const long long M = dw.M[w], n = dw.n[w];
    const double T = dw.T[w], awl = dw.avg_wl[w];
    __syncwarp();
    if (flags & fatal) return;
    if (2 * M > dw.M_cap || 2 * n > dw.n_cap) {
        if (lane == 0) dw.flags[w] = flags | (2 * M > dw.M_cap ? SSE_FLAG_M_OVERFLOW : SSE_FLAG_N_OVERFLOW);
        return;
    }
    uint32_t *ops = dw.ops + (size_t)w * dw.M_cap;
    for (long long p = lane; p < M; p += 32) ops[M + p] = ops[p];
    if (lane == 0) {
        dw.M[w] = (int)(2 * M);
        dw.n[w] = (int)(2 * n);
        dw.T[w] = T * 0.5;
        // worm-count controller (sse.jl:204-217): worms get at least twice as long at twice the inverse temperature;
        // carrying the old average over would keep num_worms (target = twlf * n / avg_wl) far too high for many sweeps
        dw.avg_wl[w] = awl * 2.0;
    }
}

}  // namespace sse
