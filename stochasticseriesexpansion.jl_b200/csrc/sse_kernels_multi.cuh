// sse_kernels_multi.cuh — CH walkers per warp (CH = 2 or 4): the sweep kernel for batches larger than one walker
// per resident warp.  Opt-in (sse_walkers_opts is unchanged: environment SSE_B200_CHAINS=2|4 at sse_walkers_create);
// the default remains sse::k_walkers, one walker per warp.
//
// Why: the worm phase is one dependent load chain per walker; with one chain per warp a B200 holds
// 28 warps x 148 SMs = 4144 chains and runs at the latency floor of that many chains.  The microbenchmark
// (profiles/r1_chase_floor.txt) shows the memory system sustains 1.7x more hops/s at 8192 chains.  Registers, not
// memory, cap the warps per SM, so the extra chains are run as instruction-level parallelism inside each warp:
// every lane still executes every chain uniformly, the chains' record loads are in flight together.
//
// The streaming phases (diagonal update + records, hints, commit + estimators) are the single-walker phase functions
// of sse_kernels.cuh, executed by the whole warp for one walker at a time; only the worm phase is interleaved.  The
// warp's walkers are not kept in step with each other: each one moves through its own sweeps (multi_run), so the
// chase always runs over every walker that is in the middle of a worm.
// Results are bit-identical to the one-walker-per-warp kernel (walkers are independent, and each walker's draws
// come from its own stream position).
#pragma once
#include "sse_kernels.cuh"

namespace sse {

// Per-warp shared scratch of the multi-chain kernel: per chain {random draws, state[N] (level >= 1)}, then ONE
// mark[N] (level >= 1) and ONE vlast[N] (level 2) shared by the warp's walkers (only used inside phase_diag_build,
// which runs for one walker at a time).
__host__ __device__ inline int multi_chain_bytes(int n_sites, int level) {
    return RNG_WORDS * 8 + (level >= 1 ? ((n_sites + 15) & ~15) : 0);
}
__host__ __device__ inline int multi_warp_bytes(int n_sites, int level, int ch) {
    int b = ch * multi_chain_bytes(n_sites, level);
    if (level >= 1) b += (n_sites + 15) & ~15;
    if (level >= 2) b += 4 * ((n_sites + 3) & ~3);
    return (b + 15) & ~15;
}

// uniform doubles for the draws [2*j0, 2*j0 + 64) of one walker -> rbuf[64] (cf. fill_u01)
template <bool INJ>
__device__ __forceinline__ void fill_u01_chain(uint32_t rbuf_s, uint32_t lane, unsigned long long seed, unsigned long long wid,
                                               const unsigned long long *inj, long long inj_len, unsigned long long j0) {
    uint64_t x0, x1;
    if (INJ) {
        const unsigned long long k = 2ull * (j0 + lane);
        x0 = (long long)k < inj_len ? (uint64_t)__ldg(inj + k) : 0ull;
        x1 = (long long)(k + 1) < inj_len ? (uint64_t)__ldg(inj + k + 1) : 0ull;
    } else {
        uint32_t b[4];
        sse_philox_block(seed, wid, j0 + lane, b);
        x0 = (uint64_t)b[0] | ((uint64_t)b[1] << 32);
        x1 = (uint64_t)b[2] | ((uint64_t)b[3] << 32);
    }
    sts_f64x2(rbuf_s + 16u * lane, sse_u01(x0), sse_u01(x1));
}

// State of the interleaved chase, kept in memory between calls of worm_multi_loop (it is re-entered once per worm).
template <int CH>
struct MultiArgs {
    uint32_t t1_s, outc_s, maxw, lane, variant;
    unsigned long long seed;
    long long inj_len;
    uint32_t act;      // bit c: chain c is in the middle of a worm
    uint32_t closed1;  // out: chains whose worm closed through the first stop test (sse.jl:288-290), which does not count the visit
    uint4 *rec[CH];
    const unsigned long long *inj[CH];
    unsigned long long wid[CH], j0[CH];
    uint32_t rbuf_s[CH], ri[CH];
    uint4 R[CH], H[CH];  // the chain's current record (links, {op code, hints}): already loaded
    uint32_t pos[CH], wf[CH], patch[CH], pval[CH], pos0[CH], w0[CH], fell[CH];
    unsigned long long draws0[CH];  // stream position at the start of the chase: one draw per visit, so the worm length
                                    // returned by worm_traverse! is 1 + (draws at the end - draws0) - (closed by stop test 1)
};

// The interleaved worm_traverse! inner loops (src/sse.jl:274-300) of up to CH walkers.  One iteration = one visit of
// every active chain; a chain's next record is requested as soon as its exit leg is known and is consumed one
// iteration later, after the other chains' visits, so up to CH record loads are in flight per warp.  Unlike the
// single-chain loop the record registers are reused in place (everything needed from the old record, including the
// prefetch hint, is taken before the new loads are issued), so no ping-pong copies exist.  Returns the mask of
// chains whose worm closed; everything else is written back to `a`.
// W1 = every site has dimension 2 (max_worm == 1, e.g. all S = 1/2 models): there is one worm type, so the worm index is not
// tracked, the transition index needs no worm term and the worm-type halves of both stop tests are constant.
template <bool INJ, int CH, bool W1>
__device__ __noinline__ uint32_t worm_multi_loop(MultiArgs<CH> &a) {
    const uint32_t t1_s = a.t1_s, outc_s = a.outc_s, maxw4 = W1 ? 4u : a.maxw * 4u;
    const bool pref = !(a.variant & 2u);
    uint4 *rec[CH];
    uint4 R[CH], H[CH];
    uint32_t pos[CH], wf[CH], patch[CH], pval[CH], pos0[CH], w0[CH], ri[CH], rbuf_s[CH], fell = 0, closed1 = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        rec[c] = a.rec[c];
        R[c] = a.R[c];
        H[c] = a.H[c];
        pos[c] = a.pos[c];
        wf[c] = W1 ? 1u : a.wf[c];
        patch[c] = a.patch[c];
        pval[c] = a.pval[c];
        pos0[c] = a.pos0[c];
        w0[c] = W1 ? 1u : a.w0[c];
        ri[c] = a.ri[c];
        rbuf_s[c] = a.rbuf_s[c];
    }
    const uint32_t act = a.act;
    uint32_t closed = 0;
    while (!closed) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (!((act >> c) & 1u)) continue;
            if (ri[c] >= 64u) {  // next 64 draws of this walker
                a.j0[c] += 32;
                ri[c] = 0;
                __syncwarp();
                fill_u01_chain<INJ>(rbuf_s[c], a.lane, a.seed, a.wid[c], a.inj[c], a.inj_len, a.j0[c]);
                __syncwarp();
            }
            const double r = lds_f64(rbuf_s[c] + 8u * ri[c]++);  // rand(rng) (sse.jl:282)
            const uint32_t p = pos[c];
            const uint32_t x = patch[c] ? pval[c] : H[c].x;
            // transitions[leg_in, worm_in, vi] fused with its first outcome (vertex_data.jl:115-123)
            uint4 e = lds128(t1_s + 16u * (W1 ? ((x & VMASK) | (p & 3u)) : op_gv(x) * maxw4 + (((wf[c] - 1u) << 2) | (p & 3u))));
            if (!(r < __hiloint2double((int)e.y, (int)e.x))) {
                const uint32_t off = (e.w >> 6) & 0x3ffffu, cnt = e.w & 63u;
                bool hit = false;
                for (uint32_t j = 0; !hit && j < cnt; ++j) {
                    e = lds128(outc_s + 16u * (off + j));
                    hit = r < __hiloint2double((int)e.y, (int)e.x);
                }
                if (!hit) fell |= 1u << c;  // vertex_data.jl:124; clamped to the last outcome
            }
            const uint32_t leg_out = (e.z >> 16) & 3u;
            const uint32_t posn = rec_sel(R[c], leg_out);  // (leg_in, p) = vertices[leg_out, p] (sse.jl:295)
            const uint32_t hint = rec_link(H[c], leg_out);
            // record k lives at byte offset 32 k = (link & ~3) << 3, which fits 32 bits (n_capacity <= 2^22): one
            // 32-bit shift + one 64-bit add per address instead of a 64-bit shift
            uint8_t *const base = reinterpret_cast<uint8_t *>(rec[c]);
            const uint4 *const rn = reinterpret_cast<const uint4 *>(base + ((posn & ~3u) << 3));
            R[c] = ldg_cg128(rn);
            H[c] = ldg_cg128(rn + 1);
            // ---- everything below overlaps with the loads (and with the other chains' visits) ----
            const uint32_t newop = (x & ~(VMASK | 2u)) | (e.z & (VMASK | 2u));  // OperCode(bond, new_vertex) (sse.jl:285)
            stg_u32(base + ((p & ~3u) << 3) + 16u, newop);
            if (pref) prefetch_l2(base + ((hint & ~3u) << 3));
            const uint32_t w_out = W1 ? 1u : e.z >> 24, dim_out = W1 ? 2u : e.w >> 24;
            const bool stop1 = (((p & ~3u) | leg_out) == pos0[c]) && (w_out + w0[c] == dim_out);  // sse.jl:288-290
            if (!W1) wf[c] = w_out;
            patch[c] = ((posn >> 2) == (p >> 2)) ? 1u : 0u;  // the link re-enters this record: its load preceded the store
            pval[c] = newop;
            pos[c] = posn;
            const bool stop2 = (posn == pos0[c]) && (w_out == w0[c]);  // sse.jl:297-299
            if (stop1) closed1 |= 1u << c;
            if (stop1 || stop2) closed |= 1u << c;
        }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        a.R[c] = R[c];
        a.H[c] = H[c];
        a.pos[c] = pos[c];
        a.wf[c] = wf[c];
        a.patch[c] = patch[c];
        a.pval[c] = pval[c];
        a.ri[c] = ri[c];
        if ((fell >> c) & 1u) a.fell[c] = 1;
    }
    a.closed1 = closed1;
    return closed;
}

// Everything a warp needs to (re)build the Ctx of one of its walkers.
struct MultiWarp {
    int w0;             // index of the warp's first walker
    int level;          // dw.smem_state for this launch
    uint32_t lane;
    uint8_t *scratch;   // the warp's shared scratch (multi_warp_bytes)
    uint32_t live;      // bit c: walker w0 + c exists and carries no fatal flag
};

__device__ __forceinline__ uint8_t *multi_chain_scratch(const DevModel &dm, const MultiWarp &mw, int c) {
    return mw.scratch + (size_t)c * multi_chain_bytes(dm.n_sites, mw.level);
}

// Ctx of chain c from the walker arrays (the scalars live in global memory between phases)
template <bool INJ, int CH>
__device__ __forceinline__ Ctx multi_open(const DevModel &dm, const DevWalkers &dw, const MultiWarp &mw, int ci) {
    const int N = dm.n_sites, w = mw.w0 + ci;
    Ctx c;
    c.lane = mw.lane;
    uint8_t *cs = multi_chain_scratch(dm, mw, ci);
    uint8_t *shared = mw.scratch + (size_t)CH * multi_chain_bytes(N, mw.level);
    c.rng = reinterpret_cast<unsigned long long *>(cs);
    if (mw.level) {
        c.state = cs + RNG_WORDS * 8;
        c.mark = shared;
    } else {
        c.state = dw.state + (size_t)w * N;
        c.mark = dw.mark + (size_t)w * N;
    }
    c.vlast = mw.level >= 2 ? reinterpret_cast<uint32_t *>(shared + ((N + 15) & ~15)) : dw.vlast + (size_t)w * N;
    c.ops = dw.ops + (size_t)w * dw.M_cap;
    c.rec = dw.rec + 2 * (size_t)w * dw.n_cap;
    c.vfirst = dw.vfirst + (size_t)w * N;
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
    return c;
}

__device__ __forceinline__ void multi_close(const DevWalkers &dw, MultiWarp &mw, int ci, const Ctx &c) {
    const int w = mw.w0 + ci;
    const uint32_t fatal = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;
    __syncwarp();  // every lane has read the scalars it needs before lane 0 overwrites them
    if (c.lane == 0) {
        dw.draws[w] = c.draws;
        dw.num_worms[w] = c.num_worms;
        dw.avg_wl[w] = c.avg_wl;
        dw.last_wlf[w] = c.last_wlf;
        dw.M[w] = c.M;
        dw.n[w] = c.n;
        dw.flags[w] = c.flags;
    }
    __syncwarp();
    if (c.flags & fatal) mw.live &= ~(1u << ci);
}

// Totals a warp reports at the end of a launch.
struct MultiStats {
    unsigned long long visits, sweeps, sum_n, sum_M, cyc_stream, cyc_commit;
};

// n_sweeps x Carlo.sweep! (src/sse.jl:62-68) for the warp's live walkers.  The walkers are NOT kept in step: each one runs
// through   diagonal update + records + hints  ->  its worms  ->  worm_finish + commit (+ estimators)   at its own pace.
// Whenever a walker's worm closes, cold code here decides what it does next (its next worm, or the end of its sweep and
// the streaming phases of its next one, executed by the whole warp) and then the interleaved chase resumes with every
// walker that is in the middle of a worm.  So a warp always chases as many chains as it has unfinished walkers, and the
// spread of worm work per sweep costs nothing until the very end of the launch.
template <bool INJ, int CH>
__device__ __noinline__ void multi_run(const SmTab &st, const DevModel &dm, const DevWalkers &dw, MultiWarp &mw, int n_sweeps,
                                       bool thermalized, bool measure, MultiStats &ms) {
    const uint32_t fatal = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;
    MultiArgs<CH> a;
    a.t1_s = (uint32_t)__cvta_generic_to_shared(st.t1);
    a.outc_s = (uint32_t)__cvta_generic_to_shared(st.outc);
    a.maxw = (uint32_t)dm.max_worm;
    a.lane = mw.lane;
    a.variant = dm.variant;
    a.seed = dw.seed;
    a.inj_len = dw.inj_len;
    a.act = 0;
    a.closed1 = 0;
    int nworms[CH] = {}, wi[CH] = {}, sweeps_left[CH] = {};
    double total[CH] = {};

    // diagonal_update + make_vertex_list! (+ hints) of walker ci's next sweep; false if the walker hit a fatal flag
    auto begin_sweep = [&](int ci) -> bool {
        const long long t0 = clock64();
        Ctx c = multi_open<INJ, CH>(dm, dw, mw, ci);
        phase_diag_build<INJ>(st, dm, dw, c, true, true);
        if (!(c.flags & fatal) && !(dm.variant & 4u)) phase_hints(dm, c);
        nworms[ci] = (int)ceil(c.num_worms);
        wi[ci] = 0;
        total[ci] = 1.0;  // sse.jl:194
        multi_close(dw, mw, ci, c);
        ms.cyc_stream += (unsigned long long)(clock64() - t0);
        return (mw.live >> ci) & 1u;
    };
    // start the next worm of walker ci (sets its bit in a.act); false if it has none left or its stream ran out
    auto next_worm = [&](int ci) -> bool {
        Ctx c = multi_open<INJ, CH>(dm, dw, mw, ci);
        bool started = false;
        while (!started && wi[ci] < nworms[ci]) {
            ++wi[ci];
            if (c.n == 0) continue;  // worm_traverse! returns 0 without drawing (sse.jl:234-236)
            uint32_t k0 = 0, l0 = 0, w0 = 0;
            if (!worm_pick_start<INJ>(dm, c, k0, l0, w0)) break;  // stream exhausted: flag set, the walker retires
            a.rec[ci] = c.rec;
            a.inj[ci] = c.inj;
            a.wid[ci] = c.wid;
            a.rbuf_s[ci] = (uint32_t)__cvta_generic_to_shared(c.rng);
            a.pos0[ci] = (k0 << 2) | l0;
            a.w0[ci] = w0;
            a.pos[ci] = a.pos0[ci];
            a.wf[ci] = w0;
            a.draws0[ci] = c.draws;
            a.patch[ci] = 0;
            a.pval[ci] = 0;
            a.fell[ci] = 0;
            a.j0[ci] = c.draws >> 1;
            a.ri[ci] = (uint32_t)(c.draws & 1ull);
            a.R[ci] = ldg_cg128(c.rec + 2u * k0);
            a.H[ci] = ldg_cg128(c.rec + 2u * k0 + 1u);
            __syncwarp();
            fill_u01_chain<INJ>(a.rbuf_s[ci], a.lane, a.seed, a.wid[ci], a.inj[ci], a.inj_len, a.j0[ci]);
            __syncwarp();
            started = true;
        }
        if (started) a.act |= 1u << ci;
        else a.act &= ~(1u << ci);
        multi_close(dw, mw, ci, c);
        return started;
    };
    // the rest of worm_update (sse.jl:200-228), then commit (+ Carlo.measure!) of walker ci's finished sweep
    auto end_sweep = [&](int ci) {
        const int w = mw.w0 + ci;
        Ctx c = multi_open<INJ, CH>(dm, dw, mw, ci);
        worm_finish<INJ>(st, dm, dw, c, thermalized, w, total[ci]);
        const long long t0 = clock64();
        if (!(c.flags & fatal)) {
            double *out = dw.obs_out + (size_t)w * dw.n_obs;
            phase_commit_measure(st, dm, dw, c, true, measure, out);
            ++ms.sweeps;
            ms.sum_n += (unsigned long long)c.n;
            ms.sum_M += (unsigned long long)c.M;
            if (measure) {
                __syncwarp();
                for (int i = c.lane; i < dw.n_obs; i += 32)
                    if (i != SSE_OBS_WORM_LENGTH_FRACTION) dw.acc[(size_t)w * dw.n_obs + i] += out[i];
                if (c.lane == 0) dw.acc_cnt[2 * w] += 1;
                __syncwarp();
            }
        }
        multi_close(dw, mw, ci, c);
        ms.cyc_commit += (unsigned long long)(clock64() - t0);
    };
    // drive walker ci until it is in the middle of a worm again, has done all its sweeps, or carries a fatal flag
    auto advance = [&](int ci) {
        while ((mw.live >> ci) & 1u) {
            if (next_worm(ci)) return;
            if (!((mw.live >> ci) & 1u)) return;  // stream exhausted while picking a start
            end_sweep(ci);
            if (!((mw.live >> ci) & 1u)) return;
            if (--sweeps_left[ci] <= 0) return;
            if (!begin_sweep(ci)) return;
        }
    };

    for (int ci = 0; ci < CH; ++ci) {
        if (!((mw.live >> ci) & 1u)) continue;
        sweeps_left[ci] = n_sweeps;
        if (begin_sweep(ci)) advance(ci);
    }
    while (a.act) {
        const uint32_t closed = dm.max_worm == 1 ? worm_multi_loop<INJ, CH, true>(a) : worm_multi_loop<INJ, CH, false>(a);
        for (int ci = 0; ci < CH; ++ci) {
            if (!((closed >> ci) & 1u)) continue;
            const int w = mw.w0 + ci;
            // the walker's stream position after the worm; flags
            const unsigned long long draws = 2ull * a.j0[ci] + a.ri[ci];
            // worm_traverse!'s return value (sse.jl:264,291,302): 1 + visits, the closing visit of stop test 1 not counted
            const uint32_t len = 1u + (uint32_t)(draws - a.draws0[ci]) - ((a.closed1 >> ci) & 1u);
            uint32_t fl = dw.flags[w];
            if (a.fell[ci]) fl |= SSE_FLAG_SCATTER_FALLTHROUGH;
            if (INJ && (long long)draws > dw.inj_len) fl |= SSE_FLAG_STREAM_EXHAUSTED;
            __syncwarp();
            if (mw.lane == 0) {
                dw.draws[w] = draws;
                dw.flags[w] = fl;
            }
            __syncwarp();
            total[ci] += (double)len;
            ms.visits += len;
            a.act &= ~(1u << ci);
            if (fl & SSE_FLAG_STREAM_EXHAUSTED) {
                mw.live &= ~(1u << ci);
                continue;
            }
            advance(ci);
        }
    }
}

// Carlo.sweep! x n_sweeps for CH walkers per warp (MODE_SWEEP only; every other mode runs on sse::k_walkers).
template <bool INJ, int CH, int MINB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB) k_walkers_multi(const DevModel dm, const DevWalkers dw, const LaunchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const SmTab st = stage_tables(dm, smem);
    const int warp = threadIdx.x >> 5;
    const int N = dm.n_sites;
    MultiWarp mw;
    mw.w0 = (blockIdx.x * WARPS_PER_CTA + warp) * CH;
    if (mw.w0 >= dw.W) return;
    mw.level = dw.smem_state;
    mw.lane = threadIdx.x & 31;
    mw.scratch = smem + dm.tl.bytes + (size_t)warp * multi_warp_bytes(N, mw.level, CH);
    mw.live = 0;
    const uint32_t fatal = SSE_FLAG_M_OVERFLOW | SSE_FLAG_N_OVERFLOW | SSE_FLAG_STREAM_EXHAUSTED;
    for (int ci = 0; ci < CH; ++ci) {
        const int w = mw.w0 + ci;
        if (w >= dw.W || (dw.flags[w] & fatal)) continue;
        mw.live |= 1u << ci;
        if (mw.level) {
            uint8_t *s = multi_chain_scratch(dm, mw, ci) + RNG_WORDS * 8;
            const uint8_t *g = dw.state + (size_t)w * N;
            for (int i = mw.lane; i < N; i += 32) s[i] = g[i];
        } else {
            uint8_t *m = dw.mark + (size_t)w * N;
            for (int i = mw.lane; i < N; i += 32) m[i] = 0;
        }
    }
    if (mw.level) {
        uint8_t *m = mw.scratch + (size_t)CH * multi_chain_bytes(N, mw.level);
        for (int i = mw.lane; i < N; i += 32) m[i] = 0;
    }
    __syncwarp();
    const uint32_t loaded = mw.live;  // walkers whose state[] is held in shared memory during this launch
    MultiStats ms = {0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    multi_run<INJ, CH>(st, dm, dw, mw, a.n_sweeps, a.thermalized != 0, a.measure != 0, ms);
    const unsigned long long cyc_total = (unsigned long long)(clock64() - t_begin);
    __syncwarp();
    if (mw.level)
        for (int ci = 0; ci < CH; ++ci) {
            if (!((loaded >> ci) & 1u)) continue;
            const uint8_t *s = multi_chain_scratch(dm, mw, ci) + RNG_WORDS * 8;
            uint8_t *g = dw.state + (size_t)(mw.w0 + ci) * N;
            for (int i = mw.lane; i < N; i += 32) g[i] = s[i];
        }
    if (mw.lane == 0) {
        if (ms.visits) atomicAdd(dw.counters + 0, ms.visits);
        if (ms.sweeps) {
            atomicAdd(dw.counters + 1, ms.sweeps);
            atomicAdd(dw.counters + 2, ms.sum_n);
            atomicAdd(dw.counters + 3, ms.sum_M);
            // SM cycles per phase, per warp (= CH walkers): the worm share is what is left of the launch
            atomicAdd(dw.counters + 4, ms.cyc_stream);
            atomicAdd(dw.counters + 5, cyc_total - ms.cyc_stream - ms.cyc_commit);
            atomicAdd(dw.counters + 6, ms.cyc_commit);
        }
    }
}

}  // namespace sse
