// sse_sweep.cuh — the kernels.
//   sse::k_sweep        Carlo.sweep! (src/sse.jl:62-68) for every walker: persistent, one CTA per SM, worm warps (one lane =
//                       one walker) + stream warps (one warp = one walker), see sse_common.cuh.
//   sse::k_phase        single phases for Carlo.init!, Carlo.measure! and the parity hooks (one warp = one walker).
//   sse::k_double_beta  thermalisation aid (no reference counterpart).
#pragma once
#include "sse_stream.cuh"
#include "sse_worm.cuh"

namespace sse {

// ---- walker status inside a k_sweep launch (shared memory, one word per walker of the CTA) ----
enum : uint32_t {
    WS_NEED_STREAM = 0,  // waits for a stream warp (end of a sweep and/or start of the next one)
    WS_STREAMING = 1,    // a stream warp works on it
    WS_READY_WORM = 2,   // records built (or parked in its worm phase): waits for its worm lane
    WS_IN_WORM = 3,      // its worm lane chases
    WS_DONE = 4          // finished for this launch (sweep quota reached, visit budget used up, or fatal flag)
};

struct SweepArgs {
    int n_sweeps;                // sweeps per walker of this sse_sweep call (applied when reset != 0)
    unsigned long long budget;   // worm visits per walker and launch (~0 = unlimited)
    int reset;                   // 1 = first launch of a call: set the per-walker quota
    int thermalized, measure;
    int worm_warps, stream_warps, level, nloc_max;
};

struct SweepStats {
    unsigned long long visits, sweeps, sum_n, sum_M, cyc_build, cyc_finish, cyc_idle, tasks;
};

// The scheduler block of a CTA in shared memory: {n_done, q_head, q_tail, pad, status[nloc_max], queue[nloc_max]}.
// queue[] is a FIFO of the walkers that wait for a stream warp (entry = local index + 1, 0 = slot not written yet):
// first come, first served, so no walker falls behind the others and a launch that gives every walker the same visit
// budget ends for all of them at about the same time.
__host__ __device__ inline int sched_bytes(int nloc_max) { return (16 + 8 * nloc_max + 15) & ~15; }

// a worm lane (or the launch prologue) queues walker j for streaming
__device__ __forceinline__ void stream_enqueue(uint32_t *sched, int nloc_max, int j) {
    const uint32_t t = atomicAdd(sched + 2, 1u);
    st_volatile_shared(sched + 4 + nloc_max + (int)(t % (uint32_t)nloc_max), (uint32_t)j + 1u);
}

// One streaming task: whatever walker w needs until it is ready for its worm lane again (or done for this launch).
// Returns the walker's next status (WS_READY_WORM or WS_DONE).
template <bool INJ>
__device__ uint32_t stream_task(const SmTab &st, const DevModel &dm, const DevWalkers &dw, const SweepArgs &a, int w,
                                uint8_t *scratch, uint32_t lane, SweepStats &ss) {
    Ctx c = ctx_open(dm, dw, w, scratch, a.level, lane);
    WalkerCtl *ctl = dw.ctl + w;
    uint32_t phase = __ldcg(&ctl->phase);
    int sweeps_left = __ldcg(&ctl->sweeps_left);
    unsigned long long budget_left = __ldcg(&ctl->budget_left), sweep_visits = __ldcg(&ctl->sweep_visits);
    unsigned long long sweeps_done = __ldcg(&ctl->sweeps_done);
    uint32_t worms_left = 0, next = WS_DONE;
    __syncwarp();
    while (true) {
        bool fuse_measure = false;  // Carlo.measure! of the sweep that ends here rides on the next sweep's diagonal update
        if (phase == 1) {  // the worms of the sweep in flight are done: rest of worm_update, then Carlo.measure!
            const long long t0 = clock64();
            worm_finish<INJ>(st, dm, dw, c, a.thermalized != 0, w, 1.0 + (double)sweep_visits);  // sse.jl:194
            if (!(c.flags & FATAL_FLAGS)) {
                if (a.measure) {
                    fuse_measure = st.estrows != nullptr && sweeps_left > 1 && budget_left != 0;  // a build follows below
                    if (!fuse_measure) {
                        double *out = dw.obs_out + (size_t)w * dw.n_obs;
                        phase_measure(st, dm, dw, c, out);
                        accumulate_obs(dw, w, lane, out);
                    }
                }
                ++ss.sweeps;
                ss.sum_n += (unsigned long long)c.n;
                ss.sum_M += (unsigned long long)c.M;
                ++sweeps_done;
                --sweeps_left;
            }
            phase = 0;
            ss.cyc_finish += (unsigned long long)(clock64() - t0);
        }
        if ((c.flags & FATAL_FLAGS) || sweeps_left <= 0 || budget_left == 0) break;
        const long long t0 = clock64();
        if (fuse_measure) {
            double *out = dw.obs_out + (size_t)w * dw.n_obs;
            phase_diag_build<INJ, true>(st, dm, dw, c, true, out);  // sse.jl:63-64 + the measurement of the sweep before
            if (!(c.flags & FATAL_FLAGS)) accumulate_obs(dw, w, lane, out);
        } else {
            phase_diag_build<INJ>(st, dm, dw, c, true);  // sse.jl:63-64
        }
        ss.cyc_build += (unsigned long long)(clock64() - t0);
        if (c.flags & FATAL_FLAGS) break;
        phase = 1;
        sweep_visits = 0;
        worms_left = (uint32_t)(int)ceil(c.num_worms);  // sse.jl:196
        if (c.n == 0 || worms_left == 0) {  // worm_traverse! returns 0 without drawing (sse.jl:234-236): nothing to chase
            if (budget_left != ~0ull) --budget_left;  // an empty sweep costs one unit, so a budgeted launch always ends
            continue;
        }
        next = WS_READY_WORM;
        break;
    }
    ++ss.tasks;
    ctx_close(dm, dw, w, a.level, c);
    if (lane == 0) {
        ctl->phase = phase;
        ctl->sweeps_left = sweeps_left;
        ctl->budget_left = budget_left;
        ctl->sweep_visits = sweep_visits;
        ctl->sweeps_done = sweeps_done;
        ctl->worms_left = worms_left;
        ctl->inworm = 0;
    }
    __syncwarp();
    return next;
}

// Role of a worm warp inside k_sweep: one lane = one walker (walkers j = lane index + m * lanes of the CTA's status table).
template <bool INJ>
__device__ __forceinline__ void worm_warp_role(const SmTab &st, const DevModel &dm, const DevWalkers &dw, const SweepArgs &a, uint32_t *sched,
                                            int nloc, int warp, uint32_t lane) {
    uint32_t *n_done = sched, *status = sched + 4;
    // ------------------------------ worm warp: one lane = one walker ------------------------------
    const LaneEnv env = lane_env(st, dm, dw);
    const int nlanes = a.worm_warps * 32, me = warp * 32 + (int)lane;
    int cur = -1;
    bool finished = me >= nloc;
    WormLane L;
    unsigned long long visits = 0, it_active = 0, it_total = 0;
    const long long t_begin = clock64();
    while (true) {
        if (cur < 0 && !finished) {
            bool all_done = true;
            for (int j = me; j < nloc; j += nlanes) {
                const uint32_t s = ld_volatile_shared(status + j);
                if (s == WS_READY_WORM) { cur = j; break; }
                if (s != WS_DONE) all_done = false;
            }
            if (cur >= 0) {
                __threadfence();  // the stream warp's writes (records, words, control block) are visible
                st_volatile_shared(status + cur, WS_IN_WORM);
                const int w = (int)blockIdx.x + cur * (int)gridDim.x;
                lane_open(dw, w, L);
                bool ok = true;
                if (__ldcg(&dw.ctl[w].inworm)) lane_resume(env, L);
                else ok = lane_pick_start<INJ>(env, L);
                if (!ok) {  // injected stream exhausted
                    lane_store(dw, w, L, 0, SSE_FLAG_STREAM_EXHAUSTED);
                    __threadfence();
                    st_volatile_shared(status + cur, WS_DONE);
                    atomicAdd(n_done, 1u);
                    cur = -1;
                }
            } else if (all_done) {
                finished = true;
            }
        }
        if (cur >= 0) {
            ++it_active;
            const int w = (int)blockIdx.x + cur * (int)gridDim.x;
            const bool closed = lane_visit<INJ>(env, L);
            if (L.budget_left != ~0ull) --L.budget_left;
            uint32_t post = 0xffffffffu, inworm = 0, extra = 0;
            if (INJ && (long long)L.draws > env.inj_len) {
                extra = SSE_FLAG_STREAM_EXHAUSTED;
                post = WS_DONE;
                if (closed) { L.sweep_visits += L.len; visits += L.len; }
            } else if (closed) {
                L.sweep_visits += L.len;  // total_worm_length += worm_traverse!(...) (sse.jl:197)
                visits += L.len;
                if (--L.worms_left == 0) post = WS_NEED_STREAM;
                else if (L.budget_left == 0) post = WS_DONE;
                else if (!lane_pick_start<INJ>(env, L)) { extra = SSE_FLAG_STREAM_EXHAUSTED; post = WS_DONE; }
            } else if (L.budget_left == 0) {
                post = WS_DONE;  // park in the middle of the worm
                inworm = 1;
            }
            if (post != 0xffffffffu) {
                lane_store(dw, w, L, inworm, extra);
                __threadfence();  // op-code stores and the control block before the hand-over
                st_volatile_shared(status + cur, post);
                if (post == WS_DONE) atomicAdd(n_done, 1u);
                else stream_enqueue(sched, a.nloc_max, cur);  // WS_NEED_STREAM: join the queue of the stream warps
                cur = -1;
            }
        }
        ++it_total;
        const uint32_t fin = __ballot_sync(FULL, finished);
        if (fin == FULL) break;
        if (!__ballot_sync(FULL, cur >= 0)) backoff(256);
    }
    const unsigned long long cyc = (unsigned long long)(clock64() - t_begin);
    visits = warp_sum_u64(visits);
    it_active = warp_sum_u64(it_active);
    if (lane == 0) {
        if (visits) atomicAdd(dw.counters + SSE_CNT_VISITS, visits);
        atomicAdd(dw.counters + SSE_CNT_CYC_WORM, cyc);
        atomicAdd(dw.counters + SSE_CNT_LANE_ITERS, it_active);
        atomicAdd(dw.counters + SSE_CNT_WARP_ITERS, it_total);
    }
}

// Role of a stream warp inside k_sweep: claim walkers that wait for streaming until every walker of the CTA is done.
template <bool INJ>
__device__ __forceinline__ void stream_warp_role(const SmTab &st, const DevModel &dm, const DevWalkers &dw, const SweepArgs &a, uint32_t *sched,
                                              int nloc, uint8_t *scratch, uint32_t lane) {
    // ------------------------------ stream warp: one warp = one walker ------------------------------
    uint32_t *n_done = sched, *q_head = sched + 1, *q_tail = sched + 2, *status = sched + 4, *queue = sched + 4 + a.nloc_max;
    SweepStats ss = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned nap = 256;
    while (true) {
        // take the walker that has waited longest (lane 0 pops the FIFO, all lanes agree); an idle warp only watches the
        // queue counters and sleeps with exponential back-off
        int j = -1;
        uint32_t done = 0;
        const long long t0 = clock64();
        if (lane == 0) {
            while (true) {
                const uint32_t h = ld_volatile_shared(q_head);
                if (h == ld_volatile_shared(q_tail)) break;
                if (atomicCAS(q_head, h, h + 1u) != h) continue;
                uint32_t *slot = queue + (int)(h % (uint32_t)a.nloc_max);
                uint32_t e;
                while ((e = ld_volatile_shared(slot)) == 0u) backoff(32);  // ticket taken, entry about to be written
                st_volatile_shared(slot, 0u);
                j = (int)e - 1;
                break;
            }
            done = ld_volatile_shared(n_done);
        }
        j = __shfl_sync(FULL, j, 0);
        done = __shfl_sync(FULL, done, 0);
        if (j < 0) {
            if (done >= (uint32_t)nloc) break;
            backoff(nap);
            if (nap < 4096) nap *= 2;
            ss.cyc_idle += (unsigned long long)(clock64() - t0);
            continue;
        }
        nap = 256;
        if (lane == 0) st_volatile_shared(status + j, WS_STREAMING);
        __threadfence();  // the worm lane's writes (op codes, control block) are visible
        const int w = (int)blockIdx.x + j * (int)gridDim.x;
        const uint32_t next = stream_task<INJ>(st, dm, dw, a, w, scratch, lane, ss);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            st_volatile_shared(status + j, next);
            if (next == WS_DONE) atomicAdd(n_done, 1u);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (ss.sweeps) {
            atomicAdd(dw.counters + SSE_CNT_SWEEPS, ss.sweeps);
            atomicAdd(dw.counters + SSE_CNT_SUM_N, ss.sum_n);
            atomicAdd(dw.counters + SSE_CNT_SUM_M, ss.sum_M);
        }
        atomicAdd(dw.counters + SSE_CNT_CYC_BUILD, ss.cyc_build);
        atomicAdd(dw.counters + SSE_CNT_CYC_FINISH, ss.cyc_finish);
        atomicAdd(dw.counters + SSE_CNT_CYC_IDLE, ss.cyc_idle);
        atomicAdd(dw.counters + SSE_CNT_TASKS, ss.tasks);
    }
}

template <bool INJ>
__global__ void __launch_bounds__(SWEEP_MAX_WARPS * 32, 1) k_sweep(const DevModel dm, const DevWalkers dw, const SweepArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const SmTab st = stage_tables(dm, smem);
    uint32_t *sched = reinterpret_cast<uint32_t *>(smem + dm.tl.bytes);
    uint32_t *n_done = sched;          // walkers of this CTA that are WS_DONE
    uint32_t *status = sched + 4;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    // walkers of this CTA: w = blockIdx.x + j * gridDim.x
    const int nloc = (dw.W - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (threadIdx.x == 0) { sched[0] = 0; sched[1] = 0; sched[2] = 0; }
    for (int j = threadIdx.x; j < a.nloc_max; j += blockDim.x) sched[4 + a.nloc_max + j] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < nloc; j += blockDim.x) {
        WalkerCtl *ctl = dw.ctl + ((size_t)blockIdx.x + (size_t)j * gridDim.x);
        if (a.reset) {
            ctl->sweeps_left = a.n_sweeps;
        }
        ctl->budget_left = a.budget;
        const uint32_t fl = ctl->flags, ph = ctl->phase;
        uint32_t s;
        if (fl & FATAL_FLAGS) s = WS_DONE;
        else if (ph == 1 && (ctl->worms_left != 0 || ctl->inworm)) s = a.budget ? WS_READY_WORM : WS_DONE;  // parked in its worm phase
        else if (ph == 1) s = WS_NEED_STREAM;
        else s = ((a.reset ? a.n_sweeps : ctl->sweeps_left) > 0 && a.budget) ? WS_NEED_STREAM : WS_DONE;
        status[j] = s;
        if (s == WS_DONE) atomicAdd(n_done, 1u);
        if (s == WS_NEED_STREAM) stream_enqueue(sched, a.nloc_max, j);
    }
    __syncthreads();

    if (SPLIT_REGS) {
        // warps [0, 8) form the two worm warpgroups, warps [8, 24) the four stream warpgroups; a.worm_warps / a.stream_warps
        // of them are active (the others only take part in the register hand-over and leave)
        if (warp < WORM_GROUP_WARPS) {
            regs_shrink<SSE_WORM_REGS>();
            if (warp < a.worm_warps) worm_warp_role<INJ>(st, dm, dw, a, sched, nloc, warp, lane);
        } else {
            regs_grow<SSE_STREAM_REGS>();
            const int sw = warp - WORM_GROUP_WARPS;
            if (sw < a.stream_warps) {
                uint8_t *scratch = smem + dm.tl.bytes + sched_bytes(a.nloc_max) + (size_t)sw * stream_scratch_bytes(dm.n_sites, a.level);
                stream_warp_role<INJ>(st, dm, dw, a, sched, nloc, scratch, lane);
            }
        }
    } else if (warp < a.worm_warps) {
        worm_warp_role<INJ>(st, dm, dw, a, sched, nloc, warp, lane);
    } else if (warp < a.worm_warps + a.stream_warps) {
        uint8_t *scratch = smem + dm.tl.bytes + sched_bytes(a.nloc_max) +
                           (size_t)(warp - a.worm_warps) * stream_scratch_bytes(dm.n_sites, a.level);
        stream_warp_role<INJ>(st, dm, dw, a, sched, nloc, scratch, lane);
    }
}

// ------------------------------------------------------------------------------------------------------
// Single phases (one warp = one walker): Carlo.init!, Carlo.measure! and the parity hooks.
// ------------------------------------------------------------------------------------------------------
enum Mode : int { MODE_INIT = 1, MODE_DIAG, MODE_MAKE_VL, MODE_WORM_UPDATE, MODE_WORM_TRAVERSE, MODE_MEASURE };

struct PhaseArgs {
    int mode, thermalized, warmup, level;
    int l0, w0;       // MODE_WORM_TRAVERSE (0-based leg, 1-based worm)
    long long p0;     // 0-based slot
};

template <bool INJ>
__global__ void __launch_bounds__(PHASE_WARPS * 32) k_phase(const DevModel dm, const DevWalkers dw, const PhaseArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const SmTab st = stage_tables(dm, smem);
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int w = blockIdx.x * PHASE_WARPS + warp;
    if (w >= dw.W) return;
    WalkerCtl *ctl = dw.ctl + w;
    if (__ldcg(&ctl->flags) & FATAL_FLAGS) return;
    const int N = dm.n_sites;
    uint8_t *scratch = smem + dm.tl.bytes + (size_t)warp * stream_scratch_bytes(N, a.level);
    Ctx c = ctx_open(dm, dw, w, scratch, a.level, lane);
    switch (a.mode) {
        case MODE_INIT: {  // Carlo.init! (sse.jl:47-60): M and the empty string are set by the host
            for (int s = lane; s < N; s += 32)
                c.state[s] = (uint8_t)(1u + (uint32_t)sse_uint_below(draw<INJ>(c, c.draws + s), dm.site_dim[s]));
            c.draws += N;
            __syncwarp();
            for (int i = 0; i < a.warmup && !(c.flags & FATAL_FLAGS); ++i) phase_diag_build<INJ>(st, dm, dw, c, true);
            break;
        }
        case MODE_DIAG:
            phase_diag_build<INJ>(st, dm, dw, c, true);
            break;
        case MODE_MAKE_VL:
            phase_diag_build<INJ>(st, dm, dw, c, false);
            break;
        case MODE_WORM_UPDATE:
        case MODE_WORM_TRAVERSE: {
            // lane 0 chases (the worm code is per-lane); the warp then finishes worm_update together
            unsigned long long total = 0, draws = c.draws;
            uint32_t fl = 0;
            long long len = -1;
            if (lane == 0) {
                const LaneEnv env = lane_env(st, dm, dw);
                WormLane L;
                lane_open(dw, w, L);
                L.G = c.G;
                L.M = (uint32_t)c.M;
                L.draws = c.draws;
                if (a.mode == MODE_WORM_UPDATE) {
                    const int nworms = (int)ceil(c.num_worms);
                    for (int wi = 0; wi < nworms && c.n != 0; ++wi) {  // worm_traverse! returns 0 without drawing if n == 0
                        if (!lane_pick_start<INJ>(env, L)) { fl |= SSE_FLAG_STREAM_EXHAUSTED; break; }
                        while (!lane_visit<INJ>(env, L)) {
                            if (INJ && (long long)L.draws > env.inj_len) break;
                        }
                        if (INJ && (long long)L.draws > env.inj_len) { fl |= SSE_FLAG_STREAM_EXHAUSTED; break; }
                        total += L.len;
                    }
                } else {
                    const uint2 wd = lane_ld64(L.words + (a.p0 >> 5));
                    if ((wd.x >> (a.p0 & 31)) & 1u) {
                        lane_set_start(env, L, wd.y + __popc(wd.x & ((1u << (a.p0 & 31)) - 1u)), (uint32_t)a.l0, (uint32_t)a.w0);
                        while (!lane_visit<INJ>(env, L)) {
                            if (INJ && (long long)L.draws > env.inj_len) break;
                        }
                        if (INJ && (long long)L.draws > env.inj_len) fl |= SSE_FLAG_STREAM_EXHAUSTED;
                        len = (long long)L.len;
                        total = L.len;
                    }
                    dw.dbg_len[w] = len;
                }
                if (L.fell) fl |= SSE_FLAG_SCATTER_FALLTHROUGH;
                draws = L.draws;
                if (total) atomicAdd(dw.counters + SSE_CNT_VISITS, total);
                __threadfence();
            }
            total = __shfl_sync(FULL, total, 0);
            c.draws = __shfl_sync(FULL, draws, 0);
            c.flags |= __shfl_sync(FULL, fl, 0);
            if (a.mode == MODE_WORM_UPDATE && !(c.flags & FATAL_FLAGS))
                worm_finish<INJ>(st, dm, dw, c, a.thermalized != 0, w, 1.0 + (double)total);
            break;
        }
        case MODE_MEASURE:
            phase_measure(st, dm, dw, c, dw.obs_out + (size_t)w * dw.n_obs);
            break;
    }
    if (INJ && (long long)c.draws > c.inj_len) c.flags |= SSE_FLAG_STREAM_EXHAUSTED;
    ctx_close(dm, dw, w, a.level, c);
}

// ------------------------------------------------------------------------------------------------------
// beta doubling (thermalisation aid; NOT part of the reference): for a periodic configuration (state, S_M) the
// doubled string S_M S_M with the same state is a valid configuration at inverse temperature 2*beta with 2n
// operators, so a cold walker can be grown from a cheap hot one in log2(beta) steps instead of thousands of
// full-size sweeps.  Needs walkers between sweeps.  One warp per walker; M, n and the controller's average worm
// length double, T halves.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PHASE_WARPS * 32) k_double_beta(const DevWalkers dw) {
    const int w = blockIdx.x * PHASE_WARPS + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (w >= dw.W) return;
    WalkerCtl *ctl = dw.ctl + w;
    const uint32_t flags = __ldcg(&ctl->flags);
    const long long M = __ldcg(&ctl->M), n = __ldcg(&ctl->n);
    const uint32_t G = __ldcg(&ctl->G), Rcap = (uint32_t)dw.R_cap;
    const double T = __ldcg(&ctl->T), awl = __ldcg(&ctl->avg_wl);
    __syncwarp();
    if (flags & FATAL_FLAGS) return;
    if (2 * M > dw.M_cap || 2 * n > dw.n_cap) {
        if (lane == 0) {
            ctl->flags = flags | (2 * M > dw.M_cap ? SSE_FLAG_M_OVERFLOW : SSE_FLAG_N_OVERFLOW);
            atomicOr(reinterpret_cast<unsigned long long *>(dw.counters + SSE_CNT_ANY_FATAL), 1ull);
        }
        return;
    }
    uint2 *words = dw.words + (size_t)w * dw.Mw_cap;
    uint4 *rec = dw.rec + (size_t)w * dw.R_cap;
    // slot p of the doubled string holds what slot p mod M held; a rewritten word only gains bits at slots >= M, so the
    // words can be rewritten in place, front to back
    const int nchunks = (int)((2 * M + 31) >> 5);
    uint32_t kbase = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        const long long p = (long long)ch * 32 + lane;
        bool bit = false;
        if (p < 2 * M) {
            const long long src = p < M ? p : p - M;
            bit = (__ldcg(&words[src >> 5].x) >> (src & 31)) & 1u;
        }
        const uint32_t nb = __ballot_sync(FULL, bit);
        __syncwarp();
        if (lane == 0) words[ch] = make_uint2(nb, kbase);
        kbase += __popc(nb);
        __syncwarp();
    }
    for (long long k = lane; k < n; k += 32) rec[ring(G, Rcap, (uint32_t)(n + k))] = __ldcg(&rec[ring(G, Rcap, (uint32_t)k)]);
    if (lane == 0) {
        ctl->M = (int)(2 * M);
        ctl->n = (int)(2 * n);
        ctl->T = T * 0.5;
        // worm-count controller (sse.jl:204-217): worms get at least twice as long at twice the inverse temperature;
        // carrying the old average over would keep num_worms (target = twlf * n / avg_wl) far too high for many sweeps
        ctl->avg_wl = awl * 2.0;
    }
}

}  // namespace sse
