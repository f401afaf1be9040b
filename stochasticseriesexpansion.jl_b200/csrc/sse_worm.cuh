// sse_worm.cuh — the worm update (src/sse.jl:193-303) executed by ONE LANE for ONE walker.
// Every lane of a worm warp carries its own dependent chain: its own record addresses, its own Philox stream position,
// its own worm registers.  Per visit: one 16-byte ld.global.cg (the record: op code + 4 leg links = half a DRAM sector),
// one 16-byte ld.shared (transition entry fused with its first outcome), one f64 compare against the lane's next uniform
// draw, one 4-byte st.global (the new op code, into the record just read).  The next record's load is issued as soon as
// the exit leg is known; the store, the stop tests and the next draw run in its shadow.
#pragma once
#include "sse_common.cuh"

namespace sse {

struct WormLane {
    // walker (constant while the lane holds it)
    uint4 *rec;
    const uint2 *words;
    const unsigned long long *inj;
    unsigned long long wid;
    uint32_t G, M;
    // stream position; one Philox block yields two draws, the odd one is kept
    unsigned long long draws, spare;
    uint32_t have_spare;
    // the worm in flight (worm_traverse!, src/sse.jl:262-303)
    uint4 R;          // the current record (requested, maybe still in flight)
    uint4 *cur;       // its address
    uint32_t pos, wf, pos0, w0, patch, pval;
    unsigned long long len;
    // the sweep in flight
    uint32_t worms_left, fell;
    unsigned long long sweep_visits, budget_left;
};

struct LaneEnv {  // launch constants of the worm lanes
    unsigned long long pol;  // L2 policy of the record accesses (evict_first)
    uint32_t t1_s, outc_s, maxw4, Rcap;
    unsigned long long seed;
    long long inj_len;
    const uint4 *bond_info;
};

__device__ __forceinline__ LaneEnv lane_env(const SmTab &st, const DevModel &dm, const DevWalkers &dw) {
    LaneEnv e;
    e.pol = policy_evict_first();
    e.t1_s = st.t1_s;
    e.outc_s = st.outc_s;
    e.maxw4 = (uint32_t)dm.max_worm * 4u;
    e.Rcap = (uint32_t)dw.R_cap;
    e.seed = dw.seed;
    e.inj_len = dw.inj_len;
    e.bond_info = dm.bond_info;
    return e;
}

// next raw draw of the lane's stream (sse_rng.h: draw 2j = words (0,1), draw 2j+1 = words (2,3) of Philox block j)
template <bool INJ>
__device__ __forceinline__ uint64_t lane_draw(const LaneEnv &e, WormLane &L) {
    const unsigned long long k = L.draws++;
    if (INJ) return (long long)k < e.inj_len ? (uint64_t)__ldg(L.inj + k) : 0ull;
    if ((k & 1ull) && L.have_spare) {
        L.have_spare = 0;
        return L.spare;
    }
    uint32_t b[4];
    sse_philox_block(e.seed, L.wid, k >> 1, b);
    const uint64_t x0 = (uint64_t)b[0] | ((uint64_t)b[1] << 32), x1 = (uint64_t)b[2] | ((uint64_t)b[3] << 32);
    if (k & 1ull) return x1;
    L.spare = x1;
    L.have_spare = 1;
    return x0;
}

// Take over walker w: everything the lane needs, from the walker's control block.
__device__ __forceinline__ void lane_open(const DevWalkers &dw, int w, WormLane &L) {
    const WalkerCtl *ctl = dw.ctl + w;
    L.rec = dw.rec + (size_t)w * dw.R_cap;
    L.words = dw.words + (size_t)w * dw.Mw_cap;
    L.inj = dw.inj ? dw.inj + (size_t)w * dw.inj_len : nullptr;
    L.wid = dw.wid_off + (unsigned long long)w;
    L.G = __ldcg(&ctl->G);
    L.M = (uint32_t)__ldcg(&ctl->M);
    L.draws = __ldcg(&ctl->draws);
    L.spare = 0;
    L.have_spare = 0;
    L.worms_left = __ldcg(&ctl->worms_left);
    L.sweep_visits = __ldcg(&ctl->sweep_visits);
    L.budget_left = __ldcg(&ctl->budget_left);
    L.fell = 0;
    L.patch = 0;
    L.pval = 0;
    L.pos = __ldcg(&ctl->pos);
    L.wf = __ldcg(&ctl->wf);
    L.pos0 = __ldcg(&ctl->pos0);
    L.w0 = __ldcg(&ctl->w0);
    L.len = __ldcg(&ctl->worm_len);
    L.cur = L.rec;
    L.R = make_uint4(0, 0, 0, 0);
}
// resume a parked worm: request its current record again
__device__ __forceinline__ void lane_resume(const LaneEnv &e, WormLane &L) {
    L.cur = L.rec + ring(L.G, e.Rcap, L.pos >> 2);
    L.R = lane_ld128(L.cur, e.pol);
}
// Give the walker back: stream position, sweep progress and (inworm) the worm in flight.
__device__ __forceinline__ void lane_store(const DevWalkers &dw, int w, const WormLane &L, uint32_t inworm, uint32_t extra_flags) {
    WalkerCtl *ctl = dw.ctl + w;
    ctl->draws = L.draws;
    ctl->worms_left = L.worms_left;
    ctl->sweep_visits = L.sweep_visits;
    ctl->budget_left = L.budget_left;
    ctl->inworm = inworm;
    if (inworm) {
        ctl->pos = L.pos;
        ctl->wf = L.wf;
        ctl->pos0 = L.pos0;
        ctl->w0 = L.w0;
        ctl->worm_len = L.len;
    }
    uint32_t fl = extra_flags;
    if (L.fell) fl |= SSE_FLAG_SCATTER_FALLTHROUGH;
    if (fl) {
        ctl->flags = __ldcg(&ctl->flags) | fl;
        if (fl & FATAL_FLAGS) atomicOr(reinterpret_cast<unsigned long long *>(dw.counters + SSE_CNT_ANY_FATAL), 1ull);
    }
}

// worm_traverse! outer, start selection (src/sse.jl:241-251): rejection loop over (slot, leg) until the slot holds an
// operator, then the worm type.  Leaves the worm at its start with the first record requested.  Returns false if the
// injected stream ran out.
template <bool INJ>
__device__ __forceinline__ bool lane_pick_start(const LaneEnv &e, WormLane &L) {
    uint32_t k0, l0;
    while (true) {
        if (INJ && (long long)L.draws >= e.inj_len) return false;
        const uint32_t p0 = (uint32_t)sse_uint_below(lane_draw<INJ>(e, L), (uint64_t)L.M);  // rand(rng, 1:M) - 1 (sse.jl:242)
        l0 = (uint32_t)sse_uint_below(lane_draw<INJ>(e, L), 4u);                            // rand(rng, 1:leg_count) - 1 (:243)
        const uint2 wd = lane_ld64(L.words + (p0 >> 5));
        if ((wd.x >> (p0 & 31u)) & 1u) {                                                     // vertices[l0, p0][1] > 0 (:244)
            k0 = wd.y + __popc(wd.x & ((1u << (p0 & 31u)) - 1u));
            break;
        }
    }
    L.pos0 = (k0 << 2) | l0;
    L.pos = L.pos0;
    L.cur = L.rec + ring(L.G, e.Rcap, k0);
    L.R = lane_ld128(L.cur, e.pol);
    const uint4 bi = __ldg(e.bond_info + op_bond(L.R.x));
    const uint32_t dim0 = (l0 & 1u) ? (bi.y >> 24) : (bi.x >> 24);                // site_of_leg (sse.jl:250)
    L.w0 = 1u + (uint32_t)sse_uint_below(lane_draw<INJ>(e, L), dim0 - 1u);        // sse.jl:251
    L.wf = L.w0;
    L.len = 1;
    L.patch = 0;
    return true;
}
// worm_traverse!((l0, p0, wormfunc0), ...) with an explicit start (parity hook): k0 = record index, 0-based leg
__device__ __forceinline__ void lane_set_start(const LaneEnv &e, WormLane &L, uint32_t k0, uint32_t l0, uint32_t w0) {
    L.pos0 = (k0 << 2) | l0;
    L.pos = L.pos0;
    L.cur = L.rec + ring(L.G, e.Rcap, k0);
    L.R = lane_ld128(L.cur, e.pol);
    L.w0 = w0;
    L.wf = w0;
    L.len = 1;
    L.patch = 0;
}

// One visit (the body of the reference's `while true`, src/sse.jl:274-300) with scatter (src/vertex_data.jl:106-125).
// Returns true when the worm closed.
template <bool INJ>
__device__ __forceinline__ bool lane_visit(const LaneEnv &e, WormLane &L) {
    const double r = sse_u01(lane_draw<INJ>(e, L));  // rand(rng) (sse.jl:282); independent of the record in flight
    const uint32_t pos = L.pos;
    const uint4 Rc = L.R;
    const uint32_t x = L.patch ? L.pval : Rc.x;
    // transitions[leg_in, worm_in, vi] fused with its first outcome (vertex_data.jl:115-123)
    uint4 t = lds128(e.t1_s + 16u * (op_gv(x) * e.maxw4 + (((L.wf - 1u) << 2) | (pos & 3u))));
    if (!(r < __hiloint2double((int)t.y, (int)t.x))) {
        const uint32_t off = (t.w >> 6) & 0x3ffffu, cnt = t.w & 63u;
        bool hit = false;
        for (uint32_t j = 0; !hit && j < cnt; ++j) {  // first out with random < cumprob
            t = lds128(e.outc_s + 16u * (off + j));
            hit = r < __hiloint2double((int)t.y, (int)t.x);
        }
        if (!hit) L.fell = 1;  // vertex_data.jl:124; clamped to the last outcome
    }
    const uint32_t leg_out = (t.z >> 16) & 3u;
    const uint32_t posn = rec_link(Rc, leg_out);  // (leg_in, p) = vertices[leg_out, p] (sse.jl:295)
    uint4 *const nxt = L.rec + ring(L.G, e.Rcap, posn >> 2);
    L.R = lane_ld128(nxt, e.pol);
    // ---- everything below overlaps with the load ----
    const uint32_t newop = (x & ~(VMASK | 2u)) | (t.z & (VMASK | 2u));  // OperCode(bond, new_vertex) (sse.jl:285)
    lane_st32(L.cur, newop, e.pol);
    const uint32_t w_out = t.z >> 24, dim_out = t.w >> 24;
    const bool stop1 = (((pos & ~3u) | leg_out) == L.pos0) && (w_out + L.w0 == dim_out);  // sse.jl:288-290
    L.len += stop1 ? 0u : 1u;
    L.wf = w_out;
    L.patch = ((posn >> 2) == (pos >> 2)) ? 1u : 0u;  // the link re-enters this record: its load preceded the store
    L.pval = newop;
    L.pos = posn;
    L.cur = nxt;
    const bool stop2 = (posn == L.pos0) && (w_out == L.w0);  // sse.jl:297-299
    return stop1 || stop2;
}

}  // namespace sse
