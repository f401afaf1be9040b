// sse_stream.cuh — the streaming phases, executed by ONE WARP for ONE walker (32 string slots per step):
//   phase_diag_build   diagonal_update (src/sse.jl:137-191) fused with make_vertex_list! (src/vertex_list.jl:15-54)
//   worm_finish        the end of worm_update (src/sse.jl:200-228): WormLengthFraction, controller, state rebuild
//   phase_measure      Carlo.measure! (src/sse.jl:70-87): measure_sign, scalar observables, measure_opstring!
// Warp primitives do the scans: ballot/popc prefix sums give each slot its random-stream offset (2/1/0 draws by
// pre-update slot type) and its compact record index; one Philox block per lane feeds a whole chunk; shuffles resolve
// same-site ordering only in the rare chunks where two operators share a site.
#pragma once
#include "sse_common.cuh"

namespace sse {

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
// Draw number `rel` of the scratch filled by fill_draws(j0), i.e. draw k = 2*j0 + rel of the stream (or of the injected one).
template <bool INJ>
__device__ __forceinline__ uint64_t scratch_draw(const Ctx &c, uint32_t rel, unsigned long long k) {
    if (INJ) return (long long)k < c.inj_len ? (uint64_t)__ldg(c.inj + k) : 0ull;
    return c.rng[rel];
}

// Sequential reader of a walker's operator string: per 32-slot chunk every lane gets the op code of its slot (0 =
// identity), gathered from the record ring through the occupancy bitmap.  The bitmap words of 32 chunks are held one
// per lane (the next 32 are already requested).  The op codes come from DRAM; they are requested OP_AHEAD chunks before
// their use with asynchronous copies into a small shared-memory ring, one group per chunk (measured with plain loads:
// the register moves that rotate a software pipeline wait for the load they move, which shortened the lead to one chunk
// and left the pass waiting on the scoreboard for a quarter of its time).
// Group accounting: request() leaves the commit to the caller, which commits exactly one op group and `extra` other
// groups per chunk, in that order; take() waits until the chunk's own group has landed.
template <int EXTRA>
struct OpReader {
    const uint4 *rec;
    const uint2 *words;
    uint32_t G, Rcap, lane, lt;
    uint32_t dst_s;             // shared-space address of this lane's word in slot 0 of the op-code ring
    unsigned long long pol;     // L2 policy of the op-code gathers: read once (evict_first)
    int nchunks, next_req;
    // bitmap words, one per lane: of the block of 32 chunks before the one being requested from, of that block, and of the
    // next one (already requested).  take() lags request() by OP_AHEAD < 32 chunks, so its chunk is in wprev or wcur.
    uint32_t wprev, wcur, wnext;
    uint32_t kreq;              // first record of the next chunk to request
    uint32_t kold;              // first record of the next chunk to take

    __device__ __forceinline__ uint32_t word_at(int ch) const { return ch < nchunks ? __ldcg(&words[ch].x) : 0u; }
    // request the next chunk (chunks are requested in order 0, 1, 2, ...); the caller commits the group
    __device__ __forceinline__ void request() {
        const int c = next_req++;
        if ((c & 31) == 0 && c > 0) {  // c opens a new block: the one after it is requested
            wprev = wcur;
            wcur = wnext;
            wnext = word_at(c + 32 + (int)lane);
        }
        uint32_t bits = __shfl_sync(FULL, wcur, c & 31);
        if (c >= nchunks) bits = 0u;
        const bool have = (bits >> lane) & 1u;
        const uint32_t idx = have ? ring(G, Rcap, kreq + __popc(bits & lt)) : 0u;
        cp_async4(dst_s + 128u * (uint32_t)(c & (OP_RING - 1)), rec + idx, have, pol);
        kreq += __popc(bits);
    }
    __device__ __forceinline__ void init(const uint2 *words_, const uint4 *rec_, uint32_t G_, uint32_t Rcap_, int nchunks_,
                                         uint32_t lane_, uint32_t ring_s_) {
        words = words_;
        rec = rec_;
        G = G_;
        Rcap = Rcap_;
        lane = lane_;
        lt = lanemask_lt();
        dst_s = ring_s_ + 4u * lane_;
        pol = policy_evict_first();
        nchunks = nchunks_;
        next_req = 0;
        wprev = 0;
        wcur = word_at((int)lane);
        wnext = word_at(32 + (int)lane);
        kreq = 0;
        kold = 0;
        for (int i = 0; i < OP_AHEAD; ++i) {  // the pipeline's lead: chunks 0 .. OP_AHEAD-1, with the group pattern of a chunk
            request();
            cp_async_commit();
            for (int e = 0; e < EXTRA; ++e) cp_async_commit();
        }
    }
    // chunk ch (called for ch = 0, 1, 2, ... in order, BEFORE this iteration's request()): bits = occupancy of its 32
    // slots, op = this lane's op code; returns the index of the chunk's first record
    __device__ __forceinline__ uint32_t take(int ch, uint32_t &bits, uint32_t &op) {
        // groups committed after chunk ch's own: EXTRA of its iteration + (OP_AHEAD - 1) later chunks x (1 + EXTRA)
        cp_async_wait<EXTRA + (OP_AHEAD - 1) * (1 + EXTRA)>();
        // the requests are at chunk next_req - 1 = ch + OP_AHEAD - 1: same block as ch, or the one after
        const uint32_t w = ((next_req - 1) >> 5) == (ch >> 5) ? wcur : wprev;
        bits = __shfl_sync(FULL, w, ch & 31);
        op = lds32(dst_s + 128u * (uint32_t)(ch & (OP_RING - 1)));
        const uint32_t k0 = kold;
        kold += __popc(bits);
        return k0;
    }
};

// ------------------------------------------------------------------------------------------------------
// diagonal_update (src/sse.jl:137-191) fused with make_vertex_list! (src/vertex_list.jl:15-54).
// do_diag = false rebuilds the records of the unchanged string (make_vertex_list! alone).
// Reads generation g of the record ring (op codes only) and writes generation g+1 (op codes + links) behind it.
//
// The pass is a two-stage software pipeline over 32-slot chunks, because a warp that waits for every global load in
// turn spends its time on L2 round trips (measured: 7400 cycles per chunk).  What a chunk needs from global memory
// depends only on the OLD string: the random-stream offset of a slot follows from the pre-update slot types before it
// (2 draws per identity slot, 1 per diagonal operator, Appendix A), so the draws, the proposed bonds and the bond-table
// rows of chunk c+1 are fetched (stage B) while chunk c is decided and linked (stage C) from registers and shared memory.
// ------------------------------------------------------------------------------------------------------
// make_vertex_list! (src/vertex_list.jl:15-54) for the next `m` <= 32 operators of the new generation, records k0 .. k0+m-1,
// taken from the warp's queue {op code, site a, site b}: lane i links operator k0 + i.  The diagonal update queues the
// operators it leaves in the string (about 13 per 32-slot chunk at the BASELINE sizes), so this step runs with all lanes
// busy instead of once per chunk with most lanes idle.
constexpr uint32_t BUILD_QUEUE = 64;  // entries; at most 31 waiting + 32 new ones

// Out of line on purpose, like every cold path of this file: it keeps the chunk loop of the diagonal update small (10 KB
// of SASS instead of 60 KB) and its register allocation free of the link passes' live ranges.  Out-of-line helpers return
// their results by value: a reference parameter would pin the caller's variable in local memory.
struct BuildArgs {
    unsigned long long pol;  // L2 policy of the record stores and link patches (evict_last)
    uint32_t *queue;
    uint8_t *mark;
    uint32_t *vfirst, *vlast;
    uint4 *rec;
    uint32_t Rcap, Gn, lane;
};

// The links are built WITHOUT touching a record twice at random.  A first version patched the forward link of the previous
// operator on a site when the next one came by (as make_vertex_list! does, vertex_list.jl:36-38): 2 scattered partial
// stores per operator into sectors written ~N/2 records earlier, which had usually left L2 by then — a DRAM
// read-modify-write each, as much DRAM traffic as the whole worm phase (ncu: 97 of 251 GB per launch).  Instead:
//   forward  (build_records, slot order):   record k = {op, backward links, site a, site b}, the previous operator on each site from vlast[];
//   backward (finish_links, reverse order): forward links from vnext[site] = first leg of the next operator on the site,
//                                           which is known because the later records were handled first.
// Both passes read and write the ring sequentially.  vnext[] lives in the vfirst[] array: it starts as vfirst (the world
// line is periodic: after the last operator comes the first, vertex_list.jl:46-51) and ends as vfirst again.

__device__ __forceinline__ unsigned long long pack_links(uint32_t a, uint32_t b, bool fa, bool fb) {
    return (unsigned long long)a | ((unsigned long long)b << 24) | ((unsigned long long)fa << 48) | ((unsigned long long)fb << 49);
}
// nearest earlier (pa/pb: its top leg) operator of the group on each of a lane's sites, and whether a later one exists
// (all out-of-line helpers return their results packed in registers: a reference parameter of a __noinline__ function
// pins the caller's variable in local memory, and the hot loops then wait on LDL/STL round trips — measured)
__device__ __noinline__ unsigned long long group_resolve_earlier(uint32_t inv_mask, uint32_t k0, uint32_t lane, bool nn, uint32_t sa,
                                                                 uint32_t sb) {
    uint32_t pa = NONE24, pb = NONE24;
    bool later_a = false, later_b = false;
    for (uint32_t mm = inv_mask; mm;) {
        const int L = __ffs(mm) - 1;
        mm &= mm - 1;
        const uint32_t qa = __shfl_sync(FULL, sa, L), qb = __shfl_sync(FULL, sb, L);
        const uint32_t qk = (k0 + (uint32_t)L) << 2;
        if (nn && (int)lane > L) {
            if (sa == qa) pa = qk | 2u;
            if (sa == qb) pa = qk | 3u;
            if (sb == qa) pb = qk | 2u;
            if (sb == qb) pb = qk | 3u;
        } else if (nn && (int)lane < L) {
            if (sa == qa || sa == qb) later_a = true;
            if (sb == qa || sb == qb) later_b = true;
        }
    }
    return pack_links(pa, pb, later_a, later_b);
}
// nearest later (sua/sub: its bottom leg) operator of the group on each of a lane's sites, and whether an earlier one exists
__device__ __noinline__ unsigned long long group_resolve_later(uint32_t inv_mask, uint32_t k0, uint32_t lane, bool nn, uint32_t sa,
                                                               uint32_t sb) {
    uint32_t sua = NONE24, sub = NONE24;
    bool earlier_a = false, earlier_b = false;
    for (uint32_t mm = inv_mask; mm;) {
        const int L = __ffs(mm) - 1;
        mm &= mm - 1;
        const uint32_t qa = __shfl_sync(FULL, sa, L), qb = __shfl_sync(FULL, sb, L);
        const uint32_t qk = (k0 + (uint32_t)L) << 2;
        if (nn && (int)lane < L) {
            if (sua == NONE24) { if (sa == qa) sua = qk; else if (sa == qb) sua = qk | 1u; }
            if (sub == NONE24) { if (sb == qa) sub = qk; else if (sb == qb) sub = qk | 1u; }
        } else if (nn && (int)lane > L) {
            if (sa == qa || sa == qb) earlier_a = true;
            if (sb == qa || sb == qb) earlier_b = true;
        }
    }
    return pack_links(sua, sub, earlier_a, earlier_b);
}
// operators of the group that share a site with another one: every operator tags its two sites, a lost tag reveals a
// collision, the losers flag the contested sites, and every operator on a flagged site takes part in the search
__device__ __forceinline__ uint32_t group_collisions(uint8_t *mark, uint32_t lane, bool nn, uint32_t sa, uint32_t sb) {
    if (nn) {
        mark[sa] = (uint8_t)lane;
        mark[sb] = (uint8_t)lane;
    }
    __syncwarp();
    const bool lost_a = nn && mark[sa] != (uint8_t)lane, lost_b = nn && mark[sb] != (uint8_t)lane;
    if (!__ballot_sync(FULL, lost_a || lost_b)) return 0u;
    __syncwarp();
    if (lost_a) mark[sa] = 0x7f;
    if (lost_b) mark[sb] = 0x7f;
    __syncwarp();
    const bool inv = nn && (mark[sa] == 0x7f || mark[sb] == 0x7f);
    return __ballot_sync(FULL, inv);
}

// forward pass: records k0 .. k0+m-1 = {op code, backward links} from the warp's queue {op code, site a, site b}
__device__ __noinline__ void build_records(const BuildArgs b, uint32_t k0, uint32_t m) {
    const uint32_t lane = b.lane, Rcap = b.Rcap, Gn = b.Gn;
    const bool nn = lane < m;
    const uint32_t k = k0 + lane, q = k & (BUILD_QUEUE - 1);
    uint32_t newop = 0, sa = 0, sb = 0;
    if (nn) {
        newop = b.queue[q];
        sa = b.queue[BUILD_QUEUE + q];
        sb = b.queue[2 * BUILD_QUEUE + q];
    }
    uint32_t pa = NONE24, pb = NONE24;
    bool later_a = false, later_b = false;
    const uint32_t inv = group_collisions(b.mark, lane, nn, sa, sb);
    if (inv) {
        const unsigned long long r = group_resolve_earlier(inv, k0, lane, nn, sa, sb);
        pa = (uint32_t)r & NONE24;
        pb = (uint32_t)(r >> 24) & NONE24;
        later_a = (r >> 48) & 1ull;
        later_b = (r >> 49) & 1ull;
    }
    uint32_t ma = NONE32, mb = NONE32;
    if (nn) {
        if (pa == NONE24) ma = b.vlast[sa];
        if (pb == NONE24) mb = b.vlast[sb];
    }
    __syncwarp();
    if (nn) {
        const uint32_t me = k << 2;
        uint32_t bla = pa, blb = pb;  // NONE24 = first operator on the site: closed by finish_links
        if (pa == NONE24) {
            if (ma != NONE32) bla = ma;       // vertices[s,p] = (s1,p1) (vertex_list.jl:36-38)
            else b.vfirst[sa] = me;           // vertex_list.jl:40
        }
        if (pb == NONE24) {
            if (mb != NONE32) blb = mb;
            else b.vfirst[sb] = me | 1u;
        }
        if (!later_a) b.vlast[sa] = me | 2u;  // vertex_list.jl:42
        if (!later_b) b.vlast[sb] = me | 3u;
        // the two forward-link fields carry the operator's sites to the backward pass (saves its bond-table lookup)
        st128_hint(b.rec + ring(Gn, Rcap, k), rec_pack(newop, bla, blb, sa, sb), b.pol);
    }
    __syncwarp();
}

// backward pass over the n records of the new generation: forward links, and the periodic closure of the world lines.
// Software pipeline over the 32-record groups, so that no global round trip is exposed: the records are requested two groups
// ahead, and the vnext[] entries of the NEXT group are gathered before the current group stores its own entries.  Where the
// current group writes a site the next group reads (most groups have one), the early value is replaced by the one the
// current group stored, taken from the writer's registers: the tag array names the writer (tags are reset to MARK_FREE
// behind every group, so a tag can only come from the current group).
constexpr uint32_t MARK_FREE = 0x40u;  // no lane id (0..31), not the collision flag (0x7f), bit 7 clear (the diagonal update's tags)
__device__ __noinline__ void finish_links(const BuildArgs b, uint32_t n, unsigned long long pol_final, int n_sites) {
    const uint32_t lane = b.lane, Rcap = b.Rcap, Gn = b.Gn;
    uint32_t *vnext = b.vfirst;
    uint8_t *mark = b.mark;
    for (int s = (int)lane; s < n_sites; s += 32) mark[s] = (uint8_t)MARK_FREE;
    __syncwarp();
    uint32_t k_hi = n;
    uint32_t k0 = k_hi > 32u ? k_hi - 32u : 0u, m = k_hi - k0;
    uint32_t nk0 = k0 > 32u ? k0 - 32u : 0u, nm = k0 - nk0;  // the next (earlier) group
    uint4 Rc = make_uint4(0, 0, 0, 0), Rn = make_uint4(0, 0, 0, 0);
    if (lane < m) Rc = __ldcg(b.rec + ring(Gn, Rcap, k0 + lane));
    if (lane < nm) Rn = __ldcg(b.rec + ring(Gn, Rcap, nk0 + lane));
    uint32_t sa = lane < m ? rec_link(Rc, 2) : 0u, sb = lane < m ? rec_link(Rc, 3) : 0u;  // left there by build_records
    uint32_t na = 0, nb = 0;
    if (lane < m) {
        na = vnext[sa];
        nb = vnext[sb];
    }
    while (k_hi > 0u) {
        const bool nn = lane < m, nn2 = lane < nm;
        const uint32_t k = k0 + lane;
        const uint32_t nnk0 = nk0 > 32u ? nk0 - 32u : 0u, nnm = nk0 - nnk0;  // the group after the next one: request its records
        uint4 Rnn = make_uint4(0, 0, 0, 0);
        if (lane < nnm) Rnn = __ldcg(b.rec + ring(Gn, Rcap, nnk0 + lane));
        // the next group's sites and their vnext[] entries as they are BEFORE this group's stores
        const uint32_t sa2 = nn2 ? rec_link(Rn, 2) : 0u, sb2 = nn2 ? rec_link(Rn, 3) : 0u;
        uint32_t na2 = 0, nb2 = 0;
        if (nn2) {
            na2 = vnext[sa2];
            nb2 = vnext[sb2];
        }
        // ---- this group ----
        uint32_t sua = NONE24, sub = NONE24;
        bool earlier_a = false, earlier_b = false;
        const uint32_t inv = group_collisions(mark, lane, nn, sa, sb);
        if (inv) {
            const unsigned long long r = group_resolve_later(inv, k0, lane, nn, sa, sb);
            sua = (uint32_t)r & NONE24;
            sub = (uint32_t)(r >> 24) & NONE24;
            earlier_a = (r >> 48) & 1ull;
            earlier_b = (r >> 49) & 1ull;
        }
        if (nn) {
            uint32_t bla = rec_link(Rc, 0), blb = rec_link(Rc, 1);
            if (bla == NONE24) bla = b.vlast[sa];  // first operator on the site: its lower neighbour is the last one (vertex_list.jl:46-51)
            if (blb == NONE24) blb = b.vlast[sb];
            if (sua == NONE24) sua = na;
            if (sub == NONE24) sub = nb;
            if (!earlier_a) vnext[sa] = k << 2;
            if (!earlier_b) vnext[sb] = (k << 2) | 1u;
            st128_hint(b.rec + ring(Gn, Rcap, k), rec_pack(Rc.x, bla, blb, sua, sub), pol_final);
        }
        __syncwarp();
        // ---- the next group's early values, corrected where this group wrote ----
        const uint32_t ta = nn2 ? mark[sa2] : MARK_FREE, tb = nn2 ? mark[sb2] : MARK_FREE;
        const uint32_t qa = __shfl_sync(FULL, sa, ta & 31u), qb = __shfl_sync(FULL, sa, tb & 31u);  // site a of the writers
        if (ta != MARK_FREE) {
            if (ta == 0x7fu) na2 = vnext[sa2];  // several writers in the group (rare): read what they left
            else na2 = ((k0 + ta) << 2) | (sa2 == qa ? 0u : 1u);
        }
        if (tb != MARK_FREE) {
            if (tb == 0x7fu) nb2 = vnext[sb2];
            else nb2 = ((k0 + tb) << 2) | (sb2 == qb ? 0u : 1u);
        }
        __syncwarp();
        if (nn) {
            mark[sa] = (uint8_t)MARK_FREE;
            mark[sb] = (uint8_t)MARK_FREE;
        }
        __syncwarp();
        k_hi = k0;
        k0 = nk0;
        m = nm;
        nk0 = nnk0;
        nm = nnm;
        Rc = Rn;
        Rn = Rnn;
        sa = sa2;
        sb = sb2;
        na = na2;
        nb = nb2;
    }
}

// The in-order recurrence of the accept tests (sse.jl:164-166,176-178) for a chunk in which the bound test left some lane
// undecided: lane l only depends on lanes < l, so after i rounds the first i lanes are final (rare: ~600/(M-n) per chunk).
__device__ __noinline__ unsigned long long diag_resolve_exact(int n, int M, double p_make_bond_raw, double p_remove_bond_raw, bool is_id,
                                                              bool is_dg, double r, double w, uint32_t lt) {
    uint32_t ins = 0, rem = 0;
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
    return (unsigned long long)ins | ((unsigned long long)rem << 32);
}

// accept thresholds for an operator count anywhere in [lo, hi] (two divisions; recomputed every few chunks)
__device__ __noinline__ double diag_window_make(int M, int at, double p_make_bond_raw) {
    return (M - at > 0) ? p_make_bond_raw / (double)(M - at) : __longlong_as_double(0x7ff0000000000000ll);
}

// state seen by identity lanes / state written by off-diagonal lanes when operators of one chunk share sites (rare)
// in/out packed: s_a | s_b << 8 | wa << 16 | wb << 17
__device__ __noinline__ uint32_t diag_resolve_state(uint32_t offm, uint32_t lane, bool is_id, bool is_off, uint32_t sa, uint32_t sb, uint32_t ta,
                                                    uint32_t tb, uint32_t packed) {
    uint32_t s_a = packed & 0xffu, s_b = (packed >> 8) & 0xffu;
    bool wa = (packed >> 16) & 1u, wb = (packed >> 17) & 1u;
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
    return s_a | (s_b << 8) | ((uint32_t)wa << 16) | ((uint32_t)wb << 17);
}

// one more Philox block behind the 32 a chunk's lanes computed (needed when 64 draws start at an odd stream position)
__device__ __noinline__ void philox_extra_block(unsigned long long seed, unsigned long long wid, unsigned long long j, uint4 *dst) {
    uint32_t b[4];
    sse_philox_block(seed, wid, j, b);
    *dst = make_uint4(b[0], b[1], b[2], b[3]);
}

struct ChunkIn {  // what stage B hands to stage C (the bond-table rows go through shared memory)
    uint32_t op, bond, idm, dgm, kold0;
    double r;
};

template <bool INJ>
__device__ __forceinline__ ChunkIn diag_stage_b(const DevModel &dm, const Ctx &c, OpReader<1> &rd, int ch, int M, bool do_diag,
                                                unsigned long long &draws) {
    const uint32_t lane = c.lane, lt = rd.lt;
    ChunkIn in;
    uint32_t obits;
    in.kold0 = rd.take(ch, obits, in.op);
    rd.request();  // op codes of chunk ch + OP_AHEAD
    cp_async_commit();
    const int p = ch * 32 + (int)lane;
    const bool nonid = in.op != 0u;
    const bool is_id = (p < M) && !nonid;
    const bool is_dg = nonid && (in.op & 2u);
    in.bond = op_bond(in.op);
    in.r = 0.0;
    in.idm = 0;
    in.dgm = 0;
    if (do_diag) {
        // stream offsets: 2 draws per identity slot, 1 per diagonal operator, in slot order (Appendix A)
        in.idm = __ballot_sync(FULL, is_id);
        in.dgm = __ballot_sync(FULL, is_dg);
        const uint32_t D = 2u * __popc(in.idm) + __popc(in.dgm);
        if (D) {
            const uint32_t rel = (uint32_t)(draws & 1ull) + 2u * __popc(in.idm & lt) + __popc(in.dgm & lt);  // index into the scratch
            const unsigned long long j0 = draws >> 1, my = 2ull * j0 + rel;
            fill_draws<INJ>(c, j0);
            if (!INJ && (draws & 1ull) && D == 64u && lane == 0)  // the one draw beyond 32 blocks
                philox_extra_block(c.seed, c.wid, j0 + 32, reinterpret_cast<uint4 *>(c.rng) + 32);
            __syncwarp();
            if (is_id) {
                in.bond = sse_uint_below32(scratch_draw<INJ>(c, rel, my), (uint32_t)dm.n_bonds);  // rand(rng, 1:N_b) - 1 (sse.jl:152)
                in.r = sse_u01(scratch_draw<INJ>(c, rel + 1u, my + 1));                                    // sse.jl:166
            } else if (is_dg) {
                in.r = sse_u01(scratch_draw<INJ>(c, rel, my));                                             // sse.jl:178
            }
            draws += D;
            __syncwarp();  // the scratch is refilled for the next chunk
        }
    }
    // the chunk's bond-table rows travel to shared memory while the previous chunk is decided (stage C)
    cp_async16(c.biring_s + 16u * (32u * (uint32_t)(ch & 1) + lane), dm.bond_info + in.bond, is_id || nonid);
    cp_async_commit();
    return in;
}

// MEAS: Carlo.measure! of the sweep that just ended (src/sse.jl:70-87, measure_opstring! :321-376) is evaluated on the way:
// this pass reads, in slot order and with the state propagated along the string, exactly the configuration measure! would
// see, so the separate pass over the string is saved.  Needs the compressed estimator rows
// (st.estrows, at most two estimators); the observables go to meas_out[n_obs].
struct MeasAcc {  // per-lane partial sums of one estimator (magnetization_estimator.jl:152-158)
    double tmpmag, mag, absmag, mag2, mag4;
};

template <bool INJ, bool MEAS = false>
__device__ void phase_diag_build(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool do_diag,
                                 double *meas_out = nullptr) {
    const uint32_t lane = c.lane, lt = lanemask_lt();
    const int N = dm.n_sites;
    // ---- measurement, part 1: everything that refers to the configuration BEFORE this diagonal update ----
    const double meas_nops = (double)c.n;
    const int md = dm.est_max_dim;
    MeasAcc ma[2];
    uint32_t neg = 0;
    if (MEAS) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            ma[g].tmpmag = ma[g].mag = ma[g].absmag = ma[g].mag2 = ma[g].mag4 = 0.0;
            if (g < dm.n_est) {  // init (magnetization_estimator.jl:96-123)
                const double *ev = dm.est_values + (size_t)g * N * md;
                double part = 0.0;
                for (int s = lane; s < N; s += 32) part += __ldg(ev + (size_t)s * md + (c.state[s] - 1));
                ma[g].tmpmag = warp_sum_f64(part);
                if (lane == 0) {
                    ma[g].mag = ma[g].tmpmag;
                    ma[g].absmag = fabs(ma[g].tmpmag);
                    ma[g].mag2 = ma[g].tmpmag * ma[g].tmpmag;
                    ma[g].mag4 = ma[g].mag2 * ma[g].mag2;
                }
            }
        }
    }
    if (do_diag && 2ll * (long long)c.n >= (long long)c.M) {  // n >= 0.5*M  (sse.jl:138)
        long long newM = (3ll * (long long)c.M) / 2 + 100;     // floor(1.5*M + 100) (sse.jl:143)
        if (newM > dw.M_cap) { c.flags |= SSE_FLAG_M_OVERFLOW; return; }
        c.M = (int)newM;  // slots beyond the old M are identity by invariant (sse.jl:144)
    }
    for (int s = lane; s < N; s += 32) { c.vfirst[s] = NONE32; c.vlast[s] = NONE32; }
    __syncwarp();
    const int M = c.M;
    const double p_make_bond_raw = (double)dm.n_bonds / c.T;   // sse.jl:147
    const double p_remove_bond_raw = c.T / (double)dm.n_bonds; // sse.jl:148
    const uint32_t Rcap = c.Rcap, n_old = (uint32_t)c.n, n_cap32 = (uint32_t)dw.n_cap;
    const uint32_t Gn = ring(c.G, Rcap, n_old);  // the new generation starts right behind the old one
    int n = c.n;
    uint32_t kbase = 0, built = 0;  // operators of the new generation so far / already linked
    unsigned long long draws = c.draws;
    const int nchunks = (M + 31) >> 5;
    OpReader<1> rd;
    rd.init(c.words, c.rec, c.G, Rcap, nchunks, lane, c.opring_s);
    uint2 wout = make_uint2(0u, 0u);  // new {bits, rank} of chunk 32*j + lane, written 32 words at a time
    // accept thresholds (see below), valid while the operator count stays inside [win_lo, win_hi]
    int win_lo = 1, win_hi = 0;
    double pm_lo = 0, pm_hi = 0, rm_sure = 0, rm_maybe = 0;

    BuildArgs ba;
    ba.pol = policy_evict_last();
    ba.queue = c.queue;
    ba.mark = c.mark;
    ba.vfirst = c.vfirst;
    ba.vlast = c.vlast;
    ba.rec = c.rec;
    ba.Rcap = Rcap;
    ba.Gn = Gn;
    ba.lane = lane;
    ChunkIn in, nxt;
    in.op = in.bond = in.idm = in.dgm = in.kold0 = 0;
    in.r = 0.0;
    nxt = in;
    for (int ch = -1; ch < nchunks; ++ch) {  // iteration ch: stage B of chunk ch + 1, then stage C of chunk ch
        if (ch + 1 < nchunks) {
            nxt = diag_stage_b<INJ>(dm, c, rd, ch + 1, M, do_diag, draws);
        } else {  // no chunk left to prepare: keep the group pattern of an iteration
            cp_async_commit();
            cp_async_commit();
        }
        if (ch < 0) {
            in = nxt;
            continue;
        }
        // ------------------------------ stage C of chunk ch ------------------------------
        const int p = ch * 32 + (int)lane;
        const bool active = p < M;
        const uint32_t op = in.op;
        const bool nonid = op != 0u;
        const bool is_id = active && !nonid;
        const bool is_dg = nonid && (op & 2u);
        const bool is_off = nonid && !(op & 2u);
        const uint32_t bond = in.bond, gv = op_gv(op), idm = in.idm, dgm = in.dgm;
        const double r = in.r;
        cp_async_wait<2>();  // this chunk's bond rows have landed; the two groups of stage B above may still be in flight
        const uint4 bi = lds128(c.biring_s + 16u * (32u * (uint32_t)(ch & 1) + lane));
        const uint32_t sa = bi.x & NONE24, sb = bi.y & NONE24;
        uint32_t newop = op;

        if (MEAS) {  // ---- measurement, part 2: this chunk's (old) operators ----
            neg += __popc(__ballot_sync(FULL, nonid && st.vneg[gv]));  // measure_sign (sse.jl:305-314)
            const uint32_t moffm = __ballot_sync(FULL, is_off);
            const uint32_t mvi = is_off ? st.vinfo[gv] : 0u;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (g >= dm.n_est) break;
                double scan = 0.0;
                if (is_off) {  // off-diagonal: tmpmag += sum_l sign*(m(top_l) - m(bottom_l)) (:134-150)
                    const double *ea = st.estrows + ((size_t)g * dm.tl.est_nrows + ((bi.w >> (16 * g)) & 0xffu)) * md;
                    const double *eb = st.estrows + ((size_t)g * dm.tl.est_nrows + ((bi.w >> (16 * g + 8)) & 0xffu)) * md;
                    scan = (ea[((mvi >> 16) & 0xffu) - 1] - ea[(mvi & 0xffu) - 1]) + (eb[(mvi >> 24) - 1] - eb[((mvi >> 8) & 0xffu) - 1]);
                }
                if (moffm) {  // inclusive prefix sum over the chunk, in slot order
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double up = shfl_up_f64(scan, d);
                        if ((int)lane >= d) scan += up;
                    }
                }
                if (nonid) {  // every non-identity operator is one sample (:152-158)
                    const double v = ma[g].tmpmag + scan, v2 = v * v;
                    ma[g].mag += v;
                    ma[g].absmag += fabs(v);
                    ma[g].mag2 += v2;
                    ma[g].mag4 += v2 * v2;
                }
                if (moffm) ma[g].tmpmag += shfl_f64(scan, 31);
            }
        }

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
                    const uint32_t rs = diag_resolve_state(offm, lane, is_id, is_off, sa, sb, ta, tb,
                                                           s_a | (s_b << 8) | ((uint32_t)wa << 16) | ((uint32_t)wb << 17));
                    s_a = rs & 0xffu;
                    s_b = (rs >> 8) & 0xffu;
                    wa = (rs >> 16) & 1u;
                    wb = (rs >> 17) & 1u;
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
            // multiplication are monotone), so evaluating them at the two ends of a window that contains this
            // range decides every lane whose draw is not between the two thresholds.  The window (and its two
            // divisions) is kept for as many chunks as n stays inside it.  Only if some lane is undecided
            // (probability ~ 600/(M-n) per chunk) the in-order recurrence is solved exactly by fixed-point iteration.
            const int n_lo = n - __popc(dgm), n_hi = n + __popc(idm);
            if (n_lo < win_lo || n_hi > win_hi) {
                win_lo = n_lo - 256;
                win_hi = n_hi + 256;
                pm_lo = diag_window_make(M, win_lo, p_make_bond_raw);
                pm_hi = diag_window_make(M, win_hi, p_make_bond_raw);
                rm_sure = (double)(M - win_hi + 1) * p_remove_bond_raw;
                rm_maybe = (double)(M - win_lo + 1) * p_remove_bond_raw;
            }
            bool acc = false, amb = false;
            if (is_id) {
                acc = r < pm_lo * w;
                amb = !acc && (r < pm_hi * w);
            } else if (is_dg) {
                const double rw = r * w;
                acc = rw < rm_sure;
                amb = !acc && (rw < rm_maybe);
            }
            uint32_t ins, rem;
            if (__ballot_sync(FULL, amb)) {
                const unsigned long long ir = diag_resolve_exact(n, M, p_make_bond_raw, p_remove_bond_raw, is_id, is_dg, r, w, lt);
                ins = (uint32_t)ir;
                rem = (uint32_t)(ir >> 32);
            } else {
                ins = __ballot_sync(FULL, is_id && acc);
                rem = __ballot_sync(FULL, is_dg && acc);
            }
            n += __popc(ins) - __popc(rem);
            if (is_id && ((ins >> lane) & 1u)) newop = op_pack(bond, gvnew, 1u);
            if (is_dg && ((rem >> lane) & 1u)) newop = 0u;
        }

        // ---- the chunk's operators join the queue of the record build (make_vertex_list!) ----
        const bool nn = newop != 0u;
        const uint32_t nm = __ballot_sync(FULL, nn);
        const uint32_t cnt = __popc(nm);
        // capacity: n_cap records, and the write head must stay clear of old records that are not consumed yet (the
        // reader is up to OP_AHEAD + 2 chunks ahead of this point: ROT_MARGIN covers them)
        if (kbase + cnt > n_cap32 || n_old + kbase + cnt + (uint32_t)ROT_MARGIN > Rcap + in.kold0) {  // all below 2^24
            c.flags |= SSE_FLAG_N_OVERFLOW;
            return;
        }
        if (nn) {
            const uint32_t q = (kbase + __popc(nm & lt)) & (BUILD_QUEUE - 1);
            c.queue[q] = newop;
            c.queue[BUILD_QUEUE + q] = sa;
            c.queue[2 * BUILD_QUEUE + q] = sb;
        }
        if (lane == (uint32_t)(ch & 31)) wout = make_uint2(nm, kbase);
        if ((ch & 31) == 31 || ch == nchunks - 1) {
            const int wi = (ch & ~31) + (int)lane;
            if (wi <= ch) c.words[wi] = wout;
        }
        kbase += cnt;
        in = nxt;
        __syncwarp();
        if (kbase - built >= 32u) {  // 32 operators are waiting: link them with every lane busy
            build_records(ba, built, 32u);
            built += 32u;
        }
    }
    while (built < kbase) {
        const uint32_t m = kbase - built < 32u ? kbase - built : 32u;
        build_records(ba, built, m);
        built += m;
    }
    finish_links(ba, kbase, policy_evict_first(), N);
    if (MEAS) {  // ---- measurement, part 3: the observables (sse.jl:73-82; result, magnetization_estimator.jl:205-230) ----
        const double sign = (neg & 1u) ? -1.0 : 1.0;  // sse.jl:313
        if (lane == 0) {
            meas_out[SSE_OBS_SIGN] = sign;
            meas_out[SSE_OBS_OPERATOR_COUNT] = meas_nops;
            meas_out[SSE_OBS_SIGN_OPERATOR_COUNT] = sign * meas_nops;
            meas_out[SSE_OBS_SIGN_OPERATOR_COUNT2] = sign * (meas_nops * meas_nops);
            meas_out[SSE_OBS_SIGN_ENERGY] = -sign * (meas_nops * c.T + dm.energy_offset) / (double)dm.norm_sites;
            meas_out[SSE_OBS_WORM_LENGTH_FRACTION] = c.last_wlf;
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (g >= dm.n_est) break;
            double m1 = warp_sum_f64(ma[g].mag), mabs = warp_sum_f64(ma[g].absmag), m2 = warp_sum_f64(ma[g].mag2), m4 = warp_sum_f64(ma[g].mag4);
            if (lane == 0) {
                const double ns = 1.0 + meas_nops;
                const double norm = 1.0 / (double)dm.norm_sites;
                m1 *= norm;
                mabs *= norm;
                m2 *= norm * norm;
                m4 *= (norm * norm) * (norm * norm);
                double *o = meas_out + SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * g;
                o[0] = sign * m1 / ns;
                o[1] = sign * mabs / ns;
                o[2] = sign * m2 / ns;
                o[3] = sign * m4 / ns;
                o[4] = sign * (1.0 / c.T / (ns + 1.0) / ns * (m1 * m1 + m2) * (double)dm.norm_sites);
            }
        }
        __syncwarp();
    }
    c.n = n;
    c.G = Gn;
    c.draws = draws;
}

// ------------------------------------------------------------------------------------------------------
// worm_update after the worms (src/sse.jl:200-228): WormLengthFraction, the worm-count controller, and the state
// rebuild from the first leg on each site.  total = 1 + sum of the worm lengths (sse.jl:194-198).
// ------------------------------------------------------------------------------------------------------
template <bool INJ>
__device__ void worm_finish(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, bool thermalized, int widx,
                            double total) {
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
            const uint32_t op = __ldcg(&c.rec[ring(c.G, c.Rcap, f >> 2)].x);
            c.state[s] = (uint8_t)((st.vinfo[op_gv(op)] >> (8u * (f & 3u))) & 0xffu);
        }
        c.draws += __popc(em);
    }
    if (INJ && (long long)c.draws > c.inj_len) c.flags |= SSE_FLAG_STREAM_EXHAUSTED;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// Carlo.measure! (src/sse.jl:70-87): measure_sign (:305-314), the scalar observables, measure_opstring! (:321-376)
// with the table-driven MagnetizationEstimator init/measure/result (magnetization_estimator.jl:96-230).
// out[n_obs] (global) receives the observables.
// ------------------------------------------------------------------------------------------------------
// One pass over the string serves the sign and up to MEASURE_GROUP estimators at once (every pass re-reads all op codes).
constexpr int MEASURE_GROUP = 2;

__device__ void phase_measure(const SmTab &st, const DevModel &dm, const DevWalkers &dw, Ctx &c, double *out) {
    const uint32_t lane = c.lane;
    const int M = c.M;
    const int nchunks = (M + 31) >> 5;
    const int N = dm.n_sites, md = dm.est_max_dim;
    const double nops = (double)c.n;
    double sign = 1.0;
    for (int e0 = 0; e0 == 0 || e0 < dm.n_est; e0 += MEASURE_GROUP) {
        const int ne = (dm.n_est - e0 < MEASURE_GROUP) ? dm.n_est - e0 : MEASURE_GROUP;  // may be 0: the sign alone
        const double *ev[MEASURE_GROUP];
        double tmpmag[MEASURE_GROUP], mag[MEASURE_GROUP], absmag[MEASURE_GROUP], mag2[MEASURE_GROUP], mag4[MEASURE_GROUP];
#pragma unroll
        for (int g = 0; g < MEASURE_GROUP; ++g) {
            ev[g] = dm.est_values + (size_t)(e0 + (g < ne ? g : 0)) * N * md;
            // init (magnetization_estimator.jl:96-123)
            double part = 0.0;
            if (g < ne)
                for (int s = lane; s < N; s += 32) part += __ldg(ev[g] + (size_t)s * md + (c.state[s] - 1));
            tmpmag[g] = warp_sum_f64(part);
            mag[g] = absmag[g] = mag2[g] = mag4[g] = 0.0;  // per-lane partial sums
            if (lane == 0) { mag[g] = tmpmag[g]; absmag[g] = fabs(tmpmag[g]); mag2[g] = tmpmag[g] * tmpmag[g]; mag4[g] = mag2[g] * mag2[g]; }
        }
        uint32_t neg = 0;
        OpReader<0> rd;
        rd.init(c.words, c.rec, c.G, c.Rcap, nchunks, lane, c.opring_s);
        for (int ch = 0; ch < nchunks; ++ch) {
            uint32_t bits, op;
            rd.take(ch, bits, op);
            rd.request();
            cp_async_commit();
            const bool nonid = op != 0u;
            if (e0 == 0) neg += __popc(__ballot_sync(FULL, nonid && st.vneg[op_gv(op)]));  // measure_sign (sse.jl:305-314)
            if (ne == 0) continue;
            const bool off = nonid && !(op & 2u);
            const uint32_t offm = __ballot_sync(FULL, off);
            uint4 bi = make_uint4(0, 0, 0, 0);
            uint32_t vi = 0;
            if (off) {
                bi = __ldg(dm.bond_info + op_bond(op));
                vi = st.vinfo[op_gv(op)];
            }
#pragma unroll
            for (int g = 0; g < MEASURE_GROUP; ++g) {
                if (g >= ne) break;
                double scan = 0.0;
                if (off) {  // off-diagonal: tmpmag += sum_l sign*(m(top_l) - m(bottom_l)) (:134-150)
                    const double *ea = ev[g] + (size_t)(bi.x & NONE24) * md, *eb = ev[g] + (size_t)(bi.y & NONE24) * md;
                    scan = (__ldg(ea + ((vi >> 16) & 0xffu) - 1) - __ldg(ea + (vi & 0xffu) - 1)) +
                           (__ldg(eb + (vi >> 24) - 1) - __ldg(eb + ((vi >> 8) & 0xffu) - 1));
                }
                if (offm) {  // inclusive prefix sum over the chunk, in slot order
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double up = shfl_up_f64(scan, d);
                        if ((int)lane >= d) scan += up;
                    }
                }
                if (nonid) {  // every non-identity operator is one sample (:152-158)
                    const double v = tmpmag[g] + scan, v2 = v * v;
                    mag[g] += v;
                    absmag[g] += fabs(v);
                    mag2[g] += v2;
                    mag4[g] += v2 * v2;
                }
                if (offm) tmpmag[g] += shfl_f64(scan, 31);
            }
        }
        if (e0 == 0) {
            sign = (neg & 1u) ? -1.0 : 1.0;  // sse.jl:313
            if (lane == 0) {
                out[SSE_OBS_SIGN] = sign;
                out[SSE_OBS_OPERATOR_COUNT] = nops;
                out[SSE_OBS_SIGN_OPERATOR_COUNT] = sign * nops;
                out[SSE_OBS_SIGN_OPERATOR_COUNT2] = sign * (nops * nops);
                out[SSE_OBS_SIGN_ENERGY] = -sign * (nops * c.T + dm.energy_offset) / (double)dm.norm_sites;
                out[SSE_OBS_WORM_LENGTH_FRACTION] = c.last_wlf;
            }
        }
#pragma unroll
        for (int g = 0; g < MEASURE_GROUP; ++g) {
            if (g >= ne) break;
            double m1 = warp_sum_f64(mag[g]), ma = warp_sum_f64(absmag[g]), m2 = warp_sum_f64(mag2[g]), m4 = warp_sum_f64(mag4[g]);
            if (lane == 0) {  // result (:205-230)
                const double ns = 1.0 + nops;
                const double norm = 1.0 / (double)dm.norm_sites;
                m1 *= norm;
                ma *= norm;
                m2 *= norm * norm;
                m4 *= (norm * norm) * (norm * norm);
                double *o = out + SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * (e0 + g);
                o[0] = sign * m1 / ns;
                o[1] = sign * ma / ns;
                o[2] = sign * m2 / ns;
                o[3] = sign * m4 / ns;
                o[4] = sign * (1.0 / c.T / (ns + 1.0) / ns * (m1 * m1 + m2) * (double)dm.norm_sites);
            }
        }
    }
    __syncwarp();
}

// Add a walker's freshly measured observables to its accumulators (one sample of the bin).
__device__ __forceinline__ void accumulate_obs(const DevWalkers &dw, int w, uint32_t lane, const double *out) {
    __syncwarp();
    for (int i = lane; i < dw.n_obs; i += 32)
        if (i != SSE_OBS_WORM_LENGTH_FRACTION) dw.acc[(size_t)w * dw.n_obs + i] += out[i];
    if (lane == 0) dw.acc_cnt[2 * w] += 1;
    __syncwarp();
}

}  // namespace sse
