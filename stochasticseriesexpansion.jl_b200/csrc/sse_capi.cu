// sse_capi.cu — C ABI of libsse_b200.so (include/sse_b200.h): table flattening to the device image,
// device memory ownership, launches of the one walker kernel (sse_kernels.cuh), checkpoint conversion
// between the device op codes and the reference's UInt64 OperCode layout (src/opercode.jl:43-47).
// No CPU fallback exists: every entry point runs on the GPU or returns an error status.
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "sse_kernels.cuh"
#include "sse_kernels_multi.cuh"

using namespace sse;

// Kernel launches go through one macro so that the test-only warp emulator (tests/emu) can compile this file with g++.
#ifndef SSE_LAUNCH_KERNEL
#define SSE_LAUNCH_KERNEL(kern, grid, block, smem_bytes, stream, ...) kern<<<grid, block, smem_bytes, stream>>>(__VA_ARGS__)
#endif

namespace {

thread_local std::string g_err;

int32_t fail(const std::string &msg) {
    g_err = msg;
    return 1;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                         \
    } while (0)

template <class T>
cudaError_t upload(T **dst, const std::vector<T> &v) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

}  // namespace

struct sse_model {
    int device = 0;
    DevModel dm{};
    // host copies for checkpoint conversion and validation
    std::vector<int32_t> bond_type, bond_sites, type_vertex_off;
    std::vector<uint8_t> is_diag;  // per global vertex
    std::vector<uint8_t> site_dim;
    int n_types = 0;
    // device allocations
    uint4 *d_bond_info = nullptr;
    uint8_t *d_site_dim = nullptr, *d_blob = nullptr;
    double *d_est = nullptr;
};

struct sse_walkers {
    const sse_model *model = nullptr;
    DevWalkers dw{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int chains = 1;           // walkers per warp in sse_sweep launches (1 = sse::k_walkers, 2/4 = sse::k_walkers_multi)
    bool indexed = false;     // string currently in indexed mode (after make_vertex_list, before commit)
    bool have_vl = false;
    unsigned long long *d_inj = nullptr;
    int64_t bytes = 0;
    std::vector<void *> allocs;
};

namespace {

int32_t check_flags(sse_walkers *w) {
    std::vector<uint32_t> f(w->dw.W);
    CU(cudaMemcpyAsync(f.data(), w->dw.flags, sizeof(uint32_t) * f.size(), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < w->dw.W; ++i) {
        if (f[i] & SSE_FLAG_M_OVERFLOW)
            return fail("walker " + std::to_string(i) + ": operator string outgrew m_capacity (recreate with a larger m_capacity)");
        if (f[i] & SSE_FLAG_N_OVERFLOW)
            return fail("walker " + std::to_string(i) + ": more operators than n_capacity (recreate with a larger n_capacity)");
        if (f[i] & SSE_FLAG_STREAM_EXHAUSTED)
            return fail("walker " + std::to_string(i) + ": injected random stream exhausted");
    }
    return 0;
}

// walkers per warp -> resident CTAs per SM the multi-chain kernel is compiled for (register budget per thread).
// Tuning only: the environment variable SSE_B200_MULTI_MINB selects one of the other compiled occupancies
// (2 walkers per warp: 7 (default) or 5 CTAs/SM = 72 or 96 registers; 4 per warp: 5 (default), 4 or 3 CTAs/SM = 96, 128 or
// 168 registers; the chase loop has no spills in any of them).
constexpr int MULTI2_MINB = 7, MULTI4_MINB = 5;

int multi_minb(int ch) {
    int minb = ch == 2 ? MULTI2_MINB : MULTI4_MINB;
    if (const char *e = getenv("SSE_B200_MULTI_MINB")) minb = atoi(e);
    return minb;
}

// highest shared-memory level of the multi-chain kernel whose CTA still fits MINB times into an SM (227 KB, 1 KB
// reserved per CTA); level 0 (rng scratch only) always fits
int multi_level(const sse_model *m, int ch, int minb) {
    const int budget = 227 * 1024 / minb - 1024;
    for (int level = 2; level >= 1; --level)
        if (m->dm.tl.bytes + WARPS_PER_CTA * multi_warp_bytes(m->dm.n_sites, level, ch) <= budget) return level;
    return 0;
}

typedef void (*multi_kernel_t)(const DevModel, const DevWalkers, const LaunchArgs);

template <bool INJ>
multi_kernel_t multi_kernel(int ch, int minb) {
    if (ch == 2 && minb == 7) return k_walkers_multi<INJ, 2, 7>;
    if (ch == 2 && minb == 5) return k_walkers_multi<INJ, 2, 5>;
    if (ch == 4 && minb == 4) return k_walkers_multi<INJ, 4, 4>;
    if (ch == 4 && minb == 5) return k_walkers_multi<INJ, 4, 5>;
    if (ch == 4 && minb == 3) return k_walkers_multi<INJ, 4, 3>;
    return nullptr;
}

template <bool INJ>
int32_t launch_multi(sse_walkers *w, const LaunchArgs &a) {
    const sse_model *m = w->model;
    const int ch = w->chains, minb = multi_minb(ch);
    multi_kernel_t kern = multi_kernel<INJ>(ch, minb);
    if (!kern) return fail("SSE_B200_MULTI_MINB: no kernel compiled for " + std::to_string(ch) + " walkers per warp at " +
                           std::to_string(minb) + " CTAs per SM (available: 2 -> 7, 5;  4 -> 4, 5, 3)");
    DevWalkers dw = w->dw;
    dw.smem_state = multi_level(m, ch, minb);
    if (const char *lv = getenv("SSE_B200_SMEM_LEVEL")) dw.smem_state = std::min(dw.smem_state, std::max(0, atoi(lv)));
    if (!dw.smem_state && !dw.mark) return fail("internal: mark[] scratch missing for the multi-chain kernel");
    const int per_cta = WARPS_PER_CTA * ch;
    const int grid = (dw.W + per_cta - 1) / per_cta;
    const int block = WARPS_PER_CTA * 32;
    const size_t smem = (size_t)m->dm.tl.bytes + (size_t)WARPS_PER_CTA * multi_warp_bytes(m->dm.n_sites, dw.smem_state, ch);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSE_LAUNCH_KERNEL(kern, grid, block, smem, w->stream, m->dm, dw, a);
    CU(cudaGetLastError());
    return 0;
}

int32_t launch(sse_walkers *w, const LaunchArgs &a) {
    const sse_model *m = w->model;
    CU(cudaSetDevice(m->device));
    if (a.mode == MODE_SWEEP && w->chains > 1) return w->dw.inj ? launch_multi<true>(w, a) : launch_multi<false>(w, a);
    const int grid = (w->dw.W + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int block = WARPS_PER_CTA * 32;
    const size_t smem = (size_t)m->dm.tl.bytes + (size_t)WARPS_PER_CTA * warp_scratch_bytes(m->dm.n_sites, w->dw.smem_state);
    if (smem > 48 * 1024) {
        CU(cudaFuncSetAttribute(k_walkers<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_walkers<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (w->dw.inj)
        SSE_LAUNCH_KERNEL(k_walkers<true>, grid, block, smem, w->stream, m->dm, w->dw, a);
    else
        SSE_LAUNCH_KERNEL(k_walkers<false>, grid, block, smem, w->stream, m->dm, w->dw, a);
    CU(cudaGetLastError());
    return 0;
}

template <class T>
int32_t dev_alloc(sse_walkers *w, T **p, size_t count, bool zero) {
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CU(cudaMalloc((void **)p, bytes));
    w->allocs.push_back(*p);
    w->bytes += (int64_t)bytes;
    if (zero) CU(cudaMemset(*p, 0, bytes));
    return 0;
}

int32_t ensure_committed(sse_walkers *w) {
    if (w->indexed) {
        LaunchArgs a{};
        a.mode = MODE_COMMIT;
        if (int32_t s = launch(w, a)) return s;
        w->indexed = false;
    }
    return 0;
}

}  // namespace

extern "C" {

int32_t sse_model_destroy(sse_model *m);

const char *sse_last_error(void) { return g_err.c_str(); }
int32_t sse_abi_version(void) { return SSE_B200_ABI_VERSION; }

int32_t sse_model_create(const sse_model_desc *d, sse_model **out) {
    if (!d || !out) return fail("sse_model_create: null argument");
    if (d->n_sites <= 0 || d->n_bonds <= 0 || d->n_types <= 0 || d->n_vertices <= 0)
        return fail("sse_model_create: empty model");
    if (d->n_vertices >= (1 << VBITS)) return fail("sse_model_create: more than 4095 vertices in total is not supported");
    if (d->n_bonds >= (1 << (32 - BOND_SHIFT))) return fail("sse_model_create: more than 262143 bonds is not supported");
    if (d->n_sites >= (1 << 24)) return fail("sse_model_create: more than 2^24 sites is not supported");
    if (d->max_worm < 1 || d->max_worm > 254) return fail("sse_model_create: max_worm out of range");
    std::unique_ptr<sse_model, int32_t (*)(sse_model *)> guard(new sse_model(), sse_model_destroy);  // freed on every error path
    sse_model *m = guard.get();
    CU(cudaGetDevice(&m->device));
    m->n_types = d->n_types;
    m->bond_type.assign(d->bond_type, d->bond_type + d->n_bonds);
    m->bond_sites.assign(d->bond_sites, d->bond_sites + 2 * d->n_bonds);
    m->type_vertex_off.assign(d->type_vertex_off, d->type_vertex_off + d->n_types + 1);
    m->site_dim.assign(d->site_dim, d->site_dim + d->n_sites);
    const int nv = d->n_vertices;
    m->is_diag.resize(nv);
    std::vector<int> vtype(nv);
    for (int t = 0; t < d->n_types; ++t)
        for (int v = d->type_vertex_off[t]; v < d->type_vertex_off[t + 1]; ++v) vtype[v] = t;
    for (int v = 0; v < nv; ++v) {
        const uint8_t *ls = d->leg_states + 4 * v;
        m->is_diag[v] = (ls[0] == ls[2] && ls[1] == ls[3]);
    }
    // bond table
    std::vector<uint4> bi(d->n_bonds);
    for (int b = 0; b < d->n_bonds; ++b) {
        int t = d->bond_type[b];
        int sa = d->bond_sites[2 * b], sb = d->bond_sites[2 * b + 1];
        if (t < 0 || t >= d->n_types || sa < 0 || sb < 0 || sa >= d->n_sites || sb >= d->n_sites) {
            return fail("sse_model_create: bond " + std::to_string(b) + " out of range");
        }
        if (sa == sb) { return fail("sse_model_create: bonds connecting a site to itself are not supported"); }
        int da = d->type_dims[2 * t], db = d->type_dims[2 * t + 1];
        if (da != d->site_dim[sa] || db != d->site_dim[sb]) {
            return fail("SSEData: site dimensions set by VertexData are inconsistent (bond " + std::to_string(b) + ")");
        }
        bi[b] = make_uint4((uint32_t)sa | ((uint32_t)da << 24), (uint32_t)sb | ((uint32_t)db << 24),
                           (uint32_t)d->type_diag_off[t], (uint32_t)t);
    }
    // shared-memory image
    const int n_out = d->n_outcomes, n_diag = d->type_diag_off[d->n_types];
    const int n_trans = nv * d->max_worm * 4;
    TabLayout tl{};
    int off = 0;
    auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
    tl.off_t1 = take(16 * n_trans);
    tl.off_outc = take(16 * n_out);
    tl.off_weights = take(8 * nv);
    tl.off_vinfo = take(4 * nv);
    tl.off_diagv = take(2 * n_diag);
    tl.off_vneg = take(nv);
    tl.bytes = off;
    if (tl.bytes > 64 * 1024) { return fail("sse_model_create: vertex tables exceed the 64 KB shared-memory budget"); }
    std::vector<uint8_t> blob(tl.bytes, 0);
    auto *outc = reinterpret_cast<uint4 *>(blob.data() + tl.off_outc);
    auto *wts = reinterpret_cast<double *>(blob.data() + tl.off_weights);
    auto *t1 = reinterpret_cast<uint4 *>(blob.data() + tl.off_t1);
    auto *vinfo = reinterpret_cast<uint32_t *>(blob.data() + tl.off_vinfo);
    auto *diagv = reinterpret_cast<uint16_t *>(blob.data() + tl.off_diagv);
    auto *vneg = blob.data() + tl.off_vneg;
    for (int v = 0; v < nv; ++v) {
        const uint8_t *ls = d->leg_states + 4 * v;
        wts[v] = d->weights[v];
        vinfo[v] = (uint32_t)ls[0] | ((uint32_t)ls[1] << 8) | ((uint32_t)ls[2] << 16) | ((uint32_t)ls[3] << 24);
        vneg[v] = d->signs[v] < 0;
    }
    for (int t = 0; t < d->n_types; ++t)
        for (int c = d->type_diag_off[t]; c < d->type_diag_off[t + 1]; ++c) {
            int lv = d->diag_vertices[c];
            diagv[c] = lv ? (uint16_t)(d->type_vertex_off[t] + lv - 1 + 1) : 0;
        }
    // outcomes: which type an outcome belongs to follows from the transition that references it
    std::vector<int> otype(n_out, -1);
    for (int i = 0; i < n_trans; ++i) {
        int o = d->trans_offset[i], c = d->trans_count[i];
        if (o < 0) continue;
        if (c < 1 || c > 64 || o + c > n_out || o + 1 >= (1 << 18)) { return fail("sse_model_create: bad transition entry (at most 64 outcomes per transition, 2^18 outcomes in total)"); }
        int v = i / (d->max_worm * 4);
        for (int j = 0; j < c; ++j) otype[o + j] = vtype[v];
    }
    for (int o = 0; o < n_out; ++o) {
        int t = otype[o];
        uint32_t pk = 0, pw = 0;
        if (t >= 0) {
            int tv = d->out_target[o];
            int leg = d->out_leg[o], worm = d->out_worm[o];
            int gv = d->type_vertex_off[t] + tv - 1;
            if (tv < 1 || gv >= d->type_vertex_off[t + 1] || leg < 0 || leg > 3 || worm < 1 || worm > 254) {
                return fail("sse_model_create: bad outcome entry");
            }
            int dim_out = d->type_dims[2 * t + (leg & 1)];
            pk = ((uint32_t)m->is_diag[gv] << 1) | ((uint32_t)gv << 2) | ((uint32_t)leg << 16) | ((uint32_t)worm << 24);
            pw = (uint32_t)dim_out << 24;
        }
        uint64_t bits;
        double cp = d->out_cumprob[o];
        memcpy(&bits, &cp, 8);
        outc[o] = make_uint4((uint32_t)bits, (uint32_t)(bits >> 32), pk, pw);
    }
    // transition header fused with its first outcome; invalid transitions get cumprob +inf and step 0
    for (int i = 0; i < n_trans; ++i) {
        int o = d->trans_offset[i], c = d->trans_count[i];
        if (o < 0) { t1[i] = make_uint4(0u, 0x7ff00000u, 0u, 0u); continue; }
        t1[i] = make_uint4(outc[o].x, outc[o].y, outc[o].z, outc[o].w | ((uint32_t)(o + 1) << 6) | (uint32_t)(c - 1));
    }
    // most likely exit leg per entrance leg, marginalised over vertices (weighted by their weight) and worms
    uint32_t pred_exit = 0;
    for (int leg = 0; leg < 4; ++leg) {
        double score[4] = {0, 0, 0, 0};
        for (int v = 0; v < nv; ++v)
            for (int wm = 0; wm < d->max_worm; ++wm) {
                int i = (v * d->max_worm + wm) * 4 + leg;
                int o = d->trans_offset[i], c = d->trans_count[i];
                if (o < 0) continue;
                double prev = 0;
                for (int j = 0; j < c; ++j) {
                    score[d->out_leg[o + j]] += d->weights[v] * (d->out_cumprob[o + j] - prev);
                    prev = d->out_cumprob[o + j];
                }
            }
        int best = 0;
        for (int j = 1; j < 4; ++j)
            if (score[j] > score[best]) best = j;
        pred_exit |= (uint32_t)best << (2 * leg);
    }
    CU(upload(&m->d_bond_info, bi));
    CU(upload(&m->d_site_dim, m->site_dim));
    CU(upload(&m->d_blob, blob));
    std::vector<double> est;
    if (d->n_estimators > 0) est.assign(d->est_values, d->est_values + (size_t)d->n_estimators * d->n_sites * d->est_max_dim);
    CU(upload(&m->d_est, est));
    DevModel &dm = m->dm;
    dm.n_sites = d->n_sites;
    dm.n_bonds = d->n_bonds;
    dm.nv = nv;
    dm.max_worm = d->max_worm;
    dm.n_est = d->n_estimators;
    dm.est_max_dim = d->est_max_dim > 0 ? d->est_max_dim : 1;
    dm.norm_sites = d->norm_site_count;
    dm.energy_offset = d->energy_offset;
    dm.bond_info = m->d_bond_info;
    dm.site_dim = m->d_site_dim;
    dm.est_values = m->d_est;
    dm.tab_blob = m->d_blob;
    dm.pred_exit = pred_exit;
    dm.variant = getenv("SSE_B200_VARIANT") ? (uint32_t)atoi(getenv("SSE_B200_VARIANT")) : 0u;
    dm.tl = tl;
    *out = guard.release();
    return 0;
}

int32_t sse_model_destroy(sse_model *m) {
    if (!m) return 0;
    cudaFree(m->d_bond_info);
    cudaFree(m->d_site_dim);
    cudaFree(m->d_blob);
    cudaFree(m->d_est);
    delete m;
    return 0;
}

int32_t sse_walkers_create(const sse_model *m, const sse_walkers_opts *o, sse_walkers **out) {
    if (!m || !o || !out) return fail("sse_walkers_create: null argument");
    if (o->n_walkers <= 0) return fail("sse_walkers_create: n_walkers must be positive");
    if (o->m_capacity < 128 || o->m_capacity >= (1ll << 31)) return fail("sse_walkers_create: m_capacity out of range [128, 2^31)");
    if (o->n_capacity < 64 || o->n_capacity > (1ll << 22)) return fail("sse_walkers_create: n_capacity out of range [64, 2^22]");
    if (o->device >= 0 && o->device != m->device) return fail("sse_walkers_create: model was created on another device");
    CU(cudaSetDevice(m->device));
    std::unique_ptr<sse_walkers, int32_t (*)(sse_walkers *)> guard(new sse_walkers(), sse_walkers_destroy);  // freed on every error path
    sse_walkers *w = guard.get();
    w->model = m;
    DevWalkers &dw = w->dw;
    const int W = o->n_walkers, N = m->dm.n_sites;
    dw.W = W;
    dw.M_cap = (o->m_capacity + 31) & ~31ll;
    dw.n_cap = o->n_capacity;
    dw.n_obs = SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * m->dm.n_est;
    int32_t s = 0;
    s |= dev_alloc(w, &dw.ops, (size_t)W * dw.M_cap, true);
    s |= dev_alloc(w, &dw.rec, 2 * (size_t)W * dw.n_cap, false);  // 32-byte records
    s |= dev_alloc(w, &dw.state, (size_t)W * N, false);
    // per-warp state[N] + mark[N] live in shared memory when 7 CTAs/SM still fit, else in global scratch
    // level 2 (+ vlast) only while 7 CTAs/SM still fit (32 KB per CTA); level 1 up to 99 KB per CTA
    dw.smem_state = 0;
    if (m->dm.tl.bytes + WARPS_PER_CTA * warp_scratch_bytes(N, 1) <= 99 * 1024) dw.smem_state = 1;
    if (m->dm.tl.bytes + WARPS_PER_CTA * warp_scratch_bytes(N, 2) <= 32 * 1024) dw.smem_state = 2;
    if (const char *lv = getenv("SSE_B200_SMEM_LEVEL")) dw.smem_state = std::min(dw.smem_state, std::max(0, atoi(lv)));  // tests: force the large-lattice paths
    // global mark[] scratch: needed by whichever kernel (one walker per warp, or 2/4 per warp) runs at level 0
    if (!dw.smem_state || !multi_level(m, 2, 7) || !multi_level(m, 4, 5) || getenv("SSE_B200_SMEM_LEVEL"))
        s |= dev_alloc(w, &dw.mark, (size_t)W * N, true);
    s |= dev_alloc(w, &dw.vfirst, (size_t)W * N, false);
    s |= dev_alloc(w, &dw.vlast, (size_t)W * N, false);
    s |= dev_alloc(w, &dw.T, W, true);
    s |= dev_alloc(w, &dw.M, W, true);
    s |= dev_alloc(w, &dw.n, W, true);
    s |= dev_alloc(w, &dw.num_worms, W, true);
    s |= dev_alloc(w, &dw.avg_wl, W, true);
    s |= dev_alloc(w, &dw.last_wlf, W, true);
    s |= dev_alloc(w, &dw.draws, W, true);
    s |= dev_alloc(w, &dw.flags, W, true);
    s |= dev_alloc(w, &dw.acc, (size_t)W * dw.n_obs, true);
    s |= dev_alloc(w, &dw.acc_cnt, (size_t)W * 2, true);
    s |= dev_alloc(w, &dw.counters, 8, true);
    s |= dev_alloc(w, &dw.dbg_len, W, true);
    s |= dev_alloc(w, &dw.obs_out, (size_t)W * dw.n_obs, true);
    if (s) return 1;
    dw.inj = nullptr;
    dw.inj_len = 0;
    dw.seed = o->seed;
    dw.wid_off = o->walker_id_offset;
    dw.twlf = o->target_worm_length_fraction;
    dw.atten = o->num_worms_attenuation_factor;
    std::vector<double> T(o->T, o->T + W), nw(W, o->init_num_worms), awl(W, 1.0), wlf(W, NAN);
    for (double t : T)
        if (!(t > 0)) return fail("sse_walkers_create: temperatures must be positive");
    CU(cudaMemcpy(dw.T, T.data(), sizeof(double) * W, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.num_worms, nw.data(), sizeof(double) * W, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.avg_wl, awl.data(), sizeof(double) * W, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dw.last_wlf, wlf.data(), sizeof(double) * W, cudaMemcpyHostToDevice));
    CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    w->own_stream = true;
    if (const char *ch = getenv("SSE_B200_CHAINS")) {
        if (sse_set_walkers_per_warp(w, atoi(ch))) return 1;
    }
    *out = guard.release();
    return 0;
}

int32_t sse_set_walkers_per_warp(sse_walkers *w, int32_t walkers_per_warp) {
    if (!w) return fail("null handle");
    if (walkers_per_warp != 1 && walkers_per_warp != 2 && walkers_per_warp != 4)
        return fail("sse_set_walkers_per_warp: supported values are 1, 2 and 4");
    CU(cudaStreamSynchronize(w->stream));
    w->chains = walkers_per_warp;
    return 0;
}

int32_t sse_walkers_destroy(sse_walkers *w) {
    if (!w) return 0;
    cudaSetDevice(w->model->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    for (void *p : w->allocs) cudaFree(p);
    if (w->d_inj) cudaFree(w->d_inj);
    if (w->own_stream && w->stream) cudaStreamDestroy(w->stream);
    delete w;
    return 0;
}

int32_t sse_set_stream(sse_walkers *w, void *cuda_stream) {
    if (!w) return fail("null handle");
    CU(cudaStreamSynchronize(w->stream));
    if (w->own_stream) { cudaStreamDestroy(w->stream); w->own_stream = false; }
    w->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int32_t sse_n_observables(const sse_walkers *w) { return w ? w->dw.n_obs : -1; }
int64_t sse_device_bytes(const sse_walkers *w) { return w ? w->bytes : -1; }

int32_t sse_init(sse_walkers *w, int64_t init_opstring_cutoff, int32_t diagonal_warmup_sweeps) {
    if (!w) return fail("null handle");
    const int W = w->dw.W;
    CU(cudaSetDevice(w->model->device));
    std::vector<double> T(W);
    CU(cudaMemcpy(T.data(), w->dw.T, sizeof(double) * W, cudaMemcpyDeviceToHost));
    std::vector<int> M(W), n(W, 0);
    for (int i = 0; i < W; ++i) {
        // round(Int, length(sites) * T) (src/sse.jl:51): round-half-even
        long long m0 = init_opstring_cutoff >= 0 ? init_opstring_cutoff : (long long)std::nearbyint((double)w->model->dm.n_sites * T[i]);
        if (m0 > w->dw.M_cap) return fail("sse_init: init_opstring_cutoff exceeds m_capacity");
        M[i] = (int)m0;
    }
    CU(cudaMemsetAsync(w->dw.ops, 0, sizeof(uint32_t) * (size_t)W * w->dw.M_cap, w->stream));
    CU(cudaMemcpyAsync(w->dw.M, M.data(), sizeof(int) * W, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(w->dw.n, n.data(), sizeof(int) * W, cudaMemcpyHostToDevice, w->stream));
    w->indexed = false;
    w->have_vl = false;
    LaunchArgs a{};
    a.mode = MODE_INIT;
    a.warmup = diagonal_warmup_sweeps;
    if (int32_t s = launch(w, a)) return s;
    return check_flags(w);
}

int32_t sse_sweep(sse_walkers *w, int32_t n_sweeps, int32_t thermalized, int32_t measure) {
    if (!w) return fail("null handle");
    if (n_sweeps <= 0) return 0;
    if (int32_t s = ensure_committed(w)) return s;
    LaunchArgs a{};
    a.mode = MODE_SWEEP;
    a.n_sweeps = n_sweeps;
    a.thermalized = thermalized;
    a.measure = measure;
    w->have_vl = false;
    return launch(w, a);
}

int32_t sse_sync(sse_walkers *w) {
    if (!w) return fail("null handle");
    CU(cudaStreamSynchronize(w->stream));
    return check_flags(w);
}

int32_t sse_measure(sse_walkers *w, double *out) {
    if (!w || !out) return fail("null argument");
    LaunchArgs a{};
    a.mode = MODE_MEASURE;
    a.indexed = w->indexed;
    if (int32_t s = launch(w, a)) return s;
    w->indexed = false;  // MODE_MEASURE commits
    CU(cudaMemcpyAsync(out, w->dw.obs_out, sizeof(double) * (size_t)w->dw.W * w->dw.n_obs, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return check_flags(w);
}

int32_t sse_fetch_accumulators(sse_walkers *w, double *sums, int64_t *counts, int32_t reset) {
    if (!w || !sums || !counts) return fail("null argument");
    const size_t W = w->dw.W;
    CU(cudaMemcpyAsync(sums, w->dw.acc, sizeof(double) * W * w->dw.n_obs, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaMemcpyAsync(counts, w->dw.acc_cnt, sizeof(long long) * W * 2, cudaMemcpyDeviceToHost, w->stream));
    if (reset) {
        CU(cudaMemsetAsync(w->dw.acc, 0, sizeof(double) * W * w->dw.n_obs, w->stream));
        CU(cudaMemsetAsync(w->dw.acc_cnt, 0, sizeof(long long) * W * 2, w->stream));
    }
    CU(cudaStreamSynchronize(w->stream));
    return check_flags(w);
}

int32_t sse_accumulators_device_ptr(sse_walkers *w, void **sums, void **counts) {
    if (!w) return fail("null handle");
    if (sums) *sums = w->dw.acc;
    if (counts) *counts = w->dw.acc_cnt;
    return 0;
}

int32_t sse_fetch_counters(sse_walkers *w, uint64_t out[8], int32_t reset) {
    if (!w || !out) return fail("null argument");
    CU(cudaMemcpyAsync(out, w->dw.counters, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, w->stream));
    if (reset) CU(cudaMemsetAsync(w->dw.counters, 0, 8 * sizeof(uint64_t), w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

int32_t sse_get_state(sse_walkers *w, int32_t i, sse_walker_state *st) {
    if (!w || !st) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("sse_get_state: walker index out of range");
    if (int32_t s = ensure_committed(w)) return s;
    const sse_model *m = w->model;
    int M = 0, n = 0;
    unsigned long long draws = 0;
    CU(cudaStreamSynchronize(w->stream));
    CU(cudaMemcpy(&M, w->dw.M + i, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&n, w->dw.n + i, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&draws, w->dw.draws + i, sizeof(draws), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&st->avg_worm_length, w->dw.avg_wl + i, sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&st->num_worms, w->dw.num_worms + i, sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&st->T, w->dw.T + i, sizeof(double), cudaMemcpyDeviceToHost));
    st->num_operators = n;
    st->rng_draws = draws;
    if (st->operators) {
        if (st->operators_len < M) { st->operators_len = M; return fail("sse_get_state: operators buffer too small"); }
        std::vector<uint32_t> ops(M);
        if (M) CU(cudaMemcpy(ops.data(), w->dw.ops + (size_t)i * w->dw.M_cap, sizeof(uint32_t) * M, cudaMemcpyDeviceToHost));
        for (int p = 0; p < M; ++p) {
            uint32_t op = ops[p];
            if (!op) { st->operators[p] = 0; continue; }
            uint32_t bond = op_bond(op), gv = op_gv(op);
            uint64_t lv = (uint64_t)(gv - m->type_vertex_off[m->bond_type[bond]] + 1);
            uint64_t vcode = ((op >> 1) & 1u) | (lv << 1);                // VertexCode(diagonal, idx) (opercode.jl:18-21)
            st->operators[p] = 1ull | (vcode << 1) | ((uint64_t)(bond + 1) << 26);  // OperCode(bond, vertex) (opercode.jl:43-47)
        }
    }
    st->operators_len = M;
    if (st->state) CU(cudaMemcpy(st->state, w->dw.state + (size_t)i * m->dm.n_sites, m->dm.n_sites, cudaMemcpyDeviceToHost));
    return 0;
}

int32_t sse_set_state(sse_walkers *w, int32_t i, const sse_walker_state *st) {
    if (!w || !st || !st->operators || !st->state) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("sse_set_state: walker index out of range");
    if (int32_t s = ensure_committed(w)) return s;
    const sse_model *m = w->model;
    const long long M = st->operators_len;
    if (M < 0 || M > w->dw.M_cap) return fail("sse_set_state: operator string longer than m_capacity");
    std::vector<uint32_t> ops((size_t)w->dw.M_cap, 0u);
    long long n = 0;
    for (long long p = 0; p < M; ++p) {
        uint64_t code = st->operators[p];
        if (code == 0) continue;
        long long bond = (long long)(code >> 26) - 1;
        uint64_t vcode = (code & ((1ull << 25) - 1)) >> 1;  // get_vertex (opercode.jl:61-62)
        long long lv = (long long)(vcode >> 1);
        if (bond < 0 || bond >= m->dm.n_bonds) return fail("sse_set_state: bond index out of range at slot " + std::to_string(p));
        int t = m->bond_type[bond];
        long long gv = m->type_vertex_off[t] + lv - 1;
        if (lv < 1 || gv >= m->type_vertex_off[t + 1]) return fail("sse_set_state: vertex index out of range at slot " + std::to_string(p));
        if ((uint32_t)(vcode & 1) != m->is_diag[gv]) return fail("sse_set_state: diagonal flag inconsistent with the vertex table at slot " + std::to_string(p));
        ops[p] = op_pack((uint32_t)bond, (uint32_t)gv, (uint32_t)(vcode & 1));
        ++n;
    }
    if (n != st->num_operators) return fail("sse_set_state: num_operators does not match the operator string");
    for (int s = 0; s < m->dm.n_sites; ++s)
        if (st->state[s] < 1 || st->state[s] > m->site_dim[s]) return fail("sse_set_state: state index out of range at site " + std::to_string(s));
    CU(cudaStreamSynchronize(w->stream));
    int Mi = (int)M, ni = (int)n;
    unsigned long long draws = st->rng_draws;
    uint32_t zero = 0;
    CU(cudaMemcpy(w->dw.ops + (size_t)i * w->dw.M_cap, ops.data(), sizeof(uint32_t) * ops.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.state + (size_t)i * m->dm.n_sites, st->state, m->dm.n_sites, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.M + i, &Mi, sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.n + i, &ni, sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.draws + i, &draws, sizeof(draws), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.avg_wl + i, &st->avg_worm_length, sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.num_worms + i, &st->num_worms, sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.T + i, &st->T, sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(w->dw.flags + i, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice));
    w->have_vl = false;
    return 0;
}

int32_t sse_get_flags(sse_walkers *w, uint32_t *flags) {
    if (!w || !flags) return fail("null argument");
    CU(cudaStreamSynchronize(w->stream));
    CU(cudaMemcpy(flags, w->dw.flags, sizeof(uint32_t) * w->dw.W, cudaMemcpyDeviceToHost));
    return 0;
}

int32_t sse_get_num_operators(sse_walkers *w, int64_t *out) {
    if (!w || !out) return fail("null argument");
    std::vector<int> n(w->dw.W);
    CU(cudaStreamSynchronize(w->stream));
    CU(cudaMemcpy(n.data(), w->dw.n, sizeof(int) * n.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n.size(); ++i) out[i] = n[i];
    return 0;
}

int32_t sse_pt_log_weight_ratio(sse_walkers *w, const double *new_T, double *out) {
    if (!w || !new_T || !out) return fail("null argument");
    const int W = w->dw.W;
    std::vector<int> n(W);
    std::vector<double> T(W);
    CU(cudaStreamSynchronize(w->stream));
    CU(cudaMemcpy(n.data(), w->dw.n, sizeof(int) * W, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(T.data(), w->dw.T, sizeof(double) * W, cudaMemcpyDeviceToHost));
    for (int i = 0; i < W; ++i) out[i] = -(double)n[i] * std::log(new_T[i] / T[i]);  // src/sse.jl:395
    return 0;
}

int32_t sse_set_temperature(sse_walkers *w, const double *T) {
    if (!w || !T) return fail("null argument");
    for (int i = 0; i < w->dw.W; ++i)
        if (!(T[i] > 0)) return fail("sse_set_temperature: temperatures must be positive");
    CU(cudaMemcpyAsync(w->dw.T, T, sizeof(double) * w->dw.W, cudaMemcpyHostToDevice, w->stream));  // src/sse.jl:403
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

int32_t sse_set_controller(sse_walkers *w, double target_worm_length_fraction, double num_worms_attenuation_factor) {
    if (!w) return fail("null handle");
    if (!(target_worm_length_fraction > 0) || !(num_worms_attenuation_factor >= 0) || !(num_worms_attenuation_factor <= 1))
        return fail("sse_set_controller: target_worm_length_fraction must be > 0 and the attenuation factor in [0, 1]");
    CU(cudaStreamSynchronize(w->stream));
    w->dw.twlf = target_worm_length_fraction;
    w->dw.atten = num_worms_attenuation_factor;
    return 0;
}

int32_t sse_double_beta(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = ensure_committed(w)) return s;
    CU(cudaSetDevice(w->model->device));
    const int grid = (w->dw.W + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    SSE_LAUNCH_KERNEL(k_double_beta, grid, WARPS_PER_CTA * 32, 0, w->stream, w->dw);
    CU(cudaGetLastError());
    w->have_vl = false;
    return sse_sync(w);
}

int32_t sse_set_injected_stream(sse_walkers *w, const uint64_t *stream, int64_t len) {
    if (!w) return fail("null handle");
    CU(cudaStreamSynchronize(w->stream));
    if (w->d_inj) { cudaFree(w->d_inj); w->d_inj = nullptr; }
    w->dw.inj = nullptr;
    w->dw.inj_len = 0;
    if (stream && len > 0) {
        size_t bytes = sizeof(uint64_t) * (size_t)len * w->dw.W;
        CU(cudaMalloc((void **)&w->d_inj, bytes));
        CU(cudaMemcpy(w->d_inj, stream, bytes, cudaMemcpyHostToDevice));
        w->dw.inj = w->d_inj;
        w->dw.inj_len = len;
        CU(cudaMemset(w->dw.draws, 0, sizeof(unsigned long long) * w->dw.W));
    }
    return 0;
}

int32_t sse_dbg_set_variant(sse_model *m, uint32_t variant) {
    if (!m) return fail("null handle");
    m->dm.variant = variant;  // DevModel travels by value with every launch: takes effect at the next one
    return 0;
}

int32_t sse_dbg_diagonal_update(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = ensure_committed(w)) return s;
    LaunchArgs a{};
    a.mode = MODE_DIAG;
    w->have_vl = false;
    if (int32_t s = launch(w, a)) return s;
    return sse_sync(w);
}

int32_t sse_dbg_make_vertex_list(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = ensure_committed(w)) return s;
    LaunchArgs a{};
    a.mode = MODE_MAKE_VL;
    if (int32_t s = launch(w, a)) return s;
    w->indexed = true;
    w->have_vl = true;
    return sse_sync(w);
}

int32_t sse_dbg_worm_update(sse_walkers *w, int32_t thermalized) {
    if (!w) return fail("null handle");
    if (!w->indexed || !w->have_vl) return fail("sse_dbg_worm_update: call sse_dbg_make_vertex_list first");
    LaunchArgs a{};
    a.mode = MODE_WORM_UPDATE;
    a.thermalized = thermalized;
    if (int32_t s = launch(w, a)) return s;
    return sse_sync(w);
}

int32_t sse_dbg_worm_traverse(sse_walkers *w, int32_t l0, int64_t p0, int32_t wormfunc0, int64_t *lengths) {
    if (!w || !lengths) return fail("null argument");
    if (!w->indexed || !w->have_vl) return fail("sse_dbg_worm_traverse: call sse_dbg_make_vertex_list first");
    if (l0 < 1 || l0 > 4 || p0 < 1 || wormfunc0 < 1) return fail("sse_dbg_worm_traverse: start out of range");
    std::vector<int> M(w->dw.W);
    CU(cudaMemcpy(M.data(), w->dw.M, sizeof(int) * M.size(), cudaMemcpyDeviceToHost));
    for (int m : M)
        if (p0 > m) return fail("sse_dbg_worm_traverse: p0 beyond the operator string");
    LaunchArgs a{};
    a.mode = MODE_WORM_TRAVERSE;
    a.l0 = l0 - 1;
    a.p0 = p0 - 1;
    a.w0 = wormfunc0;
    if (int32_t s = launch(w, a)) return s;
    if (int32_t s = sse_sync(w)) return s;
    std::vector<long long> len(w->dw.W);
    CU(cudaMemcpy(len.data(), w->dw.dbg_len, sizeof(long long) * len.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < len.size(); ++i) lengths[i] = len[i];
    return 0;
}

int32_t sse_dbg_get_vertex_list(sse_walkers *w, int32_t i, int64_t *vertices, int64_t m_len, int64_t *v_first, int64_t *v_last) {
    if (!w || !vertices || !v_first || !v_last) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("walker index out of range");
    if (!w->indexed || !w->have_vl) return fail("sse_dbg_get_vertex_list: no vertex list (call sse_dbg_make_vertex_list)");
    CU(cudaStreamSynchronize(w->stream));
    int M = 0, n = 0;
    const int N = w->model->dm.n_sites;
    CU(cudaMemcpy(&M, w->dw.M + i, sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&n, w->dw.n + i, sizeof(int), cudaMemcpyDeviceToHost));
    if (m_len < M) return fail("sse_dbg_get_vertex_list: vertices buffer too small");
    std::vector<uint32_t> ops(M), vf(N), vl(N);
    std::vector<uint4> rec2(2 * (size_t)n), rec(n);
    if (M) CU(cudaMemcpy(ops.data(), w->dw.ops + (size_t)i * w->dw.M_cap, sizeof(uint32_t) * M, cudaMemcpyDeviceToHost));
    if (n) CU(cudaMemcpy(rec2.data(), w->dw.rec + 2 * (size_t)i * w->dw.n_cap, sizeof(uint4) * 2 * n, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; ++k) rec[k] = rec2[2 * (size_t)k];
    CU(cudaMemcpy(vf.data(), w->dw.vfirst + (size_t)i * N, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(vl.data(), w->dw.vlast + (size_t)i * N, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    std::vector<int64_t> pos(n, -1);
    for (int p = 0; p < M; ++p)
        if (ops[p]) {
            if (ops[p] - 1 >= (uint32_t)n) return fail("sse_dbg_get_vertex_list: corrupt record index");
            pos[ops[p] - 1] = p;
        }
    auto link = [&](const uint4 &r, int j) -> uint32_t { return j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w; };
    for (int64_t p = 0; p < M; ++p)
        for (int l = 0; l < 4; ++l) {
            int64_t *dst = vertices + (p * 4 + l) * 2;
            dst[0] = dst[1] = -1;
            if (!ops[p]) continue;
            uint32_t lk = link(rec[ops[p] - 1], l);
            if (lk == NONE24 || (lk >> 2) >= (uint32_t)n) return fail("sse_dbg_get_vertex_list: dangling link");
            dst[0] = (lk & 3) + 1;
            dst[1] = pos[lk >> 2] + 1;
        }
    for (int s = 0; s < N; ++s) {
        for (int which = 0; which < 2; ++which) {
            uint32_t v = which ? vl[s] : vf[s];
            int64_t *dst = (which ? v_last : v_first) + 2 * s;
            if (v == NONE32) { dst[0] = dst[1] = -1; }
            else { dst[0] = (v & 3) + 1; dst[1] = pos[v >> 2] + 1; }
        }
    }
    return 0;
}

int32_t sse_dbg_commit(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = ensure_committed(w)) return s;
    return sse_sync(w);
}

}  // extern "C"
