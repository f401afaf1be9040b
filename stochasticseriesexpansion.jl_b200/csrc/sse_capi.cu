// sse_capi.cu — C ABI of libsse_b200.so (include/sse_b200.h): table flattening to the device image, device memory
// ownership, launches of the kernels (sse_sweep.cuh), checkpoint conversion between the device layout (occupancy
// bitmap + record ring) and the reference's UInt64 OperCode strings (src/opercode.jl:43-47).
// No CPU fallback exists: every entry point runs on the GPU or returns an error status.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "sse_sweep.cuh"

#ifndef SSE_EMU
#include <dlfcn.h>
#endif

using namespace sse;

// Kernel launches go through one macro so that the test-only warp emulator (tests/emu) can compile this file with g++.
// A launch first discards whatever error another library left in the runtime's per-thread slot (NCCL's device probing
// leaves cudaErrorInvalidDevice behind), so that the cudaGetLastError() after it reports this launch only.
#ifndef SSE_LAUNCH_KERNEL
#define SSE_LAUNCH_KERNEL(kern, grid, block, smem_bytes, stream, ...) \
    do {                                                              \
        (void)cudaGetLastError();                                     \
        kern<<<grid, block, smem_bytes, stream>>>(__VA_ARGS__);       \
    } while (0)
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
    int n_sm = 1;
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
    int worm_warps = 0, stream_warps = 0;  // launch shape of k_sweep; 0 = automatic
    bool have_vl = false;                  // leg links valid (parity hooks)
    bool maybe_in_flight = false;          // sse_advance may have parked walkers inside a sweep
    unsigned long long *d_inj = nullptr;
    int64_t bytes = 0;
    std::vector<void *> allocs;
    void *comm = nullptr;                  // ncclComm_t of sse_comm_init (bin reduction only)
    int comm_rank = 0, comm_nranks = 1;
    void *d_red = nullptr;                 // staging of sse_reduce_bins
    size_t red_bytes = 0;
    int32_t *d_ladder = nullptr;           // sse_pt_set_ladder: walker index at each temperature rank; then the accept counter
    int n_ladder = 0;
};

namespace {

// ---- strided access to one field of the per-walker control blocks ----
template <class T>
int32_t get_field(sse_walkers *w, size_t offset, std::vector<T> &out) {
    out.resize(w->dw.W);
    CU(cudaMemcpy2DAsync(out.data(), sizeof(T), reinterpret_cast<const char *>(w->dw.ctl) + offset, sizeof(WalkerCtl), sizeof(T),
                         (size_t)w->dw.W, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}
template <class T>
int32_t set_field(sse_walkers *w, size_t offset, const T *src) {
    CU(cudaMemcpy2DAsync(reinterpret_cast<char *>(w->dw.ctl) + offset, sizeof(WalkerCtl), src, sizeof(T), sizeof(T), (size_t)w->dw.W,
                         cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}
#define CTL_OFF(f) offsetof(WalkerCtl, f)

int32_t check_flags(sse_walkers *w) {
    unsigned long long any = 0;
    CU(cudaMemcpyAsync(&any, w->dw.counters + SSE_CNT_ANY_FATAL, sizeof(any), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    if (!any) return 0;
    std::vector<uint32_t> f;
    if (int32_t s = get_field(w, CTL_OFF(flags), f)) return s;
    for (int i = 0; i < w->dw.W; ++i) {
        if (f[i] & SSE_FLAG_M_OVERFLOW)
            return fail("walker " + std::to_string(i) + ": operator string outgrew m_capacity (recreate with a larger m_capacity, or sse_grow_capacity)");
        if (f[i] & SSE_FLAG_N_OVERFLOW)
            return fail("walker " + std::to_string(i) + ": more operators than n_capacity (recreate with a larger n_capacity, or sse_grow_capacity)");
        if (f[i] & SSE_FLAG_STREAM_EXHAUSTED)
            return fail("walker " + std::to_string(i) + ": injected random stream exhausted");
    }
    return 0;
}

// shared-memory level of the streaming warps: 1 = state[N] + mark[N] in shared memory, 0 = in global memory
int phase_level(const sse_model *m) {
    int level = (m->dm.tl.bytes + PHASE_WARPS * stream_scratch_bytes(m->dm.n_sites, 1) <= 99 * 1024) ? 1 : 0;
    if (const char *lv = getenv("SSE_B200_SMEM_LEVEL")) level = std::min(level, std::max(0, atoi(lv)));
    return level;
}

int32_t launch_phase(sse_walkers *w, PhaseArgs a) {
    const sse_model *m = w->model;
    CU(cudaSetDevice(m->device));
    a.level = phase_level(m);
    if (!a.level && !w->dw.mark) return fail("internal: mark[] scratch missing");
    const int grid = (w->dw.W + PHASE_WARPS - 1) / PHASE_WARPS;
    const size_t smem = (size_t)m->dm.tl.bytes + (size_t)PHASE_WARPS * stream_scratch_bytes(m->dm.n_sites, a.level);
    if (smem > 48 * 1024) {
        CU(cudaFuncSetAttribute(k_phase<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_phase<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (w->dw.inj)
        SSE_LAUNCH_KERNEL(k_phase<true>, grid, PHASE_WARPS * 32, smem, w->stream, m->dm, w->dw, a);
    else
        SSE_LAUNCH_KERNEL(k_phase<false>, grid, PHASE_WARPS * 32, smem, w->stream, m->dm, w->dw, a);
    CU(cudaGetLastError());
    return 0;
}

// Launch shape of k_sweep: one CTA per SM (fewer if there are fewer walkers); worm warps so that every walker of a CTA
// has its own lane (up to 8 warps = 256 chains per SM: more does not raise the random-sector rate of the memory system,
// profiles/r2_chase_lanes.txt), the remaining warps stream.
struct SweepShape {
    int grid, worm_warps, stream_warps, level, nloc_max;
    size_t smem;
};

int32_t sweep_shape(const sse_walkers *w, SweepShape &sh) {
    const sse_model *m = w->model;
    const int W = w->dw.W, N = m->dm.n_sites;
    sh.grid = std::min(W, m->n_sm);
    sh.nloc_max = (W + sh.grid - 1) / sh.grid;
    const int ww_max = SPLIT_REGS ? WORM_GROUP_WARPS : std::min(8, SWEEP_MAX_WARPS - 1);
    int ww = w->worm_warps > 0 ? w->worm_warps : std::min(ww_max, (sh.nloc_max + 31) / 32);
    const int sw_max = SPLIT_REGS ? STREAM_GROUP_WARPS : SWEEP_MAX_WARPS - ww;
    int sw = w->stream_warps > 0 ? w->stream_warps : std::min(sw_max, std::max(1, sh.nloc_max));
    if (ww < 1 || sw < 1 || ww > ww_max || sw > sw_max)
        return fail("launch shape: need 1 <= worm_warps <= " + std::to_string(ww_max) + " and 1 <= stream_warps <= " + std::to_string(sw_max));
    const int budget = 227 * 1024 - 1024;
    const int fixed = m->dm.tl.bytes + sched_bytes(sh.nloc_max);
    // stream warps keep state[] and mark[] (level 1) and vlast[] (level 2) of their walker in shared memory.  The number
    // of stream warps matters more than the level (measured at L = 64: 10 warps at level 1 beat 9 at level 2 by 5 %): the
    // highest level at which all wanted warps fit wins; otherwise level 1 with what fits (at least 4), else global memory.
    int max_level = 2;
    if (const char *lv = getenv("SSE_B200_SMEM_LEVEL")) max_level = std::min(max_level, std::max(0, atoi(lv)));
    sh.level = 0;
    for (int level = max_level; level >= 1; --level) {
        const int fit = (budget - fixed) / stream_scratch_bytes(N, level);
        if (fit >= sw || (level == 1 && fit >= 4)) {
            sh.level = level;
            sw = std::min(sw, fit);
            break;
        }
    }
    if (!sh.level && !w->dw.mark) return fail("internal: mark[] scratch missing");
    sh.worm_warps = ww;
    sh.stream_warps = sw;
    sh.smem = (size_t)fixed + (size_t)sw * stream_scratch_bytes(N, sh.level);
    if ((int)sh.smem > budget) return fail("launch shape: shared memory exceeded (too many walkers per SM for the status table)");
    return 0;
}

int32_t launch_sweep(sse_walkers *w, int n_sweeps, unsigned long long budget, int reset, int thermalized, int measure) {
    const sse_model *m = w->model;
    CU(cudaSetDevice(m->device));
    SweepShape sh;
    if (int32_t s = sweep_shape(w, sh)) return s;
    SweepArgs a{};
    a.n_sweeps = n_sweeps;
    a.budget = budget;
    a.reset = reset;
    a.thermalized = thermalized;
    a.measure = measure;
    a.worm_warps = sh.worm_warps;
    a.stream_warps = sh.stream_warps;
    a.level = sh.level;
    a.nloc_max = sh.nloc_max;
    const int block = (SPLIT_REGS ? SWEEP_MAX_WARPS : sh.worm_warps + sh.stream_warps) * 32;
    if (sh.smem > 48 * 1024) {
        CU(cudaFuncSetAttribute(k_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
        CU(cudaFuncSetAttribute(k_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
    }
    if (w->dw.inj)
        SSE_LAUNCH_KERNEL(k_sweep<true>, sh.grid, block, sh.smem, w->stream, m->dm, w->dw, a);
    else
        SSE_LAUNCH_KERNEL(k_sweep<false>, sh.grid, block, sh.smem, w->stream, m->dm, w->dw, a);
    CU(cudaGetLastError());
    w->have_vl = false;
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

// sse_get_state, sse_measure, sse_double_beta and the parity hooks need every walker between two sweeps
int32_t require_between_sweeps(sse_walkers *w, const char *who) {
    CU(cudaStreamSynchronize(w->stream));
    if (!w->maybe_in_flight) return 0;
    std::vector<uint32_t> ph;
    if (int32_t s = get_field(w, CTL_OFF(phase), ph)) return s;
    for (int i = 0; i < w->dw.W; ++i)
        if (ph[i]) return fail(std::string(who) + ": walker " + std::to_string(i) + " is parked inside a sweep (sse_advance); call sse_finish_sweeps first");
    w->maybe_in_flight = false;
    return 0;
}

// ---- NCCL, resolved at run time: the library has no link-time dependency on it, and a host that already loaded an NCCL
// (PyTorch bundles one) shares that copy.  Used for ONE thing: summing binned observables over ranks. ----
struct NcclApi {
    bool ok = false;
    std::string why;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, sse_nccl_id, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
#ifdef SSE_EMU
    api.why = "NCCL is not available in the emulator build";
#else
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // an NCCL the host process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("libnccl.so.2 not found: ") + dlerror(); return api; }
    api.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, sse_nccl_id, int))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
#endif
    return api;
}
#define NC(call)                                                                                         \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != 0) return fail(std::string(#call) + " failed: " + nccl().GetErrorString(r__));        \
    } while (0)

// Per-group sums of the walkers' accumulators in a FIXED order (walker order inside each group), so a bin is
// reproducible bit for bit: thread (g, i) adds column i of the walkers listed for group g.
// out[g][n_obs + 2]: n_obs sums, then the two counts (as doubles: exact below 2^53).
__global__ void k_reduce_bins(const DevWalkers dw, const int32_t *start, const int32_t *members, int n_groups, double *out) {
    const int ncol = dw.n_obs + 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_groups * ncol) return;
    const int g = t / ncol, i = t % ncol;
    double s = 0.0;
    for (int k = start[g]; k < start[g + 1]; ++k) {
        const int w = members[k];
        s += i < dw.n_obs ? dw.acc[(size_t)w * dw.n_obs + i] : (double)dw.acc_cnt[2 * w + (i - dw.n_obs)];
    }
    out[t] = s;
}

// One round of neighbour swaps on the temperature ladder (parallel tempering, src/sse.jl:390-405): thread i proposes to
// exchange the temperatures of the walkers at ranks r = parity + 2i and r + 1 and accepts with
// min(1, exp(lw_a + lw_b)), lw_x = parallel_tempering_log_weight_ratio(x, :T, T_other) = -n_x * log(T_other / T_x)
// (sse.jl:395).  The uniform is draw i of the Philox stream (seed, step): the walkers' own streams are not touched.
// Configurations never move, only the temperature labels (parallel_tempering_change_parameter!, sse.jl:398-405).
__global__ void k_pt_exchange(const DevWalkers dw, int32_t *ladder, int n_ladder, int parity, unsigned long long seed,
                              unsigned long long step, int32_t *n_accepted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = parity + 2 * i;
    if (r + 1 >= n_ladder) return;
    const int a = ladder[r], b = ladder[r + 1];
    WalkerCtl *ca = dw.ctl + a, *cb = dw.ctl + b;
    if ((ca->flags | cb->flags) & FATAL_FLAGS) return;
    const double Ta = ca->T, Tb = cb->T;
    const double lw = -(double)ca->n * log(Tb / Ta) + -(double)cb->n * log(Ta / Tb);
    const double u = sse_u01(sse_philox_draw(seed, step, (unsigned long long)i));
    if (log(u > 1e-300 ? u : 1e-300) < lw) {
        ca->T = Tb;
        cb->T = Ta;
        ladder[r] = b;
        ladder[r + 1] = a;
        atomicAdd(n_accepted, 1);
    }
}

// sse_grow_capacity: every walker's bitmap words and records (unrolled from its ring, generation start back at 0) move to
// arrays of the new capacities; a pending "string outgrew m_capacity" flag is cleared, because that overflow is detected
// before the sweep modifies anything.  One warp per walker.
__global__ void k_grow(const DevWalkers src, uint2 *new_words, long long new_Mw_cap, uint4 *new_rec, long long new_R_cap) {
    const int w = blockIdx.x * PHASE_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= src.W) return;
    WalkerCtl *ctl = src.ctl + w;
    const int M = ctl->M, n = ctl->n;
    const uint32_t G = ctl->G;
    const uint2 *ow = src.words + (size_t)w * src.Mw_cap;
    uint2 *nw = new_words + (size_t)w * new_Mw_cap;
    const int used = (M + 31) / 32;
    for (long long i = lane; i < new_Mw_cap; i += 32) nw[i] = i < used ? ow[i] : make_uint2(0u, 0u);
    const uint4 *orec = src.rec + (size_t)w * src.R_cap;
    uint4 *nrec = new_rec + (size_t)w * new_R_cap;
    for (int k = lane; k < n; k += 32) nrec[k] = orec[ring(G, (uint32_t)src.R_cap, (uint32_t)k)];
    __syncwarp();
    if (lane == 0) {
        ctl->G = 0;
        ctl->flags &= ~SSE_FLAG_M_OVERFLOW;
    }
}

int64_t ring_size(int64_t n_cap) { return n_cap + std::max<int64_t>(2048, n_cap / 16); }

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
    CU(cudaDeviceGetAttribute(&m->n_sm, cudaDevAttrMultiProcessorCount, m->device));
    if (const char *e = getenv("SSE_B200_CTAS")) m->n_sm = std::max(1, atoi(e));  // tests: force a grid size
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
                           (uint32_t)d->type_diag_off[t], 0u);
    }
    // Estimator tables est_values[e][site][state] usually hold a handful of distinct rows (e.g. +-(1/2, -1/2) for a
    // staggered magnetization).  With at most two estimators and at most 255 distinct rows each, the rows go to shared
    // memory and the row index of a bond's two sites rides in the bond table, so the measurement can be fused into the
    // diagonal-update pass without a dependent global load.
    const int md = d->est_max_dim > 0 ? d->est_max_dim : 1;
    std::vector<double> est_rows;
    int est_nrows = 0;
    bool est_compressed = d->n_estimators >= 1 && d->n_estimators <= 2;
    if (est_compressed) {
        std::vector<std::vector<std::vector<double>>> rows(d->n_estimators);
        std::vector<std::vector<int>> rowid(d->n_estimators, std::vector<int>(d->n_sites, 0));
        for (int e = 0; e < d->n_estimators && est_compressed; ++e)
            for (int s = 0; s < d->n_sites; ++s) {
                const double *r = d->est_values + ((size_t)e * d->n_sites + s) * md;
                int id = -1;
                for (size_t k = 0; k < rows[e].size(); ++k)
                    if (memcmp(rows[e][k].data(), r, sizeof(double) * md) == 0) { id = (int)k; break; }
                if (id < 0) {
                    if (rows[e].size() >= 255) { est_compressed = false; break; }
                    id = (int)rows[e].size();
                    rows[e].emplace_back(r, r + md);
                }
                rowid[e][s] = id;
            }
        if (est_compressed) {
            for (auto &re : rows) est_nrows = std::max(est_nrows, (int)re.size());
            est_rows.assign((size_t)d->n_estimators * est_nrows * md, 0.0);
            for (int e = 0; e < d->n_estimators; ++e)
                for (size_t k = 0; k < rows[e].size(); ++k)
                    std::copy(rows[e][k].begin(), rows[e][k].end(), est_rows.begin() + ((size_t)e * est_nrows + k) * md);
            for (int b = 0; b < d->n_bonds; ++b) {
                const int sa = d->bond_sites[2 * b], sb = d->bond_sites[2 * b + 1];
                uint32_t wv = (uint32_t)rowid[0][sa] | ((uint32_t)rowid[0][sb] << 8);
                if (d->n_estimators > 1) wv |= ((uint32_t)rowid[1][sa] << 16) | ((uint32_t)rowid[1][sb] << 24);
                bi[b].w = wv;
            }
        }
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
    tl.off_estrows = est_compressed ? take(8 * (int)est_rows.size()) : -1;
    tl.est_nrows = est_nrows;
    tl.bytes = off;
    if (tl.bytes > 64 * 1024) { return fail("sse_model_create: vertex tables exceed the 64 KB shared-memory budget"); }
    std::vector<uint8_t> blob(tl.bytes, 0);
    auto *outc = reinterpret_cast<uint4 *>(blob.data() + tl.off_outc);
    auto *wts = reinterpret_cast<double *>(blob.data() + tl.off_weights);
    auto *t1 = reinterpret_cast<uint4 *>(blob.data() + tl.off_t1);
    auto *vinfo = reinterpret_cast<uint32_t *>(blob.data() + tl.off_vinfo);
    auto *diagv = reinterpret_cast<uint16_t *>(blob.data() + tl.off_diagv);
    auto *vneg = blob.data() + tl.off_vneg;
    if (est_compressed) memcpy(blob.data() + tl.off_estrows, est_rows.data(), 8 * est_rows.size());
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
    if (o->n_capacity < 64 || o->n_capacity > (1ll << 22) - 1) return fail("sse_walkers_create: n_capacity out of range [64, 2^22 - 1]");
    if (o->device >= 0 && o->device != m->device) return fail("sse_walkers_create: model was created on another device");
    CU(cudaSetDevice(m->device));
    std::unique_ptr<sse_walkers, int32_t (*)(sse_walkers *)> guard(new sse_walkers(), sse_walkers_destroy);  // freed on every error path
    sse_walkers *w = guard.get();
    w->model = m;
    DevWalkers &dw = w->dw;
    const int W = o->n_walkers, N = m->dm.n_sites;
    dw.W = W;
    dw.M_cap = (o->m_capacity + 31) & ~31ll;
    dw.Mw_cap = dw.M_cap / 32;
    dw.n_cap = o->n_capacity;
    dw.R_cap = ring_size(dw.n_cap);
    dw.n_obs = SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * m->dm.n_est;
    int32_t s = 0;
    s |= dev_alloc(w, &dw.words, (size_t)W * dw.Mw_cap, true);
    s |= dev_alloc(w, &dw.rec, (size_t)W * dw.R_cap, false);
    s |= dev_alloc(w, &dw.state, (size_t)W * N, false);
    // stream warps keep state[N] + mark[N] in shared memory when they fit, else in global scratch
    {
        const int fixed = m->dm.tl.bytes + sched_bytes(1024);
        const bool fits_sweep = (227 * 1024 - 1024 - fixed) / stream_scratch_bytes(N, 1) >= 4;
        if (!fits_sweep || !phase_level(m) || getenv("SSE_B200_SMEM_LEVEL")) s |= dev_alloc(w, &dw.mark, (size_t)W * N, true);
    }
    s |= dev_alloc(w, &dw.vfirst, (size_t)W * N, false);
    s |= dev_alloc(w, &dw.vlast, (size_t)W * N, false);
    s |= dev_alloc(w, &dw.ctl, (size_t)W, true);
    s |= dev_alloc(w, &dw.acc, (size_t)W * dw.n_obs, true);
    s |= dev_alloc(w, &dw.acc_cnt, (size_t)W * 2, true);
    s |= dev_alloc(w, &dw.counters, SSE_N_COUNTERS, true);
    s |= dev_alloc(w, &dw.dbg_len, W, true);
    s |= dev_alloc(w, &dw.obs_out, (size_t)W * dw.n_obs, true);
    if (s) return 1;
    dw.inj = nullptr;
    dw.inj_len = 0;
    dw.seed = o->seed;
    dw.wid_off = o->walker_id_offset;
    dw.twlf = o->target_worm_length_fraction;
    dw.atten = o->num_worms_attenuation_factor;
    std::vector<WalkerCtl> ctl(W);
    for (int i = 0; i < W; ++i) {
        if (!(o->T[i] > 0)) return fail("sse_walkers_create: temperatures must be positive");
        memset(&ctl[i], 0, sizeof(WalkerCtl));
        ctl[i].T = o->T[i];
        ctl[i].num_worms = o->init_num_worms;
        ctl[i].avg_wl = 1.0;
        ctl[i].last_wlf = NAN;
    }
    CU(cudaMemcpy(dw.ctl, ctl.data(), sizeof(WalkerCtl) * W, cudaMemcpyHostToDevice));
    // everything above went through the legacy default stream (memsets, pageable copies that return once staged); the
    // handle's own stream is non-blocking, so it is not ordered against them: drain before the first launch can happen
    CU(cudaDeviceSynchronize());
    CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    w->own_stream = true;
    if (const char *e = getenv("SSE_B200_WORM_WARPS")) w->worm_warps = atoi(e);
    if (const char *e = getenv("SSE_B200_STREAM_WARPS")) w->stream_warps = atoi(e);
    {
        SweepShape sh;
        if (sweep_shape(w, sh)) return 1;
    }
    *out = guard.release();
    return 0;
}

int32_t sse_grow_capacity(sse_walkers *w, int64_t m_capacity, int64_t n_capacity) {
    if (!w) return fail("null handle");
    DevWalkers &dw = w->dw;
    if (((m_capacity + 31) & ~31ll) < dw.M_cap || n_capacity < dw.n_cap) return fail("sse_grow_capacity: capacities can only grow");
    if (m_capacity >= (1ll << 31) || n_capacity > (1ll << 22) - 1) return fail("sse_grow_capacity: capacity out of range (m < 2^31, n <= 2^22 - 1)");
    CU(cudaSetDevice(w->model->device));
    CU(cudaStreamSynchronize(w->stream));
    {  // walkers parked inside a sweep keep links into their ring: finish those sweeps first
        std::vector<uint32_t> ph;
        if (int32_t s = get_field(w, CTL_OFF(phase), ph)) return s;
        for (int i = 0; i < dw.W; ++i)
            if (ph[i]) return fail("sse_grow_capacity: walker " + std::to_string(i) + " is parked inside a sweep; call sse_finish_sweeps first");
    }
    const int64_t M_cap = (m_capacity + 31) & ~31ll, Mw_cap = M_cap / 32, R_cap = ring_size(n_capacity);
    uint2 *nwords = nullptr;
    uint4 *nrec = nullptr;
    CU(cudaMalloc((void **)&nwords, sizeof(uint2) * (size_t)dw.W * Mw_cap));
    if (cudaMalloc((void **)&nrec, sizeof(uint4) * (size_t)dw.W * R_cap) != cudaSuccess) {
        cudaFree(nwords);
        return fail("sse_grow_capacity: out of device memory (old and new arrays must coexist during the move)");
    }
    const int grid = (dw.W + PHASE_WARPS - 1) / PHASE_WARPS;
    SSE_LAUNCH_KERNEL(k_grow, grid, PHASE_WARPS * 32, 0, w->stream, dw, nwords, (long long)Mw_cap, nrec, (long long)R_cap);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(w->stream));
    for (void *&p : w->allocs) {
        if (p == dw.words) { cudaFree(p); p = nwords; }
        else if (p == dw.rec) { cudaFree(p); p = nrec; }
    }
    w->bytes += (int64_t)(sizeof(uint2) * (size_t)dw.W * (Mw_cap - dw.Mw_cap) + sizeof(uint4) * (size_t)dw.W * (R_cap - dw.R_cap));
    dw.words = nwords;
    dw.rec = nrec;
    dw.M_cap = M_cap;
    dw.Mw_cap = Mw_cap;
    dw.n_cap = n_capacity;
    dw.R_cap = R_cap;
    w->have_vl = false;
    // the fatal marker is rebuilt from the remaining flags
    std::vector<uint32_t> f;
    if (int32_t s = get_field(w, CTL_OFF(flags), f)) return s;
    unsigned long long any = 0;
    for (uint32_t x : f) any |= (x & FATAL_FLAGS) ? 1ull : 0ull;
    CU(cudaMemcpyAsync(dw.counters + SSE_CNT_ANY_FATAL, &any, sizeof(any), cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

int32_t sse_set_launch_shape(sse_walkers *w, int32_t worm_warps, int32_t stream_warps) {
    if (!w) return fail("null handle");
    if (worm_warps < 0 || stream_warps < 0) return fail("sse_set_launch_shape: negative warp count");
    CU(cudaStreamSynchronize(w->stream));
    const int ow = w->worm_warps, os = w->stream_warps;
    w->worm_warps = worm_warps;
    w->stream_warps = stream_warps;
    SweepShape sh;
    if (sweep_shape(w, sh)) {
        w->worm_warps = ow;
        w->stream_warps = os;
        return 1;
    }
    return 0;
}

int32_t sse_walkers_destroy(sse_walkers *w) {
    if (!w) return 0;
    cudaSetDevice(w->model->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    for (void *p : w->allocs) cudaFree(p);
    if (w->d_inj) cudaFree(w->d_inj);
    if (w->d_red) cudaFree(w->d_red);
    if (w->d_ladder) cudaFree(w->d_ladder);
    if (w->comm && nccl().ok) nccl().CommDestroy(w->comm);
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

int64_t sse_walker_bytes(const sse_model *m, int64_t m_capacity, int64_t n_capacity) {
    if (!m || m_capacity < 0 || n_capacity < 0) return -1;
    const int64_t N = m->dm.n_sites, n_obs = SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * m->dm.n_est;
    int64_t b = ((m_capacity + 31) / 32) * (int64_t)sizeof(uint2) + ring_size(n_capacity) * (int64_t)sizeof(uint4);
    b += N * (1 + 4 + 4) + (int64_t)sizeof(WalkerCtl) + n_obs * 8 * 2 + 16 + 8;
    const int fixed = m->dm.tl.bytes + sched_bytes(1024);
    if ((227 * 1024 - 1024 - fixed) / stream_scratch_bytes((int)N, 1) < 4 || !phase_level(m)) b += N;  // global mark[] scratch
    return b;
}

int32_t sse_n_observables(const sse_walkers *w) { return w ? w->dw.n_obs : -1; }
int64_t sse_device_bytes(const sse_walkers *w) { return w ? w->bytes : -1; }

int32_t sse_init(sse_walkers *w, int64_t init_opstring_cutoff, int32_t diagonal_warmup_sweeps) {
    if (!w) return fail("null handle");
    const int W = w->dw.W;
    CU(cudaSetDevice(w->model->device));
    CU(cudaStreamSynchronize(w->stream));
    std::vector<WalkerCtl> ctl(W);
    CU(cudaMemcpyAsync(ctl.data(), w->dw.ctl, sizeof(WalkerCtl) * W, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < W; ++i) {
        // round(Int, length(sites) * T) (src/sse.jl:51): round-half-even
        long long m0 = init_opstring_cutoff >= 0 ? init_opstring_cutoff : (long long)std::nearbyint((double)w->model->dm.n_sites * ctl[i].T);
        if (m0 > w->dw.M_cap) return fail("sse_init: init_opstring_cutoff exceeds m_capacity");
        WalkerCtl &c = ctl[i];
        c.M = (int)m0;
        c.n = 0;
        c.G = 0;
        c.flags = 0;  // a walker that overflowed earlier starts afresh
        c.phase = 0;
        c.worms_left = 0;
        c.inworm = 0;
        c.sweep_visits = 0;
        c.sweeps_done = 0;
        c.sweeps_left = 0;
    }
    CU(cudaMemsetAsync(w->dw.words, 0, sizeof(uint2) * (size_t)W * w->dw.Mw_cap, w->stream));
    CU(cudaMemcpyAsync(w->dw.ctl, ctl.data(), sizeof(WalkerCtl) * W, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemsetAsync(w->dw.counters + SSE_CNT_ANY_FATAL, 0, sizeof(unsigned long long), w->stream));
    CU(cudaStreamSynchronize(w->stream));  // ctl is a host vector: the copy must be done before it goes out of scope
    w->have_vl = false;
    w->maybe_in_flight = false;
    PhaseArgs a{};
    a.mode = MODE_INIT;
    a.warmup = diagonal_warmup_sweeps;
    if (int32_t s = launch_phase(w, a)) return s;
    return sse_sync(w);
}

int32_t sse_sweep(sse_walkers *w, int32_t n_sweeps, int32_t thermalized, int32_t measure) {
    if (!w) return fail("null handle");
    if (n_sweeps <= 0) return 0;
    return launch_sweep(w, n_sweeps, ~0ull, 1, thermalized, measure);
}

int32_t sse_advance(sse_walkers *w, int32_t max_sweeps, uint64_t visit_budget, int32_t thermalized, int32_t measure) {
    if (!w) return fail("null handle");
    if (max_sweeps <= 0 || visit_budget == 0) return 0;
    w->maybe_in_flight = true;
    return launch_sweep(w, max_sweeps, (unsigned long long)visit_budget, 1, thermalized, measure);
}

int32_t sse_finish_sweeps(sse_walkers *w, int32_t thermalized, int32_t measure) {
    if (!w) return fail("null handle");
    // quota 0: sweeps in flight are completed (their worm phase runs to its end), no new sweep starts
    if (int32_t s = launch_sweep(w, 0, ~0ull, 1, thermalized, measure)) return s;
    w->maybe_in_flight = false;
    return 0;
}

int32_t sse_continue_sweeps(sse_walkers *w, int32_t thermalized, int32_t measure) {
    if (!w) return fail("null handle");
    // reset = 0: every walker keeps the quota the interrupted sse_sweep left it with
    return launch_sweep(w, 0, ~0ull, 0, thermalized, measure);
}

int32_t sse_get_progress(sse_walkers *w, uint64_t *sweeps_done, uint8_t *in_flight) {
    if (!w || !sweeps_done) return fail("null argument");
    std::vector<unsigned long long> sd;
    if (int32_t s = get_field(w, CTL_OFF(sweeps_done), sd)) return s;
    for (int i = 0; i < w->dw.W; ++i) sweeps_done[i] = sd[i];
    if (in_flight) {
        std::vector<uint32_t> ph;
        if (int32_t s = get_field(w, CTL_OFF(phase), ph)) return s;
        for (int i = 0; i < w->dw.W; ++i) in_flight[i] = (uint8_t)(ph[i] != 0);
    }
    return 0;
}

int32_t sse_sync(sse_walkers *w) {
    if (!w) return fail("null handle");
    CU(cudaStreamSynchronize(w->stream));
    return check_flags(w);
}

int32_t sse_measure(sse_walkers *w, double *out) {
    if (!w || !out) return fail("null argument");
    if (int32_t s = require_between_sweeps(w, "sse_measure")) return s;
    PhaseArgs a{};
    a.mode = MODE_MEASURE;
    if (int32_t s = launch_phase(w, a)) return s;
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

int32_t sse_comm_unique_id(sse_nccl_id *id) {
    if (!id) return fail("null argument");
    if (!nccl().ok) return fail("sse_comm_unique_id: " + nccl().why);
    NC(nccl().GetUniqueId(id));
    return 0;
}

int32_t sse_comm_init(sse_walkers *w, const sse_nccl_id *id, int32_t rank, int32_t nranks) {
    if (!w || !id) return fail("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("sse_comm_init: rank out of range");
    if (!nccl().ok) return fail("sse_comm_init: " + nccl().why);
    CU(cudaSetDevice(w->model->device));
    if (w->comm) { nccl().CommDestroy(w->comm); w->comm = nullptr; }
    NC(nccl().CommInitRank(&w->comm, nranks, *id, rank));
    (void)cudaGetLastError();
    w->comm_rank = rank;
    w->comm_nranks = nranks;
    return 0;
}

int32_t sse_reduce_bins(sse_walkers *w, const int32_t *group, int32_t n_groups, double *sums, int64_t *counts, int32_t reset) {
    if (!w || !sums || !counts) return fail("null argument");
    if (n_groups < 1) return fail("sse_reduce_bins: n_groups must be positive");
    const int W = w->dw.W, ncol = w->dw.n_obs + 2;
    CU(cudaSetDevice(w->model->device));
    // group -> walkers (CSR), walker order inside a group
    std::vector<int32_t> start(n_groups + 1, 0), members(W);
    for (int i = 0; i < W; ++i) {
        const int g = group ? group[i] : 0;
        if (g < 0 || g >= n_groups) return fail("sse_reduce_bins: group index out of range at walker " + std::to_string(i));
        ++start[g + 1];
    }
    for (int g = 0; g < n_groups; ++g) start[g + 1] += start[g];
    {
        std::vector<int32_t> fill(start.begin(), start.end() - 1);
        for (int i = 0; i < W; ++i) members[fill[group ? group[i] : 0]++] = i;
    }
    const size_t b_out = sizeof(double) * (size_t)n_groups * ncol, b_start = sizeof(int32_t) * (n_groups + 1), b_mem = sizeof(int32_t) * W;
    const size_t off_start = (b_out + 15) & ~(size_t)15, off_mem = (off_start + b_start + 15) & ~(size_t)15, need = off_mem + b_mem;
    if (need > w->red_bytes) {
        CU(cudaStreamSynchronize(w->stream));
        if (w->d_red) cudaFree(w->d_red);
        w->d_red = nullptr;
        CU(cudaMalloc(&w->d_red, need));
        w->red_bytes = need;
    }
    char *base = static_cast<char *>(w->d_red);
    double *d_out = reinterpret_cast<double *>(base);
    CU(cudaMemcpyAsync(base + off_start, start.data(), b_start, cudaMemcpyHostToDevice, w->stream));
    CU(cudaMemcpyAsync(base + off_mem, members.data(), b_mem, cudaMemcpyHostToDevice, w->stream));
    const int threads = 128, blocks = (n_groups * ncol + threads - 1) / threads;
    SSE_LAUNCH_KERNEL(k_reduce_bins, blocks, threads, 0, w->stream, w->dw, reinterpret_cast<const int32_t *>(base + off_start),
                      reinterpret_cast<const int32_t *>(base + off_mem), (int)n_groups, d_out);
    CU(cudaGetLastError());
    if (w->comm) NC(nccl().AllReduce(d_out, d_out, (size_t)n_groups * ncol, 8 /* ncclFloat64 */, 0 /* ncclSum */, w->comm, w->stream));
    std::vector<double> host((size_t)n_groups * ncol);
    CU(cudaMemcpyAsync(host.data(), d_out, b_out, cudaMemcpyDeviceToHost, w->stream));
    if (reset) {
        CU(cudaMemsetAsync(w->dw.acc, 0, sizeof(double) * (size_t)W * w->dw.n_obs, w->stream));
        CU(cudaMemsetAsync(w->dw.acc_cnt, 0, sizeof(long long) * (size_t)W * 2, w->stream));
    }
    CU(cudaStreamSynchronize(w->stream));  // also keeps start/members alive until their copies are done
    for (int g = 0; g < n_groups; ++g) {
        for (int i = 0; i < w->dw.n_obs; ++i) sums[(size_t)g * w->dw.n_obs + i] = host[(size_t)g * ncol + i];
        counts[2 * g] = (int64_t)host[(size_t)g * ncol + w->dw.n_obs];
        counts[2 * g + 1] = (int64_t)host[(size_t)g * ncol + w->dw.n_obs + 1];
    }
    return check_flags(w);
}

int32_t sse_fetch_counters(sse_walkers *w, uint64_t out[SSE_N_COUNTERS], int32_t reset) {
    if (!w || !out) return fail("null argument");
    CU(cudaMemcpyAsync(out, w->dw.counters, SSE_N_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost, w->stream));
    if (reset) CU(cudaMemsetAsync(w->dw.counters, 0, SSE_CNT_ANY_FATAL * sizeof(uint64_t), w->stream));  // the fatal marker stays
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

namespace {

// device op code -> reference OperCode (opercode.jl:43-47)
inline uint64_t to_opercode(const sse_model *m, uint32_t op) {
    const uint32_t bond = op_bond(op), gv = op_gv(op);
    const uint64_t lv = (uint64_t)(gv - m->type_vertex_off[m->bond_type[bond]] + 1);
    const uint64_t vcode = ((op >> 1) & 1u) | (lv << 1);  // VertexCode(diagonal, idx) (opercode.jl:18-21)
    return 1ull | (vcode << 1) | ((uint64_t)(bond + 1) << 26);
}

// Download walker i's string: bits/ranks of its M slots and its n records (unrolled from the ring).
int32_t download_string(sse_walkers *w, int i, const WalkerCtl &c, std::vector<uint2> &words, std::vector<uint4> &rec) {
    const int nw = (c.M + 31) / 32;
    words.resize(nw);
    rec.resize(c.n);
    if (nw) CU(cudaMemcpyAsync(words.data(), w->dw.words + (size_t)i * w->dw.Mw_cap, sizeof(uint2) * nw, cudaMemcpyDeviceToHost, w->stream));
    const uint4 *ring0 = w->dw.rec + (size_t)i * w->dw.R_cap;
    const int64_t first = std::min<int64_t>(c.n, w->dw.R_cap - c.G);
    if (first > 0) CU(cudaMemcpyAsync(rec.data(), ring0 + c.G, sizeof(uint4) * first, cudaMemcpyDeviceToHost, w->stream));
    if (c.n > first) CU(cudaMemcpyAsync(rec.data() + first, ring0, sizeof(uint4) * (c.n - first), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

}  // namespace

// Carlo.write_checkpoint for a range of walkers with ONE round of copies: the control blocks come down first (sizes),
// then every walker's bitmap words, ring segments and state are queued on the stream and awaited together.
int32_t sse_get_states(sse_walkers *w, int32_t first, int32_t count, sse_walker_state *sts) {
    if (!w || !sts) return fail("null argument");
    if (first < 0 || count < 0 || first + count > w->dw.W) return fail("sse_get_states: walker range out of bounds");
    if (count == 0) return 0;
    const sse_model *m = w->model;
    const int N = m->dm.n_sites;
    std::vector<WalkerCtl> ctl(count);
    CU(cudaMemcpyAsync(ctl.data(), w->dw.ctl + first, sizeof(WalkerCtl) * count, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    for (int j = 0; j < count; ++j)  // only the walkers asked for must be between two sweeps
        if (ctl[j].phase)
            return fail("sse_get_state: walker " + std::to_string(first + j) + " is parked inside a sweep (sse_advance); call sse_finish_sweeps first");
    std::vector<std::vector<uint2>> words(count);
    std::vector<std::vector<uint4>> rec(count);
    for (int j = 0; j < count; ++j) {
        const WalkerCtl &c = ctl[j];
        sse_walker_state *st = sts + j;
        st->num_operators = c.n;
        st->rng_draws = c.draws;
        st->avg_worm_length = c.avg_wl;
        st->num_worms = c.num_worms;
        st->T = c.T;
        if (st->operators) {
            if (st->operators_len < c.M) {
                st->operators_len = c.M;
                return fail("sse_get_state: operators buffer too small (walker " + std::to_string(first + j) + ")");
            }
            const int i = first + j, nw = (c.M + 31) / 32;
            words[j].resize(nw);
            rec[j].resize(c.n);
            if (nw) CU(cudaMemcpyAsync(words[j].data(), w->dw.words + (size_t)i * w->dw.Mw_cap, sizeof(uint2) * nw, cudaMemcpyDeviceToHost, w->stream));
            const uint4 *ring0 = w->dw.rec + (size_t)i * w->dw.R_cap;
            const int64_t head = std::min<int64_t>(c.n, w->dw.R_cap - c.G);
            if (head > 0) CU(cudaMemcpyAsync(rec[j].data(), ring0 + c.G, sizeof(uint4) * head, cudaMemcpyDeviceToHost, w->stream));
            if (c.n > head) CU(cudaMemcpyAsync(rec[j].data() + head, ring0, sizeof(uint4) * (c.n - head), cudaMemcpyDeviceToHost, w->stream));
        }
        if (st->state) CU(cudaMemcpyAsync(st->state, w->dw.state + (size_t)(first + j) * N, N, cudaMemcpyDeviceToHost, w->stream));
    }
    CU(cudaStreamSynchronize(w->stream));
    for (int j = 0; j < count; ++j) {
        const WalkerCtl &c = ctl[j];
        sse_walker_state *st = sts + j;
        if (st->operators) {
            int64_t k = 0;
            for (int p = 0; p < c.M; ++p) {
                if (!((words[j][p >> 5].x >> (p & 31)) & 1u)) { st->operators[p] = 0; continue; }
                if (k >= c.n) return fail("sse_get_state: corrupt occupancy bitmap");
                st->operators[p] = to_opercode(m, rec[j][k++].x);
            }
            if (k != c.n) return fail("sse_get_state: occupancy bitmap and operator count disagree");
        }
        st->operators_len = c.M;
    }
    return 0;
}

int32_t sse_get_state(sse_walkers *w, int32_t i, sse_walker_state *st) {
    if (!w || !st) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("sse_get_state: walker index out of range");
    return sse_get_states(w, i, 1, st);
}

namespace {
// host side of sse_set_state(s): the reference's UInt64 operator string -> bitmap words + records (op codes only; the links
// are rebuilt by the next diagonal update / make_vertex_list!), validated against the model's tables
struct StagedState {
    std::vector<uint2> words;
    std::vector<uint4> rec;
    WalkerCtl ctl;
};
int32_t stage_state(const sse_walkers *w, int32_t i, const sse_walker_state *st, StagedState &out) {
    const sse_model *m = w->model;
    const std::string who = "sse_set_state (walker " + std::to_string(i) + "): ";
    if (!st->operators || !st->state) return fail(who + "null operators / state");
    const long long M = st->operators_len;
    if (M < 0 || M > w->dw.M_cap) return fail(who + "operator string longer than m_capacity");
    out.words.assign((size_t)w->dw.Mw_cap, make_uint2(0u, 0u));
    out.rec.clear();
    long long n = 0;
    for (long long p = 0; p < M; ++p) {
        if ((p & 31) == 0) out.words[p >> 5].y = (uint32_t)n;
        uint64_t code = st->operators[p];
        if (code == 0) continue;
        long long bond = (long long)(code >> 26) - 1;
        uint64_t vcode = (code & ((1ull << 25) - 1)) >> 1;  // get_vertex (opercode.jl:61-62)
        long long lv = (long long)(vcode >> 1);
        if (bond < 0 || bond >= m->dm.n_bonds) return fail(who + "bond index out of range at slot " + std::to_string(p));
        int t = m->bond_type[bond];
        long long gv = m->type_vertex_off[t] + lv - 1;
        if (lv < 1 || gv >= m->type_vertex_off[t + 1]) return fail(who + "vertex index out of range at slot " + std::to_string(p));
        if ((uint32_t)(vcode & 1) != m->is_diag[gv]) return fail(who + "diagonal flag inconsistent with the vertex table at slot " + std::to_string(p));
        out.words[p >> 5].x |= 1u << (p & 31);
        out.rec.push_back(make_uint4(op_pack((uint32_t)bond, (uint32_t)gv, (uint32_t)(vcode & 1)), 0u, 0u, 0u));
        ++n;
    }
    if (n != st->num_operators) return fail(who + "num_operators does not match the operator string");
    if (n > w->dw.n_cap) return fail(who + "more operators than n_capacity");
    for (int s = 0; s < m->dm.n_sites; ++s)
        if (st->state[s] < 1 || st->state[s] > m->site_dim[s]) return fail(who + "state index out of range at site " + std::to_string(s));
    WalkerCtl &c = out.ctl;
    memset(&c, 0, sizeof(c));
    c.T = st->T;
    c.num_worms = st->num_worms;
    c.avg_wl = st->avg_worm_length;
    c.last_wlf = NAN;
    c.draws = st->rng_draws;
    c.M = (int)M;
    c.n = (int)n;
    return 0;
}
}  // namespace

int32_t sse_set_states(sse_walkers *w, int32_t first, int32_t count, const sse_walker_state *sts) {
    if (!w || !sts) return fail("null argument");
    if (first < 0 || count < 0 || first + count > w->dw.W) return fail("sse_set_states: walker range out of bounds");
    if (count == 0) return 0;
    const sse_model *m = w->model;
    // everything is validated and staged before the first byte is copied: a bad state leaves the batch untouched
    std::vector<StagedState> staged(count);
    for (int j = 0; j < count; ++j)
        if (int32_t e = stage_state(w, first + j, sts + j, staged[j])) return e;
    CU(cudaStreamSynchronize(w->stream));
    for (int j = 0; j < count; ++j) {
        const int i = first + j;
        const StagedState &g = staged[j];
        CU(cudaMemcpyAsync(w->dw.words + (size_t)i * w->dw.Mw_cap, g.words.data(), sizeof(uint2) * g.words.size(), cudaMemcpyHostToDevice, w->stream));
        if (!g.rec.empty()) CU(cudaMemcpyAsync(w->dw.rec + (size_t)i * w->dw.R_cap, g.rec.data(), sizeof(uint4) * g.rec.size(), cudaMemcpyHostToDevice, w->stream));
        CU(cudaMemcpyAsync(w->dw.state + (size_t)i * m->dm.n_sites, sts[j].state, m->dm.n_sites, cudaMemcpyHostToDevice, w->stream));
        CU(cudaMemcpyAsync(w->dw.ctl + i, &g.ctl, sizeof(WalkerCtl), cudaMemcpyHostToDevice, w->stream));
    }
    CU(cudaStreamSynchronize(w->stream));  // the staging vectors are about to be freed
    w->have_vl = false;
    return 0;
}

int32_t sse_set_state(sse_walkers *w, int32_t i, const sse_walker_state *st) {
    if (!w || !st) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("sse_set_state: walker index out of range");
    return sse_set_states(w, i, 1, st);
}

int32_t sse_get_flags(sse_walkers *w, uint32_t *flags) {
    if (!w || !flags) return fail("null argument");
    std::vector<uint32_t> f;
    if (int32_t s = get_field(w, CTL_OFF(flags), f)) return s;
    std::copy(f.begin(), f.end(), flags);
    return 0;
}

int32_t sse_get_num_operators(sse_walkers *w, int64_t *out) {
    if (!w || !out) return fail("null argument");
    std::vector<int> n;
    if (int32_t s = get_field(w, CTL_OFF(n), n)) return s;
    for (size_t i = 0; i < n.size(); ++i) out[i] = n[i];
    return 0;
}

int32_t sse_pt_log_weight_ratio(sse_walkers *w, const double *new_T, double *out) {
    if (!w || !new_T || !out) return fail("null argument");
    std::vector<int> n;
    std::vector<double> T;
    if (int32_t s = get_field(w, CTL_OFF(n), n)) return s;
    if (int32_t s = get_field(w, CTL_OFF(T), T)) return s;
    for (int i = 0; i < w->dw.W; ++i) out[i] = -(double)n[i] * std::log(new_T[i] / T[i]);  // src/sse.jl:395
    return 0;
}

int32_t sse_set_temperature(sse_walkers *w, const double *T) {
    if (!w || !T) return fail("null argument");
    for (int i = 0; i < w->dw.W; ++i)
        if (!(T[i] > 0)) return fail("sse_set_temperature: temperatures must be positive");
    return set_field(w, CTL_OFF(T), T);  // src/sse.jl:403
}

int32_t sse_get_temperatures(sse_walkers *w, double *T) {
    if (!w || !T) return fail("null argument");
    std::vector<double> t;
    if (int32_t s = get_field(w, CTL_OFF(T), t)) return s;
    std::copy(t.begin(), t.end(), T);
    return 0;
}

int32_t sse_pt_set_ladder(sse_walkers *w, const int32_t *walker_at_rank, int32_t n) {
    if (!w || !walker_at_rank) return fail("null argument");
    if (n < 2 || n > w->dw.W) return fail("sse_pt_set_ladder: need 2 <= n <= n_walkers");
    std::vector<char> seen(w->dw.W, 0);
    for (int i = 0; i < n; ++i) {
        const int x = walker_at_rank[i];
        if (x < 0 || x >= w->dw.W || seen[x]) return fail("sse_pt_set_ladder: walker indices must be distinct and in range");
        seen[x] = 1;
    }
    CU(cudaSetDevice(w->model->device));
    CU(cudaStreamSynchronize(w->stream));
    if (w->d_ladder) { cudaFree(w->d_ladder); w->d_ladder = nullptr; }
    CU(cudaMalloc((void **)&w->d_ladder, sizeof(int32_t) * ((size_t)n + 1)));  // [n] = accept counter
    CU(cudaMemcpyAsync(w->d_ladder, walker_at_rank, sizeof(int32_t) * n, cudaMemcpyHostToDevice, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    w->n_ladder = n;
    return 0;
}

int32_t sse_pt_get_ladder(sse_walkers *w, int32_t *walker_at_rank) {
    if (!w || !walker_at_rank) return fail("null argument");
    if (!w->d_ladder) return fail("sse_pt_get_ladder: no ladder set");
    CU(cudaMemcpyAsync(walker_at_rank, w->d_ladder, sizeof(int32_t) * w->n_ladder, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    return 0;
}

int32_t sse_pt_exchange(sse_walkers *w, int32_t parity, uint64_t seed, uint64_t step, int32_t *n_accepted) {
    if (!w) return fail("null handle");
    if (!w->d_ladder) return fail("sse_pt_exchange: call sse_pt_set_ladder first");
    if (parity != 0 && parity != 1) return fail("sse_pt_exchange: parity must be 0 or 1");
    if (int32_t s = require_between_sweeps(w, "sse_pt_exchange")) return s;
    CU(cudaSetDevice(w->model->device));
    int32_t *cnt = w->d_ladder + w->n_ladder;
    CU(cudaMemsetAsync(cnt, 0, sizeof(int32_t), w->stream));
    const int pairs = (w->n_ladder - parity) / 2;
    if (pairs > 0) {
        SSE_LAUNCH_KERNEL(k_pt_exchange, (pairs + 127) / 128, 128, 0, w->stream, w->dw, w->d_ladder, w->n_ladder, (int)parity,
                          (unsigned long long)seed, (unsigned long long)step, cnt);
        CU(cudaGetLastError());
    }
    if (n_accepted) {
        CU(cudaMemcpyAsync(n_accepted, cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, w->stream));
        CU(cudaStreamSynchronize(w->stream));
    }
    return 0;
}

int32_t sse_pt_uniforms(uint64_t seed, uint64_t step, int32_t n, double *out) {
    if (!out || n < 0) return fail("sse_pt_uniforms: bad argument");
    for (int i = 0; i < n; ++i) out[i] = sse_u01(sse_philox_draw(seed, step, (uint64_t)i));
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
    if (int32_t s = require_between_sweeps(w, "sse_double_beta")) return s;
    CU(cudaSetDevice(w->model->device));
    const int grid = (w->dw.W + PHASE_WARPS - 1) / PHASE_WARPS;
    SSE_LAUNCH_KERNEL(k_double_beta, grid, PHASE_WARPS * 32, 0, w->stream, w->dw);
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
        CU(cudaMemcpyAsync(w->d_inj, stream, bytes, cudaMemcpyHostToDevice, w->stream));
        w->dw.inj = w->d_inj;
        w->dw.inj_len = len;
        std::vector<unsigned long long> zero(w->dw.W, 0ull);
        if (int32_t s = set_field(w, CTL_OFF(draws), zero.data())) return s;  // synchronises the stream
    }
    return 0;
}

int32_t sse_dbg_diagonal_update(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = require_between_sweeps(w, "sse_dbg_diagonal_update")) return s;
    PhaseArgs a{};
    a.mode = MODE_DIAG;
    w->have_vl = false;
    if (int32_t s = launch_phase(w, a)) return s;
    return sse_sync(w);
}

int32_t sse_dbg_make_vertex_list(sse_walkers *w) {
    if (!w) return fail("null handle");
    if (int32_t s = require_between_sweeps(w, "sse_dbg_make_vertex_list")) return s;
    PhaseArgs a{};
    a.mode = MODE_MAKE_VL;
    if (int32_t s = launch_phase(w, a)) return s;
    w->have_vl = true;
    return sse_sync(w);
}

int32_t sse_dbg_worm_update(sse_walkers *w, int32_t thermalized) {
    if (!w) return fail("null handle");
    if (!w->have_vl) return fail("sse_dbg_worm_update: call sse_dbg_make_vertex_list first");
    PhaseArgs a{};
    a.mode = MODE_WORM_UPDATE;
    a.thermalized = thermalized;
    if (int32_t s = launch_phase(w, a)) return s;
    return sse_sync(w);
}

int32_t sse_dbg_worm_traverse(sse_walkers *w, int32_t l0, int64_t p0, int32_t wormfunc0, int64_t *lengths) {
    if (!w || !lengths) return fail("null argument");
    if (!w->have_vl) return fail("sse_dbg_worm_traverse: call sse_dbg_make_vertex_list first");
    if (l0 < 1 || l0 > 4 || p0 < 1 || wormfunc0 < 1) return fail("sse_dbg_worm_traverse: start out of range");
    std::vector<int> M;
    if (int32_t s = get_field(w, CTL_OFF(M), M)) return s;
    for (int m : M)
        if (p0 > m) return fail("sse_dbg_worm_traverse: p0 beyond the operator string");
    PhaseArgs a{};
    a.mode = MODE_WORM_TRAVERSE;
    a.l0 = l0 - 1;
    a.p0 = p0 - 1;
    a.w0 = wormfunc0;
    if (int32_t s = launch_phase(w, a)) return s;
    if (int32_t s = sse_sync(w)) return s;
    std::vector<long long> len(w->dw.W);
    CU(cudaMemcpyAsync(len.data(), w->dw.dbg_len, sizeof(long long) * len.size(), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    for (size_t i = 0; i < len.size(); ++i) lengths[i] = len[i];
    return 0;
}

int32_t sse_dbg_get_vertex_list(sse_walkers *w, int32_t i, int64_t *vertices, int64_t m_len, int64_t *v_first, int64_t *v_last) {
    if (!w || !vertices || !v_first || !v_last) return fail("null argument");
    if (i < 0 || i >= w->dw.W) return fail("walker index out of range");
    if (!w->have_vl) return fail("sse_dbg_get_vertex_list: no vertex list (call sse_dbg_make_vertex_list)");
    CU(cudaStreamSynchronize(w->stream));
    const int N = w->model->dm.n_sites;
    WalkerCtl c;
    CU(cudaMemcpyAsync(&c, w->dw.ctl + i, sizeof(c), cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    const int M = c.M, n = c.n;
    if (m_len < M) return fail("sse_dbg_get_vertex_list: vertices buffer too small");
    std::vector<uint2> words;
    std::vector<uint4> rec;
    if (int32_t s = download_string(w, i, c, words, rec)) return s;
    std::vector<uint32_t> vf(N), vl(N);
    CU(cudaMemcpyAsync(vf.data(), w->dw.vfirst + (size_t)i * N, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaMemcpyAsync(vl.data(), w->dw.vlast + (size_t)i * N, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, w->stream));
    CU(cudaStreamSynchronize(w->stream));
    std::vector<int64_t> pos(n, -1), rec_of(M, -1);
    int64_t k = 0;
    for (int p = 0; p < M; ++p)
        if ((words[p >> 5].x >> (p & 31)) & 1u) {
            if (k >= n) return fail("sse_dbg_get_vertex_list: corrupt occupancy bitmap");
            if (words[p >> 5].y + (uint32_t)__builtin_popcount(words[p >> 5].x & ((1u << (p & 31)) - 1u)) != (uint32_t)k)
                return fail("sse_dbg_get_vertex_list: corrupt rank");
            pos[k] = p;
            rec_of[p] = k++;
        }
    auto link = [&](const uint4 &r, int j) -> uint32_t {
        const uint64_t lo = (uint64_t)r.y | ((uint64_t)r.z << 32);
        switch (j) {
            case 0: return (uint32_t)(lo & NONE24);
            case 1: return (uint32_t)((lo >> 24) & NONE24);
            case 2: return (uint32_t)(((lo >> 48) | ((uint64_t)r.w << 16)) & NONE24);
            default: return r.w >> 8;
        }
    };
    for (int64_t p = 0; p < M; ++p)
        for (int l = 0; l < 4; ++l) {
            int64_t *dst = vertices + (p * 4 + l) * 2;
            dst[0] = dst[1] = -1;
            if (rec_of[p] < 0) continue;
            uint32_t lk = link(rec[rec_of[p]], l);
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

}  // extern "C"
