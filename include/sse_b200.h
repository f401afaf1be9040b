/* sse_b200.h — C ABI of libsse_b200.so, the B200 (sm_100a) sweep backend for
 * lukas-weber/StochasticSeriesExpansion.jl.
 *
 * This is the drop-in boundary: every entry point replaces one method the reference
 * implements in Julia on `MC <: Carlo.AbstractMC` (src/sse.jl) for a BATCH of independent walkers.
 * All functions are `extern "C"`, take plain pointers and sizes, return an int32 status (0 = ok) and
 * never throw; `sse_last_error()` returns a message for the last failure on the calling thread.
 *
 * Ownership: the caller owns every host buffer it passes (the library copies in/out before
 * returning; a Julia caller wraps calls in GC.@preserve).  The library owns all device memory behind
 * the opaque handles, freed by sse_*_destroy.  A handle is not re-entrant.  Work is enqueued on the
 * handle's CUDA stream (own stream by default, or the caller's via sse_set_stream); functions that
 * return data synchronise that stream before returning, sse_sweep does not (pair it with sse_sync).
 *
 * Index conventions at this boundary: sites, bonds and bond types are 0-based; vertex indices are
 * LOCAL to their bond type and 1-based (0 = invalid); leg indices 0..3 (0,1 = bottom legs of site
 * 0,1 of the bond; 2,3 = top legs); worm indices and state indices are 1-based exactly as in the
 * reference tables.  Operator strings cross the boundary in the reference's UInt64 `OperCode`
 * layout (src/opercode.jl:43-47: bit0 = non-identity, bits 1..25 = VertexCode = diagonal flag |
 * (1-based local vertex idx << 1), bits 26.. = 1-based bond idx), so reference checkpoints load.
 */
#ifndef SSE_B200_H
#define SSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSE_B200_ABI_VERSION 2

/* Flattened `SSEData` + `VertexData` tables + estimator tables.
 * Replaces: SSEData{NSites} (src/sse_data.jl:15-22), VertexData{NSites} (src/vertex_data.jl:13-28),
 * and the MagnetizationEstimator type parameters (src/models/common/magnetization_estimator.jl:32-46). */
typedef struct sse_model_desc {
    int32_t n_sites;
    const uint8_t *site_dim;        /* [n_sites] local Hilbert-space dimension */
    int32_t n_bonds;
    const int32_t *bond_type;       /* [n_bonds] */
    const int32_t *bond_sites;      /* [n_bonds*2] */
    int32_t n_types;
    const int32_t *type_dims;       /* [n_types*2] dims of the two sites of a bond of this type */
    const int32_t *type_vertex_off; /* [n_types+1] prefix offsets into the per-vertex arrays */
    const int32_t *type_diag_off;   /* [n_types+1] prefix offsets into diag_vertices */
    int32_t n_vertices;             /* = type_vertex_off[n_types] */
    const double *weights;          /* [n_vertices] */
    const int8_t *signs;            /* [n_vertices] +1/-1 */
    const uint8_t *leg_states;      /* [n_vertices*4] 1-based states, leg fastest */
    const int32_t *diag_vertices;   /* [type_diag_off[n_types]] local 1-based vertex idx of the diagonal vertex
                                       for compound state c = (s0-1) + dim0*(s1-1); 0 = no such vertex */
    int32_t max_worm;               /* max over sites of dim-1 */
    const int32_t *trans_offset;    /* [n_vertices*max_worm*4], index ((v*max_worm + worm-1)*4 + leg_in):
                                       offset of the first outcome, -1 = invalid transition */
    const int32_t *trans_count;     /* same shape: number of outcomes (>= 1 when valid) */
    int32_t n_outcomes;
    const double *out_cumprob;      /* [n_outcomes] cumulative probabilities */
    const int32_t *out_target;      /* [n_outcomes] local 1-based vertex idx of the vertex after the step */
    const int32_t *out_leg;         /* [n_outcomes] exit leg 0..3 */
    const int32_t *out_worm;        /* [n_outcomes] exit worm 1..dim-1 */
    double energy_offset;           /* SSEData.energy_offset = sum over bonds (src/sse_data.jl:38) */
    int32_t norm_site_count;        /* normalization_site_count(model) (src/abstract_model.jl:44-47) */
    int32_t n_estimators;           /* table-driven magnetization estimators */
    int32_t est_max_dim;            /* row length of est_values */
    const double *est_values;       /* [n_estimators][n_sites][est_max_dim]:
                                       staggered_sign(site) * magnetization_state(site, state), state-1 = column */
} sse_model_desc;

typedef struct sse_model sse_model;
typedef struct sse_walkers sse_walkers;

/* Parameters of `MC(params)` (src/sse.jl:26-45) for a batch of walkers. */
typedef struct sse_walkers_opts {
    int32_t n_walkers;
    const double *T;                    /* [n_walkers] temperature per walker (params[:T]) */
    int64_t m_capacity;                 /* capacity of each operator string (slots); M grows inside it (costs 0.25 B/slot) */
    int64_t n_capacity;                 /* max non-identity operators per walker, <= 2^22 - 1 (costs 17 B each) */
    int32_t device;                     /* CUDA device ordinal; -1 = current device */
    uint64_t seed;                      /* Philox key */
    uint64_t walker_id_offset;          /* global id of walker 0 (stream id = offset + index) */
    double target_worm_length_fraction; /* default 2.0  (src/sse.jl:34) */
    double num_worms_attenuation_factor;/* default 0.01 (src/sse.jl:35) */
    double init_num_worms;              /* default 5    (src/sse.jl:37) */
} sse_walkers_opts;

/* The five checkpointed fields of `MC` (src/sse.jl:89-107) + the stream position. */
typedef struct sse_walker_state {
    int64_t num_operators;
    double avg_worm_length;
    double num_worms;
    uint64_t *operators;     /* reference-format OperCodes; in: buffer, out: filled */
    int64_t operators_len;   /* get: in = buffer capacity, out = M;  set: M */
    uint8_t *state;          /* [n_sites] 1-based state indices */
    uint64_t rng_draws;      /* number of draws consumed so far (stream position) */
    double T;
} sse_walker_state;

/* Observable layout of sse_measure / sse_fetch_accumulators, per walker:
 *   0 Sign  1 OperatorCount  2 SignOperatorCount  3 SignOperatorCount2  4 SignEnergy
 *   5 WormLengthFraction  then for estimator e: 6+5e + {0 SignMag, 1 SignAbsMag, 2 SignMag2, 3 SignMag4, 4 SignMagChi} */
#define SSE_OBS_SIGN 0
#define SSE_OBS_OPERATOR_COUNT 1
#define SSE_OBS_SIGN_OPERATOR_COUNT 2
#define SSE_OBS_SIGN_OPERATOR_COUNT2 3
#define SSE_OBS_SIGN_ENERGY 4
#define SSE_OBS_WORM_LENGTH_FRACTION 5
#define SSE_OBS_FIXED 6
#define SSE_OBS_PER_ESTIMATOR 5

/* Per-walker error flags (sse_get_flags). */
#define SSE_FLAG_M_OVERFLOW 1u      /* operator string would grow beyond m_capacity */
#define SSE_FLAG_N_OVERFLOW 2u      /* more non-identity operators than n_capacity */
#define SSE_FLAG_STREAM_EXHAUSTED 4u/* injected random stream ran out */
#define SSE_FLAG_SCATTER_FALLTHROUGH 8u /* r >= last cumprob (src/vertex_data.jl:124): clamped to last outcome */

const char *sse_last_error(void);
int32_t sse_abi_version(void);

/* --- model: replaces generate_sse_data(model)::SSEData consumption (src/sse.jl:28, src/sse_data.jl:70-71) --- */
int32_t sse_model_create(const sse_model_desc *desc, sse_model **out);
int32_t sse_model_destroy(sse_model *m);

/* --- walkers: replaces MC(params) (src/sse.jl:26-45) --- */
int32_t sse_walkers_create(const sse_model *m, const sse_walkers_opts *opts, sse_walkers **out);
int32_t sse_walkers_destroy(sse_walkers *w);
int32_t sse_set_stream(sse_walkers *w, void *cuda_stream);
int32_t sse_n_observables(const sse_walkers *w);
/* Device bytes one walker of `m` costs at these capacities (0.25 B per string slot + 17 B per operator + 9 B per site +
 * accumulators): for sizing a batch to the GPU's memory. */
int64_t sse_walker_bytes(const sse_model *m, int64_t m_capacity, int64_t n_capacity);
int64_t sse_device_bytes(const sse_walkers *w);

/* The reference resizes its operator string in place (src/sse.jl:138-145); the device arrays have fixed capacities.
 * sse_grow_capacity moves every walker to larger ones (old and new arrays coexist during the move) and clears a pending
 * "string outgrew m_capacity" condition: that overflow is detected before the sweep modifies anything, so the walker
 * simply continues with its next sse_sweep.  Running out of n_capacity in the middle of a diagonal update is NOT
 * recoverable (the record ring is partly rewritten): grow n_capacity while sse_get_num_operators is still below it.
 * Needs walkers between sweeps. */
int32_t sse_grow_capacity(sse_walkers *w, int64_t m_capacity, int64_t n_capacity);

/* Carlo.init!(mc, ctx, params) (src/sse.jl:47-60): random state, `init_opstring_cutoff` identities
 * (< 0: round(n_sites*T) per walker), `diagonal_warmup_sweeps` diagonal updates. Synchronous. */
int32_t sse_init(sse_walkers *w, int64_t init_opstring_cutoff, int32_t diagonal_warmup_sweeps);

/* Carlo.sweep!(mc, ctx) (src/sse.jl:62-68) x n_sweeps for every walker inside ONE persistent launch:
 * diagonal_update -> make_vertex_list! -> worm_update.  `thermalized` = is_thermalized(ctx) (src/sse.jl:139,200,205).
 * If `measure` != 0, Carlo.measure! (src/sse.jl:70-87) runs on the device after every sweep and its
 * observables are added to the per-walker accumulators.  Asynchronous on the handle's stream.
 * Walkers that sse_advance left in the middle of a sweep finish that sweep first (it counts as one of the n_sweeps). */
int32_t sse_sweep(sse_walkers *w, int32_t n_sweeps, int32_t thermalized, int32_t measure);
int32_t sse_sync(sse_walkers *w);

/* Free-running form of sse_sweep, NOT in the reference: every walker keeps sweeping until it has completed
 * `max_sweeps` sweeps or done `visit_budget` worm visits in this call, whichever comes first, and is then parked
 * wherever it is — between sweeps, between two worms or in the middle of a worm; the next sse_advance / sse_sweep
 * resumes there.  Each walker still runs exactly the reference's Markov chain (results do not depend on where it was
 * parked); what changes is that walkers are not kept in step: a launch does the same amount of worm work for every
 * walker, and one walker inside a very long worm delays nobody else.  Asynchronous on the handle's stream.
 * sse_get_state, sse_measure, sse_double_beta and the parity hooks need walkers BETWEEN sweeps and fail otherwise
 * (sse_finish_sweeps completes the sweeps in flight). */
int32_t sse_advance(sse_walkers *w, int32_t max_sweeps, uint64_t visit_budget, int32_t thermalized, int32_t measure);
/* Complete the sweeps that sse_advance left in flight (no new sweep is started).  Asynchronous. */
int32_t sse_finish_sweeps(sse_walkers *w, int32_t thermalized, int32_t measure);
/* Resume an sse_sweep that ended with an error some walkers can recover from (sse_grow_capacity after "string outgrew
 * m_capacity"): every walker does the sweeps of that call it has not done yet, so the batch is in step again.  Asynchronous. */
int32_t sse_continue_sweeps(sse_walkers *w, int32_t thermalized, int32_t measure);
/* sweeps_done[n_walkers]: completed sweeps since sse_init / sse_set_state; in_flight[n_walkers] (may be NULL): 1 if
 * the walker is parked inside a sweep. */
int32_t sse_get_progress(sse_walkers *w, uint64_t *sweeps_done, uint8_t *in_flight);

/* Carlo.measure!(mc, ctx) (src/sse.jl:70-87, 305-376) on the current configuration of every walker:
 * out[n_walkers][n_obs] (host).  Slot 5 holds the last sweep's WormLengthFraction (NaN if none). */
int32_t sse_measure(sse_walkers *w, double *out);

/* Per-walker sums accumulated by sse_sweep(measure=1): sums[n_walkers][n_obs], counts[n_walkers][2]
 * (measurements, WormLengthFraction measurements).  reset != 0 zeroes them afterwards (one "bin"). */
int32_t sse_fetch_accumulators(sse_walkers *w, double *sums, int64_t *counts, int32_t reset);
/* One bin reduced inside the library: the accumulators of the walkers of each group (group[i] in
 * 0..n_groups-1, e.g. the index of the walker's temperature; NULL = one group) are summed on the device in walker
 * order, all-reduced over the ranks of sse_comm_init with NCCL on the handle's stream (skipped without a communicator),
 * and returned: sums[n_groups][n_obs], counts[n_groups][2].  Every rank must call it with the same n_groups.
 * Replaces Carlo's MPI merge of `measure!` accumulators (src/sse.jl:73-83,201; magnetization_estimator.jl:218-227)
 * for a host without a GPU-aware reduction of its own.  NCCL (libnccl.so.2) is resolved at run time. */
typedef struct sse_nccl_id { char internal[128]; } sse_nccl_id;   /* = ncclUniqueId */
int32_t sse_comm_unique_id(sse_nccl_id *id);                       /* rank 0 creates it, the host broadcasts it (MPI, ...) */
int32_t sse_comm_init(sse_walkers *w, const sse_nccl_id *id, int32_t rank, int32_t nranks);
int32_t sse_reduce_bins(sse_walkers *w, const int32_t *group, int32_t n_groups, double *sums, int64_t *counts, int32_t reset);
/* Device pointers of the accumulator buffers, for reductions by a host that has its own collective (no copy). */
int32_t sse_accumulators_device_ptr(sse_walkers *w, void **sums, void **counts);

/* Totals since creation or the last reset (sse_fetch_counters):
 *   visits      worm-vertex visits = sum of the lengths returned by worm_traverse! (src/sse.jl:302)
 *   sweeps      completed walker-sweeps;  sum_n / sum_M: operators / string slots summed over them
 *   cyc_build / cyc_finish / cyc_idle   SM cycles of the stream warps in diagonal update + record build / end of
 *               worm_update + measure / waiting for work;   tasks: streaming tasks
 *   cyc_worm    SM cycles of the worm warps;  lane_iters: visits they executed;  warp_iters: loop iterations they
 *               issued (lane_iters / (32 * warp_iters) = share of lanes with a walker to chase) */
#define SSE_CNT_VISITS 0
#define SSE_CNT_SWEEPS 1
#define SSE_CNT_SUM_N 2
#define SSE_CNT_SUM_M 3
#define SSE_CNT_CYC_BUILD 4
#define SSE_CNT_CYC_WORM 5
#define SSE_CNT_CYC_FINISH 6
#define SSE_CNT_CYC_IDLE 7
#define SSE_CNT_LANE_ITERS 8
#define SSE_CNT_WARP_ITERS 9
#define SSE_CNT_TASKS 10
#define SSE_CNT_ANY_FATAL 15 /* internal: some walker raised a fatal flag */
#define SSE_N_COUNTERS 16
int32_t sse_fetch_counters(sse_walkers *w, uint64_t out[SSE_N_COUNTERS], int32_t reset);

/* Carlo.write_checkpoint / read_checkpoint (src/sse.jl:89-107). */
int32_t sse_get_state(sse_walkers *w, int32_t walker, sse_walker_state *st);
/* The same for walkers first .. first+count-1 with one round of device-to-host copies (states[count]); a call with
 * NULL buffers fills the sizes (operators_len = M) and the scalars only. */
int32_t sse_get_states(sse_walkers *w, int32_t first, int32_t count, sse_walker_state *states);
int32_t sse_set_state(sse_walkers *w, int32_t walker, const sse_walker_state *st);
/* Carlo.read_checkpoint for walkers first .. first+count-1 in one round of host-to-device copies.  All states are validated
 * against the model's tables before anything is copied: an invalid state fails the call and leaves the batch untouched. */
int32_t sse_set_states(sse_walkers *w, int32_t first, int32_t count, const sse_walker_state *states);
int32_t sse_get_flags(sse_walkers *w, uint32_t *flags /* [n_walkers] */);

/* Carlo.parallel_tempering_log_weight_ratio / _change_parameter! (src/sse.jl:390-405), parameter :T only. */
int32_t sse_pt_log_weight_ratio(sse_walkers *w, const double *new_T, double *out /* [n_walkers] */);
int32_t sse_set_temperature(sse_walkers *w, const double *T /* [n_walkers] */);
int32_t sse_get_num_operators(sse_walkers *w, int64_t *out /* [n_walkers] */);
int32_t sse_get_temperatures(sse_walkers *w, double *T /* [n_walkers] */);

/* Replica exchange decided ON THE DEVICE with the two hooks above (what Carlo's parallel-tempering wrapper does with
 * them over MPI).  sse_pt_set_ladder names the walkers of a temperature ladder in rank order (walker_at_rank[n]);
 * sse_pt_exchange proposes the neighbour swaps (rank r, r+1), r = parity, parity+2, ...: pair i is accepted iff
 * log(u_i) < lw_a + lw_b with lw_x = -n_x * log(T_other / T_x) (src/sse.jl:395) and u_i = draw i of the Philox stream
 * (seed, step) — sse_pt_uniforms returns the same numbers to a host that wants to replay the decisions.  Accepted
 * pairs exchange their temperatures (src/sse.jl:398-405) and their places on the ladder; configurations never move.
 * Needs walkers between sweeps.  n_accepted may be NULL (then the call is asynchronous). */
int32_t sse_pt_set_ladder(sse_walkers *w, const int32_t *walker_at_rank, int32_t n);
int32_t sse_pt_get_ladder(sse_walkers *w, int32_t *walker_at_rank);
int32_t sse_pt_exchange(sse_walkers *w, int32_t parity, uint64_t seed, uint64_t step, int32_t *n_accepted);
int32_t sse_pt_uniforms(uint64_t seed, uint64_t step, int32_t n, double *out);

/* Launch shape of sse_sweep / sse_advance, NOT in the reference: warps per CTA that chase worms (one lane = one
 * walker) and warps that run the streaming phases (one warp = one walker); one CTA per SM.  0 = choose automatically
 * from the number of walkers (default; environment SSE_B200_WORM_WARPS / SSE_B200_STREAM_WARPS at creation).
 * worm_warps + stream_warps <= 16.  Results do not depend on this setting (bit-identical). */
int32_t sse_set_launch_shape(sse_walkers *w, int32_t worm_warps, int32_t stream_warps);

/* The two parameters of the worm-count controller (src/sse.jl:34-35,204-217), changeable between launches.
 * The reference fixes them in MC(params); a larger attenuation factor during beta doubling lets the controller
 * follow the quickly growing worm length (see sse_double_beta). */
int32_t sse_set_controller(sse_walkers *w, double target_worm_length_fraction, double num_worms_attenuation_factor);

/* Thermalisation aid, NOT in the reference (beta doubling): every walker's periodic configuration
 * (state, S_M) becomes (state, S_M S_M) at temperature T/2 with 2n operators — a valid configuration at the
 * doubled inverse temperature that is already close to equilibrium, so a cold walker is grown from a cheap
 * hot one in log2(beta) steps.  The controller's average worm length doubles as well.  Fails loudly (overflow flag) if 2M > m_capacity or 2n > n_capacity.
 * The caller keeps sweeping with thermalized = 0 afterwards; nothing here touches the random stream. */
int32_t sse_double_beta(sse_walkers *w);

/* --- parity hooks: run ONE phase on the current configuration with an injected random stream --- */
/* stream[n_walkers][len]: walker i draws stream[i*len + k]; the stream position restarts at 0.  NULL clears. */
int32_t sse_set_injected_stream(sse_walkers *w, const uint64_t *stream, int64_t len);
int32_t sse_dbg_diagonal_update(sse_walkers *w);                  /* src/sse.jl:137-191 */
int32_t sse_dbg_make_vertex_list(sse_walkers *w);                 /* src/vertex_list.jl:15-54 */
int32_t sse_dbg_worm_update(sse_walkers *w, int32_t thermalized); /* src/sse.jl:193-231; needs a vertex list */
/* worm_traverse!((l0, p0, wormfunc0), ...) (src/sse.jl:262-303) with 1-based l0, p0, wormfunc0 as in the
 * reference test (test/test_sse.jl:50); lengths[n_walkers] receives the returned worm length. */
int32_t sse_dbg_worm_traverse(sse_walkers *w, int32_t l0, int64_t p0, int32_t wormfunc0, int64_t *lengths);
/* The vertex list in the reference's layout: vertices[M][4][2] = (leg, p) 1-based or (-1,-1);
 * v_first/v_last[n_sites][2].  Valid after sse_dbg_make_vertex_list / before the next sweep. */
int32_t sse_dbg_get_vertex_list(sse_walkers *w, int32_t walker, int64_t *vertices, int64_t m_len,
                                int64_t *v_first, int64_t *v_last);

#ifdef __cplusplus
}
#endif
#endif /* SSE_B200_H */
