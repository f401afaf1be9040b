// sse_oracle.cpp — CPU ORACLE for the SSE sweep hot path.  TEST INFRASTRUCTURE ONLY.
//
// A line-faithful C++17 restatement of the reference's per-walker sweep, keeping the reference's own
// data layout (UInt64 op codes, 16-byte (leg, p) link tuples re-filled every sweep, Float64 tables,
// 1-based indices) so that it doubles as the timed CPU baseline.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product (libsse_b200.so) never does.
//
// PARITY STATUS: "parity unpinned" at the RNG boundary.  Julia is not installed in this image, so the
// reference itself cannot be run; the oracle is pinned against every golden vector the reference's
// tests hold for this path (tests/test_oracle_golden.py): test/test_vertex_list.jl:5-29 (exact link
// arrays), test/test_sse.jl:5-94 (isconsistent after worm_traverse! and after 1000 sweeps),
// test/test_ed_compare.jl (exact diagonalisation, redone in numpy) and docs/src/bani2v2o8.results.json
// (40 published observable points).  The random stream (Julia's Xoshiro in the reference) is replaced by
// the draw-index contract of include/sse_rng.h; the one libm call of the path, tanh in the worm-count
// controller (src/sse.jl:213), is replaced by sse_tanh so host and device agree bit for bit.
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

#include "../include/sse_b200.h"
#include "../include/sse_rng.h"

namespace {

using Int = int64_t;
using StateIndex = uint8_t;                       // src/worms.jl:2
using OperCode = uint64_t;                        // src/opercode.jl:1,38-40
using VertexCode = uint64_t;                      // src/opercode.jl:11-13
constexpr int vertex_code_maxbits = 8 * 3 + 1;    // src/opercode.jl:2
constexpr VertexCode VERTEX_INVALID = (VertexCode(1) << vertex_code_maxbits) + 1;  // opercode.jl:16

// --- src/opercode.jl:18-71 ---------------------------------------------------------------------
inline VertexCode make_vertex_code(bool diagonal, Int vertex_idx) { return VertexCode(diagonal) | (VertexCode(vertex_idx) << 1); }
inline bool isdiagonal(VertexCode v) { return v & 1; }
inline bool isinvalid(VertexCode v) { return v >= (VertexCode(1) << vertex_code_maxbits); }
inline Int get_vertex_idx(VertexCode v) { return Int(v >> 1); }
inline OperCode make_opercode(Int bond, VertexCode v) { return 1 | (v << 1) | (OperCode(bond) << (1 + vertex_code_maxbits)); }
inline Int get_bond(OperCode o) { return Int(o >> (1 + vertex_code_maxbits)); }
inline VertexCode get_vertex(OperCode o) { return (o & ((OperCode(1) << vertex_code_maxbits) - 1)) >> 1; }
inline bool isidentity(OperCode o) { return o == 0; }
inline bool op_isdiagonal(OperCode o) { return isdiagonal(get_vertex(o)); }

// --- src/worms.jl:4-7, src/vertex_data.jl:190 -----------------------------------------------------
inline Int worm_inverse(Int worm, Int basis_size) { return basis_size - worm; }
inline Int worm_count(Int basis_size) { return basis_size - 1; }
inline Int site_of_leg(Int leg, Int num_sites) { return leg > num_sites ? leg - num_sites : leg; }

constexpr Int NSites = 2;
constexpr Int LegCount = 2 * NSites;

struct Transition { Int offset = -1; Int length = 0; };  // src/vertex_data.jl:6-11

// src/vertex_data.jl:13-28 (1-based semantics kept; arrays padded with a dummy element 0)
struct VertexData {
    double energy_offset = 0;
    Int dims[NSites];
    std::vector<VertexCode> diagonal_vertices;  // [1..prod(dims)]
    std::vector<int8_t> signs;                   // [1..nv]
    std::vector<double> weights;                 // [1..nv]
    Int max_worm = 1, nv = 0;
    std::vector<Transition> transitions;         // [leg_in, worm_in, vertex] column-major, 1-based
    std::vector<double> transition_cumprobs;     // [1..]
    std::vector<VertexCode> transition_targets;
    std::vector<std::pair<Int, Int>> transition_step_outs;  // (leg, worm)
    std::vector<StateIndex> leg_states;          // [leg, vertex]

    const Transition &trans(Int leg_in, Int worm_in, Int vi) const {
        return transitions[(leg_in - 1) + LegCount * ((worm_in - 1) + max_worm * (vi - 1))];
    }
    // src/vertex_data.jl:92-104
    VertexCode get_diagonal_vertex(Int compound_state_idx) const { return diagonal_vertices[compound_state_idx]; }
    double get_vertex_weight(VertexCode v) const { return isinvalid(v) ? 0.0 : weights[get_vertex_idx(v)]; }
    int get_sign(VertexCode v) const { return signs[get_vertex_idx(v)]; }
    const StateIndex *get_leg_state(VertexCode v) const { return &leg_states[LegCount * (get_vertex_idx(v) - 1)] - 1; }  // 1-based [leg]
};

struct SSESite { Int dim; };                      // src/sse_data.jl:1-3
struct SSEBond { Int type; Int sites[NSites]; };  // src/sse_data.jl:9-13 (1-based)

struct SSEData {                                  // src/sse_data.jl:15-22
    std::vector<VertexData> vertex_data;          // [1..]
    std::vector<SSESite> sites;                   // [1..]
    std::vector<SSEBond> bonds;                   // [1..]
    double energy_offset = 0;
    Int norm_site_count = 1;
    Int n_estimators = 0, est_max_dim = 0;
    std::vector<double> est_values;               // [e][site0][state0]
    Int nsites() const { return Int(sites.size()) - 1; }
    Int nbonds() const { return Int(bonds.size()) - 1; }
    const VertexData &get_vertex_data(Int bond_idx) const { return vertex_data[bonds[bond_idx].type]; }  // sse_data.jl:70-71
    double est(Int e, Int site, Int state) const { return est_values[(e * nsites() + (site - 1)) * est_max_dim + (state - 1)]; }
};

// src/vertex_data.jl:106-125
struct ScatterResult { Int leg_out, worm_out; VertexCode target; };
inline ScatterResult scatter(const VertexData &vd, VertexCode v, Int leg_in, Int worm_in, double random, bool *fell_through) {
    Int vi = get_vertex_idx(v);
    const Transition &t = vd.trans(leg_in, worm_in, vi);
    for (Int out = t.offset; out <= t.offset + t.length; ++out) {
        if (random < vd.transition_cumprobs[out]) {
            return {vd.transition_step_outs[out].first, vd.transition_step_outs[out].second, vd.transition_targets[out]};
        }
    }
    // The reference returns (-1, -1, invalid) here (vertex_data.jl:124), which is unusable by its caller;
    // probability ~1e-8 per visit.  Like the device path, clamp to the last outcome and flag it.
    *fell_through = true;
    Int out = t.offset + t.length;
    return {vd.transition_step_outs[out].first, vd.transition_step_outs[out].second, vd.transition_targets[out]};
}

SSEData *build_sse_data(const sse_model_desc *d) {
    auto *s = new SSEData();
    s->sites.resize(d->n_sites + 1);
    for (Int i = 0; i < d->n_sites; ++i) s->sites[i + 1].dim = d->site_dim[i];
    s->bonds.resize(d->n_bonds + 1);
    for (Int b = 0; b < d->n_bonds; ++b) {
        s->bonds[b + 1].type = d->bond_type[b] + 1;
        for (Int k = 0; k < NSites; ++k) s->bonds[b + 1].sites[k] = d->bond_sites[b * NSites + k] + 1;
    }
    s->vertex_data.resize(d->n_types + 1);
    for (Int t = 0; t < d->n_types; ++t) {
        VertexData &vd = s->vertex_data[t + 1];
        vd.dims[0] = d->type_dims[2 * t];
        vd.dims[1] = d->type_dims[2 * t + 1];
        Int v0 = d->type_vertex_off[t], v1 = d->type_vertex_off[t + 1];
        vd.nv = v1 - v0;
        vd.max_worm = d->max_worm;
        Int nd = d->type_diag_off[t + 1] - d->type_diag_off[t];
        vd.diagonal_vertices.assign(nd + 1, VERTEX_INVALID);
        for (Int c = 0; c < nd; ++c) {
            Int lv = d->diag_vertices[d->type_diag_off[t] + c];
            vd.diagonal_vertices[c + 1] = lv ? make_vertex_code(true, lv) : VERTEX_INVALID;
        }
        vd.signs.assign(vd.nv + 1, 1);
        vd.weights.assign(vd.nv + 1, 0.0);
        vd.leg_states.assign(LegCount * vd.nv, 0);
        std::vector<uint8_t> isdiag(vd.nv + 1, 0);
        for (Int v = 0; v < vd.nv; ++v) {
            vd.signs[v + 1] = d->signs[v0 + v];
            vd.weights[v + 1] = d->weights[v0 + v];
            const uint8_t *ls = d->leg_states + 4 * (v0 + v);
            for (Int l = 0; l < 4; ++l) vd.leg_states[LegCount * v + l] = ls[l];
            isdiag[v + 1] = (ls[0] == ls[2] && ls[1] == ls[3]);
        }
        vd.transitions.assign(LegCount * vd.max_worm * vd.nv, Transition());
        vd.transition_cumprobs.assign(1, 0.0);
        vd.transition_targets.assign(1, VERTEX_INVALID);
        vd.transition_step_outs.assign(1, {0, 0});
        for (Int v = 0; v < vd.nv; ++v)
            for (Int w = 0; w < vd.max_worm; ++w)
                for (Int l = 0; l < 4; ++l) {
                    Int idx = ((v0 + v) * d->max_worm + w) * 4 + l;
                    Int off = d->trans_offset[idx];
                    if (off < 0) continue;
                    Int cnt = d->trans_count[idx];
                    Transition tr;
                    tr.offset = Int(vd.transition_cumprobs.size());
                    tr.length = cnt - 1;
                    for (Int o = 0; o < cnt; ++o) {
                        vd.transition_cumprobs.push_back(d->out_cumprob[off + o]);
                        Int tv = d->out_target[off + o];
                        vd.transition_targets.push_back(make_vertex_code(isdiag[tv], tv));
                        vd.transition_step_outs.push_back({d->out_leg[off + o] + 1, d->out_worm[off + o]});
                    }
                    vd.transitions[l + LegCount * (w + vd.max_worm * v)] = tr;
                }
    }
    s->energy_offset = d->energy_offset;
    s->norm_site_count = d->norm_site_count;
    s->n_estimators = d->n_estimators;
    s->est_max_dim = d->est_max_dim;
    if (d->n_estimators > 0)
        s->est_values.assign(d->est_values, d->est_values + size_t(d->n_estimators) * d->n_sites * d->est_max_dim);
    return s;
}

// --- random streams (include/sse_rng.h contract) -----------------------------------------------------
struct Rng {
    int kind = 0;  // 0 = Philox draw-index stream, 1 = injected array, 2 = xoshiro256++ (timing baseline only)
    uint64_t seed = 0, walker = 0, pos = 0;
    const uint64_t *stream = nullptr;
    uint64_t stream_len = 0;
    bool exhausted = false;
    uint64_t s[4] = {1, 2, 3, 4};
    uint64_t cached_j = ~0ull;
    uint32_t cached[4] = {0, 0, 0, 0};
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    void seed_xoshiro(uint64_t sd) {  // splitmix64 expansion
        for (int i = 0; i < 4; ++i) {
            uint64_t z = (sd += 0x9e3779b97f4a7c15ULL);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    inline uint64_t next() {
        if (kind == 2) {
            uint64_t result = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
            s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
            ++pos;
            return result;
        }
        uint64_t k = pos++;
        if (kind == 1) {
            if (k >= stream_len) { exhausted = true; return 0; }
            return stream[k];
        }
        if ((k >> 1) != cached_j) {  // one Philox block = two draws (include/sse_rng.h)
            cached_j = k >> 1;
            sse_philox_block(seed, walker, cached_j, cached);
        }
        return (k & 1) ? (uint64_t(cached[2]) | (uint64_t(cached[3]) << 32)) : (uint64_t(cached[0]) | (uint64_t(cached[1]) << 32));
    }
    inline double U() { return sse_u01(next()); }                       // rand(rng)
    inline Int I(Int k) { return 1 + Int(sse_uint_below(next(), uint64_t(k))); }  // rand(rng, 1:k)
};

// src/vertex_list.jl:1-13
struct VertexList {
    std::vector<std::pair<Int, Int>> vertices;  // [leg, p] column-major 1-based: (leg-1) + 4*(p-1)
    Int ncols = 0;
    std::vector<std::pair<Int, Int>> v_first, v_last;  // [1..site_count]
    std::pair<Int, Int> &at(Int leg, Int p) { return vertices[(leg - 1) + LegCount * (p - 1)]; }
};

// src/vertex_list.jl:15-54
void make_vertex_list(VertexList &vl, const std::vector<OperCode> &operators /*1-based*/, const std::vector<SSEBond> &bonds) {
    const Int M = Int(operators.size()) - 1;
    if (vl.ncols != M) { vl.vertices.assign(size_t(LegCount) * M, {-1, -1}); vl.ncols = M; }
    std::fill(vl.vertices.begin(), vl.vertices.end(), std::make_pair<Int, Int>(-1, -1));
    std::fill(vl.v_first.begin(), vl.v_first.end(), std::make_pair<Int, Int>(-1, -1));
    std::fill(vl.v_last.begin(), vl.v_last.end(), std::make_pair<Int, Int>(-1, -1));
    for (Int p = 1; p <= M; ++p) {
        OperCode op = operators[p];
        if (isidentity(op)) continue;
        const SSEBond &b = bonds[get_bond(op)];
        for (Int s = 1; s <= NSites; ++s) {
            auto [s1, p1] = vl.v_last[b.sites[s - 1]];
            if (p1 != -1) {
                vl.at(s1, p1) = {s, p};
                vl.at(s, p) = {s1, p1};
            } else {
                vl.v_first[b.sites[s - 1]] = {s, p};
            }
            vl.v_last[b.sites[s - 1]] = {NSites + s, p};
        }
    }
    for (size_t i = 1; i < vl.v_first.size(); ++i) {
        if (vl.v_first[i].first != -1) {
            vl.at(vl.v_first[i].first, vl.v_first[i].second) = vl.v_last[i];
            vl.at(vl.v_last[i].first, vl.v_last[i].second) = vl.v_first[i];
        }
    }
}

// src/sse.jl:6-24
struct MC {
    double T = 1;
    double target_worm_length_fraction = 2.0, num_worms_attenuation_factor = 0.01;
    double avg_worm_length = 1.0, num_worms = 5.0;
    Int num_operators = 0;
    std::vector<OperCode> operators;   // 1-based (element 0 unused)
    std::vector<StateIndex> state;     // 1-based
    const SSEData *sse_data = nullptr;
    VertexList vertex_list;
    Rng rng;
    // bookkeeping outside the reference struct
    std::vector<double> acc;           // accumulated observables
    int64_t acc_count[2] = {0, 0};
    double last_wlf = NAN;
    uint64_t counters[4] = {0, 0, 0, 0};  // visits, sweeps, sum n, sum M
    uint32_t flags = 0;
    uint64_t visit_cap = 0;            // screening aid only (tests/golden/screen_seeds.py): abandon the walker beyond this many visits
    bool aborted = false;
    Int M() const { return Int(operators.size()) - 1; }
    Int n_obs() const { return SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * sse_data->n_estimators; }
};

// src/sse.jl:137-191
void diagonal_update(MC &mc) {
    const SSEData &sd = *mc.sse_data;
    if (double(mc.num_operators) >= double(mc.M()) * 0.5) {
        Int old_length = mc.M();
        Int new_length = Int(std::floor(double(old_length) * 1.5 + 100));
        mc.operators.resize(new_length + 1, OperCode(0));
    }
    const Int Mlen = mc.M();
    const Int nb = sd.nbonds();
    const double p_make_bond_raw = double(nb) / mc.T;
    const double p_remove_bond_raw = mc.T / double(nb);

    for (Int iop = 1; iop <= Mlen; ++iop) {
        OperCode op = mc.operators[iop];
        if (isidentity(op)) {
            Int bond = mc.rng.I(nb);
            const SSEBond &b = sd.bonds[bond];
            // join_idx(dims, idxs) (src/util.jl:15-23)
            Int r = 0;
            for (Int k = NSites; k >= 1; --k) {
                r *= sd.sites[b.sites[k - 1]].dim;
                r += Int(mc.state[b.sites[k - 1]]) - 1;
            }
            Int state_idx = r + 1;
            const VertexData &vd = sd.get_vertex_data(bond);
            VertexCode new_vert = vd.get_diagonal_vertex(state_idx);
            double weight = vd.get_vertex_weight(new_vert);
            double p_make_bond = p_make_bond_raw / double(Mlen - mc.num_operators);
            if (mc.rng.U() < p_make_bond * weight) {
                mc.operators[iop] = make_opercode(bond, new_vert);
                mc.num_operators += 1;
            }
        } else {
            Int bond = get_bond(op);
            const VertexData &vd = sd.get_vertex_data(bond);
            if (op_isdiagonal(op)) {
                double weight = vd.get_vertex_weight(get_vertex(op));
                double p_remove_bond = double(Mlen - mc.num_operators + 1) * p_remove_bond_raw;
                if (mc.rng.U() * weight < p_remove_bond) {
                    mc.operators[iop] = OperCode(0);
                    mc.num_operators -= 1;
                }
            } else {
                const SSEBond &b = sd.bonds[bond];
                const StateIndex *leg_state = vd.get_leg_state(get_vertex(op));
                for (Int s = 1; s <= NSites; ++s) mc.state[b.sites[s - 1]] = leg_state[NSites + s];
            }
        }
    }
}

// src/sse.jl:262-303
Int worm_traverse(MC &mc, Int l0, Int p0, Int wormfunc0) {
    const SSEData &sd = *mc.sse_data;
    Int leg_in = l0, p = p0, wormfunc = wormfunc0;
    Int worm_length = 1;
    while (true) {
        OperCode op = mc.operators[p];
        Int bond = get_bond(op);
        bool ft = false;
        ScatterResult r = scatter(sd.get_vertex_data(bond), get_vertex(op), leg_in, wormfunc, mc.rng.U(), &ft);
        if (ft) mc.flags |= SSE_FLAG_SCATTER_FALLTHROUGH;
        mc.operators[p] = make_opercode(bond, r.target);
        const SSESite &site_out = sd.sites[sd.bonds[bond].sites[site_of_leg(r.leg_out, NSites) - 1]];
        if (p == p0 && r.leg_out == l0 && r.worm_out == worm_inverse(wormfunc0, site_out.dim)) break;
        worm_length += 1;
        wormfunc = r.worm_out;
        auto lp = mc.vertex_list.at(r.leg_out, p);
        leg_in = lp.first;
        p = lp.second;
        if (p == p0 && leg_in == l0 && wormfunc == wormfunc0) break;
        if (mc.visit_cap && mc.counters[0] + uint64_t(worm_length) > mc.visit_cap) { mc.aborted = true; break; }
    }
    return worm_length;
}

// src/sse.jl:233-260
Int worm_traverse(MC &mc) {
    if (mc.num_operators == 0) return 0;
    const SSEData &sd = *mc.sse_data;
    Int p0 = 0, l0 = 0;
    while (true) {
        p0 = mc.rng.I(mc.M());
        l0 = mc.rng.I(LegCount);
        if (mc.vertex_list.at(l0, p0).first > 0) break;
        if (mc.rng.exhausted) return 0;
    }
    OperCode op0 = mc.operators[p0];
    Int site0 = sd.bonds[get_bond(op0)].sites[site_of_leg(l0, LegCount / 2) - 1];
    Int wormfunc0 = mc.rng.I(worm_count(sd.sites[site0].dim));
    return worm_traverse(mc, l0, p0, wormfunc0);
}

// src/sse.jl:193-231
void worm_update(MC &mc, bool thermalized) {
    const SSEData &sd = *mc.sse_data;
    double total_worm_length = 1.0;
    const Int nworms = Int(std::ceil(mc.num_worms));
    for (Int i = 0; i < nworms; ++i) {
        Int worm_length = worm_traverse(mc);
        total_worm_length += double(worm_length);
        mc.counters[0] += uint64_t(worm_length);
    }
    if (thermalized && mc.num_operators != 0) {
        mc.last_wlf = total_worm_length / double(mc.num_operators);  // measure!(ctx, :WormLengthFraction, ...)
        mc.acc[SSE_OBS_WORM_LENGTH_FRACTION] += mc.last_wlf;
        mc.acc_count[1] += 1;
    }
    double avg_worm_length = total_worm_length / std::ceil(mc.num_worms);
    if (!thermalized) {
        mc.avg_worm_length += mc.num_worms_attenuation_factor * (avg_worm_length - mc.avg_worm_length);
        double target_worms = mc.target_worm_length_fraction * double(mc.num_operators) / mc.avg_worm_length;
        mc.num_worms += mc.num_worms_attenuation_factor *
                        (target_worms - mc.num_worms + 100.0 * sse_tanh(target_worms - mc.num_worms));
        if (mc.num_worms_attenuation_factor != 0) {
            double lo = 1.0, hi = 1.0 + double(mc.num_operators) / 2.0;
            mc.num_worms = mc.num_worms < lo ? lo : (mc.num_worms > hi ? hi : mc.num_worms);  // clamp
        }
    }
    for (Int i = 1; i <= sd.nsites(); ++i) {
        auto [l, p] = mc.vertex_list.v_first[i];
        if (p < 0) {
            mc.state[i] = StateIndex(mc.rng.I(sd.sites[i].dim));
        } else {
            OperCode op = mc.operators[p];
            mc.state[i] = sd.get_vertex_data(get_bond(op)).get_leg_state(get_vertex(op))[l];
        }
    }
}

// src/sse.jl:305-314
double measure_sign(const MC &mc) {
    Int sign = 0;
    for (Int p = 1; p <= mc.M(); ++p) {
        OperCode op = mc.operators[p];
        if (!isidentity(op)) sign += mc.sse_data->get_vertex_data(get_bond(op)).get_sign(get_vertex(op)) < 0;
    }
    return (sign & 1) ? -1.0 : 1.0;
}

// src/models/common/magnetization_estimator.jl:33-46 (fields) with the table-driven m(site, state)
struct MagEst { double n, tmpmag, mag, absmag, mag2, mag4; };

// src/sse.jl:70-87 + measure_opstring! (:321-376) + MagnetizationEstimator init/measure/result
// (src/models/common/magnetization_estimator.jl:96-230).  out[n_obs].
void measure(MC &mc, double *out) {
    const SSEData &sd = *mc.sse_data;
    const double sign = measure_sign(mc);
    const double nops = double(mc.num_operators);
    out[SSE_OBS_SIGN] = sign;
    out[SSE_OBS_OPERATOR_COUNT] = nops;
    out[SSE_OBS_SIGN_OPERATOR_COUNT] = sign * nops;
    out[SSE_OBS_SIGN_OPERATOR_COUNT2] = sign * (nops * nops);
    out[SSE_OBS_SIGN_ENERGY] = -sign * (nops * mc.T + sd.energy_offset) / double(sd.norm_site_count);
    out[SSE_OBS_WORM_LENGTH_FRACTION] = mc.last_wlf;
    const Int ne = sd.n_estimators;
    if (ne == 0) return;
    std::vector<MagEst> est(ne);
    for (Int e = 0; e < ne; ++e) {  // init (:96-123)
        double tmpmag = 0;
        for (Int site = 1; site <= sd.nsites(); ++site) tmpmag += sd.est(e, site, mc.state[site]);
        est[e] = {1.0, tmpmag, tmpmag, std::fabs(tmpmag), tmpmag * tmpmag, (tmpmag * tmpmag) * (tmpmag * tmpmag)};
    }
    Int n = 0;
    for (Int p = 1; p <= mc.M(); ++p) {
        OperCode op = mc.operators[p];
        if (isidentity(op)) continue;
        const SSEBond &b = sd.bonds[get_bond(op)];
        const VertexData &vd = sd.get_vertex_data(get_bond(op));
        const StateIndex *leg_state = vd.get_leg_state(get_vertex(op));
        if (!op_isdiagonal(op)) {
            for (Int i = 1; i <= NSites; ++i) mc.state[b.sites[i - 1]] = leg_state[NSites + i];
        }
        if (n < mc.num_operators) {
            for (Int e = 0; e < ne; ++e) {  // measure (:125-163)
                MagEst &m = est[e];
                if (!op_isdiagonal(op)) {
                    for (Int l = 1; l <= NSites; ++l) {
                        Int site = b.sites[l - 1];
                        m.tmpmag += sd.est(e, site, leg_state[NSites + l]) - sd.est(e, site, leg_state[l]);
                    }
                }
                m.mag += m.tmpmag;
                m.absmag += std::fabs(m.tmpmag);
                double t2 = m.tmpmag * m.tmpmag;
                m.mag2 += t2;
                m.mag4 += t2 * t2;
                m.n += 1;
            }
        }
        n += 1;
    }
    for (Int e = 0; e < ne; ++e) {  // result (:205-230)
        MagEst &m = est[e];
        double norm = 1.0 / double(sd.norm_site_count);
        m.mag *= norm;
        m.absmag *= norm;
        m.mag2 *= norm * norm;
        m.mag4 *= (norm * norm) * (norm * norm);
        double *o = out + SSE_OBS_FIXED + SSE_OBS_PER_ESTIMATOR * e;
        o[0] = sign * m.mag / m.n;
        o[1] = sign * m.absmag / m.n;
        o[2] = sign * m.mag2 / m.n;
        o[3] = sign * m.mag4 / m.n;
        double chi = 1.0 / mc.T / (m.n + 1) / m.n * (m.mag * m.mag + m.mag2) * double(sd.norm_site_count);
        o[4] = sign * chi;
    }
}

// src/sse.jl:62-68 (+ :70-87 when measuring)
void sweep(MC &mc, bool thermalized, bool do_measure) {
    diagonal_update(mc);
    make_vertex_list(mc.vertex_list, mc.operators, mc.sse_data->bonds);
    worm_update(mc, thermalized);
    mc.counters[1] += 1;
    mc.counters[2] += uint64_t(mc.num_operators);
    mc.counters[3] += uint64_t(mc.M());
    if (do_measure) {
        std::vector<double> out(mc.n_obs());
        measure(mc, out.data());
        for (Int i = 0; i < mc.n_obs(); ++i)
            if (i != SSE_OBS_WORM_LENGTH_FRACTION) mc.acc[i] += out[i];
        mc.acc_count[0] += 1;
    }
}

// src/sse.jl:47-60
void init(MC &mc, Int init_opstring_cutoff, Int diagonal_warmup_sweeps) {
    const SSEData &sd = *mc.sse_data;
    mc.state.assign(sd.nsites() + 1, 0);
    for (Int i = 1; i <= sd.nsites(); ++i) mc.state[i] = StateIndex(mc.rng.I(sd.sites[i].dim));
    if (init_opstring_cutoff < 0) init_opstring_cutoff = Int(std::nearbyint(double(sd.nsites()) * mc.T));  // round(Int, N*T)
    mc.operators.assign(init_opstring_cutoff + 1, OperCode(0));
    mc.num_operators = 0;
    for (Int i = 0; i < diagonal_warmup_sweeps; ++i) diagonal_update(mc);
}

MC *new_walker(const SSEData *sd, double T, int rng_kind, uint64_t seed, uint64_t walker_id, double twlf, double atten,
               double init_num_worms) {
    MC *mc = new MC();
    mc->sse_data = sd;
    mc->T = T;
    mc->target_worm_length_fraction = twlf;
    mc->num_worms_attenuation_factor = atten;
    mc->num_worms = init_num_worms;
    mc->rng.kind = rng_kind;
    mc->rng.seed = seed;
    mc->rng.walker = walker_id;
    if (rng_kind == 2) mc->rng.seed_xoshiro(seed * 0x9E3779B97F4A7C15ULL + walker_id);
    mc->operators.assign(1, 0);
    mc->state.assign(sd->nsites() + 1, 1);
    mc->vertex_list.v_first.assign(sd->nsites() + 1, {-1, -1});
    mc->vertex_list.v_last.assign(sd->nsites() + 1, {-1, -1});
    mc->acc.assign(mc->n_obs(), 0.0);
    return mc;
}

}  // namespace

extern "C" {

void *oracle_model_create(const sse_model_desc *d) { return build_sse_data(d); }
void oracle_model_destroy(void *m) { delete static_cast<SSEData *>(m); }

void *oracle_walker_create(void *model, double T, int32_t rng_kind, uint64_t seed, uint64_t walker_id, double twlf,
                           double atten, double init_num_worms) {
    return new_walker(static_cast<SSEData *>(model), T, rng_kind, seed, walker_id, twlf, atten, init_num_worms);
}
void oracle_walker_destroy(void *w) { delete static_cast<MC *>(w); }

// injected stream: the caller keeps `stream` alive; position restarts at 0
void oracle_set_injected_stream(void *w, const uint64_t *stream, int64_t len) {
    MC *mc = static_cast<MC *>(w);
    if (stream) { mc->rng.kind = 1; mc->rng.stream = stream; mc->rng.stream_len = uint64_t(len); mc->rng.pos = 0; mc->rng.exhausted = false; }
    else { mc->rng.kind = 0; }
}
// screening aid: returns 1 if the walker exceeded `cap` total visits (its configuration is then garbage)
int32_t oracle_sweep_capped(void *w, int32_t n_sweeps, int32_t thermalized, uint64_t cap) {
    MC &mc = *static_cast<MC *>(w);
    mc.visit_cap = cap;
    for (int i = 0; i < n_sweeps && !mc.aborted; ++i) sweep(mc, thermalized != 0, false);
    return mc.aborted;
}
void oracle_set_temperature(void *w, double T) { static_cast<MC *>(w)->T = T; }  // sse.jl:403
uint64_t oracle_rng_draws(void *w) { return static_cast<MC *>(w)->rng.pos; }
void oracle_set_rng_draws(void *w, uint64_t pos) { static_cast<MC *>(w)->rng.pos = pos; }
int32_t oracle_stream_exhausted(void *w) { return static_cast<MC *>(w)->rng.exhausted; }
uint32_t oracle_flags(void *w) { return static_cast<MC *>(w)->flags; }

void oracle_init(void *w, int64_t cutoff, int32_t warmup) { init(*static_cast<MC *>(w), cutoff, warmup); }
void oracle_sweep(void *w, int32_t n_sweeps, int32_t thermalized, int32_t do_measure) {
    MC &mc = *static_cast<MC *>(w);
    for (int i = 0; i < n_sweeps; ++i) sweep(mc, thermalized != 0, do_measure != 0);
}
void oracle_diagonal_update(void *w) { diagonal_update(*static_cast<MC *>(w)); }
void oracle_make_vertex_list(void *w) {
    MC &mc = *static_cast<MC *>(w);
    make_vertex_list(mc.vertex_list, mc.operators, mc.sse_data->bonds);
}
void oracle_worm_update(void *w, int32_t thermalized) { worm_update(*static_cast<MC *>(w), thermalized != 0); }
int64_t oracle_worm_traverse(void *w, int32_t l0, int64_t p0, int32_t wormfunc0) {
    return worm_traverse(*static_cast<MC *>(w), l0, p0, wormfunc0);
}
void oracle_measure(void *w, double *out) { measure(*static_cast<MC *>(w), out); }
int32_t oracle_n_obs(void *w) { return int32_t(static_cast<MC *>(w)->n_obs()); }
void oracle_fetch_accumulators(void *w, double *sums, int64_t *counts, int32_t reset) {
    MC &mc = *static_cast<MC *>(w);
    std::copy(mc.acc.begin(), mc.acc.end(), sums);
    counts[0] = mc.acc_count[0];
    counts[1] = mc.acc_count[1];
    if (reset) { std::fill(mc.acc.begin(), mc.acc.end(), 0.0); mc.acc_count[0] = mc.acc_count[1] = 0; }
}
void oracle_fetch_counters(void *w, uint64_t out[4], int32_t reset) {
    MC &mc = *static_cast<MC *>(w);
    for (int i = 0; i < 4; ++i) { out[i] = mc.counters[i]; if (reset) mc.counters[i] = 0; }
}

int64_t oracle_opstring_length(void *w) { return static_cast<MC *>(w)->M(); }
// same struct as the product's checkpoint boundary (src/sse.jl:89-107)
void oracle_get_state(void *w, sse_walker_state *st) {
    MC &mc = *static_cast<MC *>(w);
    st->num_operators = mc.num_operators;
    st->avg_worm_length = mc.avg_worm_length;
    st->num_worms = mc.num_worms;
    Int M = mc.M();
    if (st->operators && st->operators_len >= M) std::copy(mc.operators.begin() + 1, mc.operators.end(), st->operators);
    st->operators_len = M;
    if (st->state) std::copy(mc.state.begin() + 1, mc.state.end(), st->state);
    st->rng_draws = mc.rng.pos;
    st->T = mc.T;
}
void oracle_set_state(void *w, const sse_walker_state *st) {
    MC &mc = *static_cast<MC *>(w);
    mc.num_operators = st->num_operators;
    mc.avg_worm_length = st->avg_worm_length;
    mc.num_worms = st->num_worms;
    mc.operators.assign(st->operators_len + 1, 0);
    std::copy(st->operators, st->operators + st->operators_len, mc.operators.begin() + 1);
    std::copy(st->state, st->state + mc.sse_data->nsites(), mc.state.begin() + 1);
    mc.rng.pos = st->rng_draws;
    mc.T = st->T;
}
// vertices[M][4][2] (leg, p), v_first/v_last[n_sites][2] — the reference's VertexList contents
void oracle_get_vertex_list(void *w, int64_t *vertices, int64_t *v_first, int64_t *v_last) {
    MC &mc = *static_cast<MC *>(w);
    VertexList &vl = mc.vertex_list;
    for (Int p = 1; p <= vl.ncols; ++p)
        for (Int l = 1; l <= LegCount; ++l) {
            vertices[((p - 1) * LegCount + (l - 1)) * 2 + 0] = vl.at(l, p).first;
            vertices[((p - 1) * LegCount + (l - 1)) * 2 + 1] = vl.at(l, p).second;
        }
    for (Int i = 1; i <= mc.sse_data->nsites(); ++i) {
        v_first[2 * (i - 1)] = vl.v_first[i].first; v_first[2 * (i - 1) + 1] = vl.v_first[i].second;
        v_last[2 * (i - 1)] = vl.v_last[i].first;  v_last[2 * (i - 1) + 1] = vl.v_last[i].second;
    }
}

// Timed CPU baseline: `n_threads` independent walkers (one per thread, as Carlo runs one MC per MPI rank,
// docs/src/tutorial.md:49), xoshiro256++ stream, `therm` un-thermalised + `sweeps` timed sweeps each.
// out = {seconds (max over threads), total visits, total walker-sweeps, mean n, mean M, sum of per-thread seconds}.
// `doublings` > 0: the walkers are brought to T the way bench.py brings the device walkers there (sse_double_beta restated:
// init! at T * 2^doublings, `per_level` sweeps and one doubling (state, S_M) -> (state, S_M S_M), n -> 2n, T -> T/2 per level
// with the controller's attenuation factor at 0.1), then `therm` sweeps at T.  Needed for L = beta = 64, where the
// reference's cold start takes thousands of sweeps.
void oracle_bench2(void *model, double T, int32_t n_threads, int32_t doublings, int32_t per_level, int32_t therm, int32_t sweeps,
                   uint64_t seed, double *out) {
    SSEData *sd = static_cast<SSEData *>(model);
    std::vector<double> secs(n_threads, 0.0), visits(n_threads, 0.0), nsum(n_threads, 0.0), msum(n_threads, 0.0);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    auto worker = [&](int t) {
        MC *mc = new_walker(sd, std::ldexp(T, doublings), 2, seed, uint64_t(t), 2.0, doublings > 0 ? 0.1 : 0.01, 5.0);
        init(*mc, -1, 5);
        for (int level = 0; level < doublings; ++level) {
            for (int i = 0; i < per_level; ++i) sweep(*mc, false, false);
            const Int M = mc->M();
            mc->operators.resize(2 * M + 1);
            std::copy(mc->operators.begin() + 1, mc->operators.begin() + 1 + M, mc->operators.begin() + 1 + M);
            mc->num_operators *= 2;
            mc->T *= 0.5;
            mc->avg_worm_length *= 2.0;
        }
        mc->num_worms_attenuation_factor = 0.01;
        for (int i = 0; i < therm; ++i) sweep(*mc, false, false);
        for (int i = 0; i < 4; ++i) mc->counters[i] = 0;
        ready.fetch_add(1);
        while (!go.load()) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < sweeps; ++i) sweep(*mc, true, false);
        auto t1 = std::chrono::steady_clock::now();
        secs[t] = std::chrono::duration<double>(t1 - t0).count();
        visits[t] = double(mc->counters[0]);
        nsum[t] = double(mc->counters[2]);
        msum[t] = double(mc->counters[3]);
        delete mc;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    while (ready.load() < n_threads) std::this_thread::yield();
    go.store(true);
    for (auto &x : th) x.join();
    double smax = 0, ssum = 0, v = 0, ns = 0, ms = 0;
    for (int t = 0; t < n_threads; ++t) { smax = std::max(smax, secs[t]); ssum += secs[t]; v += visits[t]; ns += nsum[t]; ms += msum[t]; }
    out[0] = smax; out[1] = v; out[2] = double(n_threads) * sweeps;
    out[3] = ns / (double(n_threads) * sweeps); out[4] = ms / (double(n_threads) * sweeps); out[5] = ssum;
}

void oracle_bench(void *model, double T, int32_t n_threads, int32_t therm, int32_t sweeps, uint64_t seed, double *out) {
    SSEData *sd = static_cast<SSEData *>(model);
    std::vector<double> secs(n_threads, 0.0), visits(n_threads, 0.0), nsum(n_threads, 0.0), msum(n_threads, 0.0);
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    auto worker = [&](int t) {
        MC *mc = new_walker(sd, T, 2, seed, uint64_t(t), 2.0, 0.01, 5.0);
        init(*mc, -1, 5);
        for (int i = 0; i < therm; ++i) sweep(*mc, false, false);
        for (int i = 0; i < 4; ++i) mc->counters[i] = 0;
        ready.fetch_add(1);
        while (!go.load()) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < sweeps; ++i) sweep(*mc, true, false);
        auto t1 = std::chrono::steady_clock::now();
        secs[t] = std::chrono::duration<double>(t1 - t0).count();
        visits[t] = double(mc->counters[0]);
        nsum[t] = double(mc->counters[2]);
        msum[t] = double(mc->counters[3]);
        delete mc;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    while (ready.load() < n_threads) std::this_thread::yield();
    go.store(true);
    for (auto &x : th) x.join();
    double smax = 0, ssum = 0, v = 0, ns = 0, ms = 0;
    for (int t = 0; t < n_threads; ++t) { smax = std::max(smax, secs[t]); ssum += secs[t]; v += visits[t]; ns += nsum[t]; ms += msum[t]; }
    out[0] = smax; out[1] = v; out[2] = double(n_threads) * sweeps;
    out[3] = ns / (double(n_threads) * sweeps); out[4] = ms / (double(n_threads) * sweeps); out[5] = ssum;
}

}  // extern "C"
