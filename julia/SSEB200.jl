# SSEB200.jl — the reference-side binding: a `Carlo.AbstractMC` whose sweep runs in libsse_b200.so.
#
# NOT runnable in the build image (julia is absent there); it is the `ccall` layer a maintainer adds
# next to StochasticSeriesExpansion.jl.  It mirrors, one C call per Carlo method, what the Python host
# (stochasticseriesexpansion.jl_b200/{capi,walkers,mc}.py) does and what the GPU tests exercise.
# Every ccall names the export of include/sse_b200.h it binds and the reference method it replaces.
#
# Usage (drop-in for `StochasticSeriesExpansion.MC` in a Carlo job file):
#     job = JobInfo("myjob", SSEB200.MC; tasks = make_tasks(tm), ...)     # tm.n_walkers = 4096 optional
module SSEB200

using Carlo
using HDF5
using LinearAlgebra
import StochasticSeriesExpansion as S

const libsse = get(ENV, "SSE_B200_LIB", "libsse_b200.so")

# struct sse_model_desc (include/sse_b200.h) — field order and types must match exactly
struct ModelDesc
    n_sites::Int32
    site_dim::Ptr{UInt8}
    n_bonds::Int32
    bond_type::Ptr{Int32}
    bond_sites::Ptr{Int32}
    n_types::Int32
    type_dims::Ptr{Int32}
    type_vertex_off::Ptr{Int32}
    type_diag_off::Ptr{Int32}
    n_vertices::Int32
    weights::Ptr{Float64}
    signs::Ptr{Int8}
    leg_states::Ptr{UInt8}
    diag_vertices::Ptr{Int32}
    max_worm::Int32
    trans_offset::Ptr{Int32}
    trans_count::Ptr{Int32}
    n_outcomes::Int32
    out_cumprob::Ptr{Float64}
    out_target::Ptr{Int32}
    out_leg::Ptr{Int32}
    out_worm::Ptr{Int32}
    energy_offset::Float64
    norm_site_count::Int32
    n_estimators::Int32
    est_max_dim::Int32
    est_values::Ptr{Float64}
end

struct WalkersOpts
    n_walkers::Int32
    T::Ptr{Float64}
    m_capacity::Int64
    n_capacity::Int64
    device::Int32
    seed::UInt64
    walker_id_offset::UInt64
    target_worm_length_fraction::Float64
    num_worms_attenuation_factor::Float64
    init_num_worms::Float64
end

mutable struct WalkerState
    num_operators::Int64
    avg_worm_length::Float64
    num_worms::Float64
    operators::Ptr{UInt64}
    operators_len::Int64
    state::Ptr{UInt8}
    rng_draws::UInt64
    T::Float64
end

"the same layout as an isbits struct: a Vector{WalkerStateC} is a C array of sse_walker_state (sse_set_states)"
struct WalkerStateC
    num_operators::Int64
    avg_worm_length::Float64
    num_worms::Float64
    operators::Ptr{UInt64}
    operators_len::Int64
    state::Ptr{UInt8}
    rng_draws::UInt64
    T::Float64
end

function check(status::Int32)
    status == 0 && return nothing
    error("libsse_b200: " * unsafe_string(ccall((:sse_last_error, libsse), Cstring, ())))
end

"Flatten `SSEData{2}` + estimator tables into the POD arrays of `sse_model_desc` (SURVEY.md Appendix B)."
function flatten(model::S.AbstractModel, sse_data::S.SSEData{2}, estimators)
    nsites = length(sse_data.sites)
    vds = sse_data.vertex_data
    max_worm = max(1, maximum(maximum(vd.dims) - 1 for vd in vds))
    voff = Int32[0]; doff = Int32[0]
    for vd in vds
        push!(voff, voff[end] + length(vd.weights))
        push!(doff, doff[end] + length(vd.diagonal_vertices))
    end
    nv = voff[end]
    trans_offset = fill(Int32(-1), nv * max_worm * 4)
    trans_count = zeros(Int32, nv * max_worm * 4)
    out_cumprob = Float64[]; out_target = Int32[]; out_leg = Int32[]; out_worm = Int32[]
    diag = Int32[]
    for (t, vd) in enumerate(vds)
        append!(diag, (S.isinvalid(c) ? Int32(0) : Int32(S.get_vertex_idx(c)) for c in vd.diagonal_vertices))
        base = length(out_cumprob)
        append!(out_cumprob, vd.transition_cumprobs)
        append!(out_target, Int32.(S.get_vertex_idx.(vd.transition_targets)))
        append!(out_leg, Int32.(first.(vd.transition_step_outs) .- 1))
        append!(out_worm, Int32.(last.(vd.transition_step_outs)))
        for v in axes(vd.transitions, 3), w in axes(vd.transitions, 2), l in axes(vd.transitions, 1)
            tr = vd.transitions[l, w, v]
            S.isinvalid(tr) && continue
            idx = ((voff[t] + v - 1) * max_worm + (w - 1)) * 4 + l
            trans_offset[idx] = base + tr.offset - 1
            trans_count[idx] = tr.length + 1
        end
    end
    max_dim = maximum(s.dim for s in sse_data.sites)
    est = zeros(Float64, max_dim, nsites, length(estimators))   # column-major == [e][site][state] in C
    for (e, E) in enumerate(estimators), site in 1:nsites, state in 1:sse_data.sites[site].dim
        (q, stag, _, _, tag) = E.parameters
        est[state, site, e] = S.staggered_sign(model, q, stag, site) * S.magnetization_state(model, Val(tag), site, state)
    end
    return (
        site_dim = UInt8[s.dim for s in sse_data.sites],
        bond_type = Int32[b.type - 1 for b in sse_data.bonds],
        bond_sites = Int32[s - 1 for b in sse_data.bonds for s in b.sites],
        type_dims = Int32[d for vd in vds for d in vd.dims],
        voff = voff, doff = doff,
        weights = reduce(vcat, (vd.weights for vd in vds)),
        signs = reduce(vcat, (vd.signs for vd in vds)),
        leg_states = reduce(vcat, (vec(vd.leg_states) for vd in vds)),
        diag = diag, max_worm = Int32(max_worm),
        trans_offset = trans_offset, trans_count = trans_count,
        out_cumprob = out_cumprob, out_target = out_target, out_leg = out_leg, out_worm = out_worm,
        est = est, max_dim = Int32(max_dim),
    )
end

mutable struct MC{Model<:S.AbstractModel} <: AbstractMC
    model::Model
    sse_data::S.SSEData{2}
    estimators::Vector{Type}
    T::Vector{Float64}
    hmodel::Ptr{Cvoid}
    hwalkers::Ptr{Cvoid}
    obs_names::Vector{Symbol}
    nobs::Int
    sweeps_per_call::Int   # > 1: batched mode, see Carlo.sweep! below
end

"""
Upper bound on <n> at temperature `T_min`: <n> = beta * sum_b <W_b> <= beta * sum_b lambda_max(W_b), W_b the (signed)
vertex-weight matrix of bond b in the compound basis.  Same bound as the Python host (mc.py operator_count_bound).
"""
function operator_count_bound(sse_data::S.SSEData{2}, T_min::Real)
    lam = map(sse_data.vertex_data) do vd
        d0, d1 = vd.dims
        W = zeros(d0 * d1, d0 * d1)
        for v in eachindex(vd.weights)
            ls = Int.(vd.leg_states[:, v]) .- 1
            W[ls[1] + d0 * ls[2] + 1, ls[3] + d0 * ls[4] + 1] = vd.weights[v] * vd.signs[v]
        end
        eigmax(Symmetric(0.5 .* (W .+ W')))
    end
    return sum(lam[b.type] for b in sse_data.bonds) / T_min
end

"""
(m_capacity, n_capacity) the string growth rule M <- 1.5 M + 100 while n >= M/2 (src/sse.jl:138-145) cannot exceed:
M <= 3 n + 100 in the worst case, n <= the spectral bound + 8 sigma.  The device cannot resize a string inside a launch
(the reference just calls resize!), so the capacity is fixed at creation; both can be overridden with
params[:m_capacity] / params[:n_capacity], and string slots are cheap (0.25 B each).
"""
function default_capacity(sse_data::S.SSEData{2}, T_min::Real)
    nb = operator_count_bound(sse_data, T_min)
    nb = nb + 8 * sqrt(nb) + 64
    return (ceil(Int, 3 * nb) + 1124, min(ceil(Int, nb) + 256, (1 << 22) - 1))
end

"`MC(params)` — replaces StochasticSeriesExpansion.MC(params) (src/sse.jl:26-45); binds sse_model_create + sse_walkers_create."
function MC(params::AbstractDict)
    model = params[:model](params)
    sse_data = S.generate_sse_data(model)
    S.leg_count(typeof(model)) == 4 || error("SSEB200 supports 2-site bonds (leg_count == 4) only")
    ests = S.get_opstring_estimators(model)
    f = flatten(model, sse_data, ests)
    nw = get(params, :n_walkers, 1)
    T = params[:T] isa AbstractVector ? Float64.(params[:T]) : fill(Float64(params[:T]), nw)
    hmodel = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve f begin
        desc = ModelDesc(length(f.site_dim), pointer(f.site_dim), length(f.bond_type), pointer(f.bond_type),
            pointer(f.bond_sites), length(f.voff) - 1, pointer(f.type_dims), pointer(f.voff), pointer(f.doff),
            f.voff[end], pointer(f.weights), pointer(f.signs), pointer(f.leg_states), pointer(f.diag), f.max_worm,
            pointer(f.trans_offset), pointer(f.trans_count), length(f.out_cumprob), pointer(f.out_cumprob),
            pointer(f.out_target), pointer(f.out_leg), pointer(f.out_worm), sse_data.energy_offset,
            S.normalization_site_count(model), length(ests), f.max_dim, pointer(f.est))
        check(ccall((:sse_model_create, libsse), Int32, (Ref{ModelDesc}, Ref{Ptr{Cvoid}}), desc, hmodel))
    end
    mdef, ndef = default_capacity(sse_data, minimum(T))
    opts = WalkersOpts(length(T), pointer(T), get(params, :m_capacity, mdef), get(params, :n_capacity, ndef), get(params, :device, -1),
        get(params, :seed, 0), get(params, :walker_id_offset, 0), get(params, :target_worm_length_fraction, 2.0),
        get(params, :num_worms_attenuation_factor, 0.01), get(params, :init_num_worms, 5))
    hw = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve T check(ccall((:sse_walkers_create, libsse), Int32, (Ptr{Cvoid}, Ref{WalkersOpts}, Ref{Ptr{Cvoid}}), hmodel[], opts, hw))
    names = [:Sign, :OperatorCount, :SignOperatorCount, :SignOperatorCount2, :SignEnergy, :WormLengthFraction]
    for E in ests, o in (:mag, :absmag, :mag2, :mag4, :magchi)
        push!(names, S.magnetization_estimator_obs_symbols(S.get_prefix(E))[2][o])
    end
    mc = MC{typeof(model)}(model, sse_data, ests, T, hmodel[], hw[], names, length(names), get(params, :sweeps_per_call, 1))
    finalizer(mc) do m
        ccall((:sse_walkers_destroy, libsse), Int32, (Ptr{Cvoid},), m.hwalkers)
        ccall((:sse_model_destroy, libsse), Int32, (Ptr{Cvoid},), m.hmodel)
    end
    return mc
end

"Carlo.init! (src/sse.jl:47-60) -> sse_init.  The stream is Philox keyed by params[:seed]; ctx.rng is not used."
function Carlo.init!(mc::MC, ctx::MCContext, params::AbstractDict)
    check(ccall((:sse_init, libsse), Int32, (Ptr{Cvoid}, Int64, Int32), mc.hwalkers,
        get(params, :init_opstring_cutoff, -1), get(params, :diagonal_warmup_sweeps, 5)))
end

"""
Carlo.sweep! (src/sse.jl:62-68) -> sse_sweep + sse_sync.

Default (`sweeps_per_call = 1`): one launch of one sweep per Carlo step, exactly the reference's call pattern; every step
then waits for the walker with the most worm work (max ~ 4x the mean, SURVEY H1b) and pays a launch ramp, which costs
about a third of the device's throughput at BASELINE sizes (bench.py, key `carlo_call_pattern`).

Batched mode (`params[:sweeps_per_call] = k > 1`): one Carlo step = k sweeps inside one persistent launch, measured on the
device after every sweep once thermalised; `measure!` then pushes the MEAN of those k measurements per walker (a bin of
k samples — Carlo's binning analysis is unchanged because it only ever sees bin means).  Set Carlo's `sweeps`,
`thermalization` and `binsize` in units of calls.
"""
function Carlo.sweep!(mc::MC, ctx::MCContext)
    measure = mc.sweeps_per_call > 1 && is_thermalized(ctx)
    check(ccall((:sse_sweep, libsse), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), mc.hwalkers, mc.sweeps_per_call, is_thermalized(ctx), measure))
    check(ccall((:sse_sync, libsse), Int32, (Ptr{Cvoid},), mc.hwalkers))
end

"Carlo.measure! (src/sse.jl:70-87) -> sse_measure (or, in batched mode, sse_fetch_accumulators); one vector observable (over walkers) per name"
function Carlo.measure!(mc::MC, ctx::MCContext)
    nw = length(mc.T)
    out = Matrix{Float64}(undef, mc.nobs, nw)   # column-major == out[walker][obs] in C
    if mc.sweeps_per_call > 1
        counts = Matrix{Int64}(undef, 2, nw)
        check(ccall((:sse_fetch_accumulators, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int32), mc.hwalkers, out, counts, 1))
        for (i, name) in enumerate(mc.obs_names)
            c = name == :WormLengthFraction ? (@view counts[2, :]) : (@view counts[1, :])
            all(>(0), c) || continue
            measure!(ctx, name, out[i, :] ./ c)
        end
    else
        check(ccall((:sse_measure, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.hwalkers, out))
        for (i, name) in enumerate(mc.obs_names)
            name == :WormLengthFraction && any(isnan, @view out[i, :]) && continue
            measure!(ctx, name, out[i, :])
        end
    end
end

"Carlo.write_checkpoint (src/sse.jl:89-97) -> sse_get_state per walker; same five fields, reference OperCode layout"
function Carlo.write_checkpoint(mc::MC, out::HDF5.Group)
    nsites = length(mc.sse_data.sites)
    for w in eachindex(mc.T)
        st = WalkerState(0, 0, 0, C_NULL, 0, C_NULL, 0, 0)
        # size query: with a null buffer the call fills operators_len (= M) and the scalars
        check(ccall((:sse_get_state, libsse), Int32, (Ptr{Cvoid}, Int32, Ref{WalkerState}), mc.hwalkers, w - 1, st))
        ops = Vector{UInt64}(undef, st.operators_len); state = Vector{UInt8}(undef, nsites)
        GC.@preserve ops state begin
            st.operators = pointer(ops); st.state = pointer(state)
            check(ccall((:sse_get_state, libsse), Int32, (Ptr{Cvoid}, Int32, Ref{WalkerState}), mc.hwalkers, w - 1, st))
        end
        # one walker: the reference's own layout (top-level datasets, src/sse.jl:89-97), so either side reads the other's file
        g = length(mc.T) == 1 ? out : create_group(out, "walker$(w)")
        g["num_operators"] = st.num_operators; g["avg_worm_length"] = st.avg_worm_length
        g["num_worms"] = st.num_worms; g["operators"] = ops; g["state"] = state
        g["rng_draws"] = st.rng_draws; g["T"] = st.T
    end
end

"Carlo.read_checkpoint (src/sse.jl:99-107) -> one sse_set_states call for the whole batch (validated as a whole before
anything is copied).  A checkpoint written by the reference itself (one walker, top-level datasets, `operators` stored as
OperCode structs) is accepted: the stream position defaults to 0 and the temperature to the task's."
function Carlo.read_checkpoint(mc::MC, in::HDF5.Group)
    W = length(mc.T)
    opsv = Vector{Vector{UInt64}}(undef, W)
    statev = Vector{Vector{UInt8}}(undef, W)
    sts = Vector{WalkerStateC}(undef, W)
    for w in 1:W
        g = (W == 1 && !haskey(in, "walker1")) ? in : in["walker$(w)"]
        raw = read(g, "operators")
        opsv[w] = eltype(raw) === UInt64 ? raw : collect(reinterpret(UInt64, raw))   # OperCode is a UInt64 wrapper (opercode.jl:33-35)
        statev[w] = UInt8.(read(g, "state"))
        draws = haskey(g, "rng_draws") ? read(g, "rng_draws") : UInt64(0)
        T = haskey(g, "T") ? read(g, "T") : mc.T[w]
        sts[w] = WalkerStateC(read(g, "num_operators"), read(g, "avg_worm_length"), read(g, "num_worms"), pointer(opsv[w]),
            length(opsv[w]), pointer(statev[w]), draws, T)
    end
    GC.@preserve opsv statev begin
        check(ccall((:sse_set_states, libsse), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{WalkerStateC}), mc.hwalkers, 0, W, sts))
    end
end

"Carlo.register_evaluables (src/sse.jl:111-134): post-processing is unchanged, delegate to the reference"
Carlo.register_evaluables(::Type{<:MC}, eval::AbstractEvaluator, params::AbstractDict) =
    Carlo.register_evaluables(S.MC, eval, params)

"Carlo.parallel_tempering_log_weight_ratio (src/sse.jl:390-396) -> sse_pt_log_weight_ratio"
function Carlo.parallel_tempering_log_weight_ratio(mc::MC, parameter::Symbol, new_value)
    parameter != :T && error("unsupported parallel tempering parameter $parameter")
    newT = new_value isa AbstractVector ? Float64.(new_value) : fill(Float64(new_value), length(mc.T))
    out = similar(newT)
    check(ccall((:sse_pt_log_weight_ratio, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), mc.hwalkers, newT, out))
    return length(out) == 1 ? out[1] : out
end

"Carlo.parallel_tempering_change_parameter! (src/sse.jl:398-405) -> sse_set_temperature"
function Carlo.parallel_tempering_change_parameter!(mc::MC, parameter::Symbol, new_value)
    parameter != :T && error("unsupported parallel tempering parameter $parameter")
    mc.T .= new_value
    check(ccall((:sse_set_temperature, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.hwalkers, mc.T))
    return nothing
end

"Thermalisation aid without a reference counterpart (beta doubling) -> sse_double_beta: every walker's string S_M
becomes S_M S_M at T/2.  Call between unthermalised sweeps; see include/sse_b200.h."
function double_beta!(mc::MC)
    check(ccall((:sse_double_beta, libsse), Int32, (Ptr{Cvoid},), mc.hwalkers))
    mc.T ./= 2
    return nothing
end

"The controller parameters of MC(params) (src/sse.jl:34-35), changeable at run time -> sse_set_controller."
set_controller!(mc::MC, target_worm_length_fraction::Real, num_worms_attenuation_factor::Real) =
    check(ccall((:sse_set_controller, libsse), Int32, (Ptr{Cvoid}, Float64, Float64), mc.hwalkers,
                target_worm_length_fraction, num_worms_attenuation_factor))

"Launch shape of sweep! without a reference counterpart -> sse_set_launch_shape (worm warps: one lane = one walker; stream
warps: one warp = one walker; 0 = automatic)."
set_launch_shape!(mc::MC, worm_warps::Integer, stream_warps::Integer) =
    check(ccall((:sse_set_launch_shape, libsse), Int32, (Ptr{Cvoid}, Int32, Int32), mc.hwalkers, worm_warps, stream_warps))

"Free-running sweeps without a reference counterpart -> sse_advance: every walker does `visit_budget` worm visits (at most
`max_sweeps` sweeps) and is parked wherever it is; `finish_sweeps!` completes the sweeps in flight (needed before
measure!/write_checkpoint).  Use for long thermalisations: no walker waits for another one's long worm."
function advance!(mc::MC, visit_budget::Integer; max_sweeps::Integer = typemax(Int32), thermalized::Bool = false, measure::Bool = false)
    check(ccall((:sse_advance, libsse), Int32, (Ptr{Cvoid}, Int32, UInt64, Int32, Int32), mc.hwalkers, max_sweeps, visit_budget, thermalized, measure))
    check(ccall((:sse_sync, libsse), Int32, (Ptr{Cvoid},), mc.hwalkers))
end
function finish_sweeps!(mc::MC; thermalized::Bool = false, measure::Bool = false)
    check(ccall((:sse_finish_sweeps, libsse), Int32, (Ptr{Cvoid}, Int32, Int32), mc.hwalkers, thermalized, measure))
    check(ccall((:sse_sync, libsse), Int32, (Ptr{Cvoid},), mc.hwalkers))
end

"One bin summed over the walkers of each group on the device and over the ranks of `comm_init!` by NCCL inside the library
-> sse_reduce_bins.  `group[w]` in 0:n_groups-1 (e.g. the temperature index).  Returns (sums[nobs, n_groups], counts[2, n_groups])."
function reduce_bins!(mc::MC, group::Vector{Int32}, n_groups::Integer; reset::Bool = true)
    sums = Matrix{Float64}(undef, mc.nobs, n_groups); counts = Matrix{Int64}(undef, 2, n_groups)
    check(ccall((:sse_reduce_bins, libsse), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Int64}, Int32),
                mc.hwalkers, group, n_groups, sums, counts, reset))
    return sums, counts
end

"NCCL communicator for reduce_bins! -> sse_comm_unique_id (rank 0) + sse_comm_init.  Broadcast the id with MPI, e.g.
`id = MPI.bcast(rank == 0 ? comm_unique_id() : nothing, comm)`."
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:sse_comm_unique_id, libsse), Int32, (Ptr{UInt8},), id))
    return id
end
comm_init!(mc::MC, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ccall((:sse_comm_init, libsse), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), mc.hwalkers, id, rank, nranks))

"Replica exchange decided on the device -> sse_pt_set_ladder / sse_pt_exchange (src/sse.jl:390-405 evaluated by a kernel).
`walker_at_rank`: 0-based walker indices in order of temperature.  Returns the number of accepted pairs; mc.T is refreshed."
pt_set_ladder!(mc::MC, walker_at_rank::Vector{Int32}) =
    check(ccall((:sse_pt_set_ladder, libsse), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32), mc.hwalkers, walker_at_rank, length(walker_at_rank)))
function pt_exchange!(mc::MC, parity::Integer, seed::Integer, step::Integer)
    acc = Ref{Int32}(0)
    check(ccall((:sse_pt_exchange, libsse), Int32, (Ptr{Cvoid}, Int32, UInt64, UInt64, Ref{Int32}), mc.hwalkers, parity, seed, step, acc))
    check(ccall((:sse_get_temperatures, libsse), Int32, (Ptr{Cvoid}, Ptr{Float64}), mc.hwalkers, mc.T))
    return acc[]
end

end # module
