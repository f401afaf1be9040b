# dump_golden.jl — golden-vector dumper for the REAL reference (SURVEY.md §8f-3).
#
# NOT runnable in the build image (no julia there).  Run it where StochasticSeriesExpansion.jl and Carlo.jl are
# installed:   julia --project julia/dump_golden.jl tests/golden/reference_dump
# It drives the reference's own `diagonal_update`, `make_vertex_list!`, `worm_update` and `Carlo.measure!` with an
# INJECTED random stream implementing the draw-index rule of include/sse_rng.h
#     U()  = (x >> 11) * 2^-53          I(k) = 1 + mulhi64(x, k)
# and writes, per case, (tables, operators, state, draws) -> (post-phase operators/state, links, observables) as JSON.
# tests/test_reference_dumps.py replays every dump through the CPU oracle and the CUDA path and demands bit-exact
# agreement, which upgrades level-1 parity from "vs restated oracle" to "vs reference binary".
#
# [ext] Carlo internals used here (MCContext constructor, field `rng`, `sweeps`, `thermalization_sweeps`) follow
# Carlo 0.2.x as far as the reference's own call sites show (src/sse.jl:47-48,139; test/test_sse.jl:79); adjust if the
# installed Carlo differs.
using Random, JSON, Carlo
import StochasticSeriesExpansion as S

mutable struct InjectedRNG <: Random.AbstractRNG
    draws::Vector{UInt64}
    pos::Int
end
InjectedRNG() = InjectedRNG(UInt64[], 0)
InjectedRNG(::Integer) = InjectedRNG()
Random.seed!(r::InjectedRNG, args...) = r

@inline function next!(r::InjectedRNG)
    r.pos += 1
    r.pos > length(r.draws) && error("injected stream exhausted at draw $(r.pos)")
    return r.draws[r.pos]
end
mulhi64(a::UInt64, b::UInt64) = UInt64((UInt128(a) * UInt128(b)) >> 64)

# rand(rng)                      -> U()
Random.rand(r::InjectedRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = Float64(next!(r) >> 11) * 2.0^-53
# rand(rng, 1:k)                 -> I(k)   (call sites src/sse.jl:152,222,242,243,251)
Random.rand(r::InjectedRNG, sp::Random.SamplerTrivial{<:AbstractUnitRange{<:Integer}}) =
    first(sp[]) + oftype(first(sp[]), mulhi64(next!(r), UInt64(length(sp[]))))
Random.Sampler(::Type{InjectedRNG}, x::AbstractUnitRange{<:Integer}, ::Random.Repetition) = Random.SamplerTrivial(x)
# rand(rng, StateIndex.(1:dim))  -> I(dim) (src/sse.jl:48)
Random.rand(r::InjectedRNG, sp::Random.SamplerTrivial{<:AbstractVector}) =
    sp[][1+Int(mulhi64(next!(r), UInt64(length(sp[]))))]
Random.Sampler(::Type{InjectedRNG}, x::AbstractVector, ::Random.Repetition) = Random.SamplerTrivial(x)

"Flatten the tables exactly like julia/SSEB200.jl so the Python side can rebuild the same sse_model_desc."
function tables_json(model, sse_data)
    include(joinpath(@__DIR__, "SSEB200.jl"))
    f = Main.SSEB200.flatten(model, sse_data, S.get_opstring_estimators(model))
    d = Dict{String,Any}(string(k) => (v isa AbstractArray ? vec(collect(v)) : v) for (k, v) in pairs(f))
    d["energy_offset"] = sse_data.energy_offset
    d["norm_site_count"] = S.normalization_site_count(model)
    d["n_est"] = length(S.get_opstring_estimators(model))
    return d
end

function dump_case(name, params, outdir; sweeps = 6, ndraws = 400_000, seed = 1)
    mc = S.MC(params)
    ctx = MCContext{InjectedRNG}(params)
    stream = rand(Random.Xoshiro(seed), UInt64, ndraws)
    ctx.rng.draws = stream
    ctx.rng.pos = 0
    Carlo.init!(mc, ctx, params)
    steps = Any[]
    for s = 1:sweeps
        before = Dict("operators" => [op.code for op in mc.operators], "state" => Int.(mc.state),
            "num_operators" => mc.num_operators, "num_worms" => mc.num_worms,
            "avg_worm_length" => mc.avg_worm_length, "draws_before" => ctx.rng.pos)
        S.diagonal_update(mc, ctx)
        after_diag = Dict("operators" => [op.code for op in mc.operators], "state" => Int.(mc.state),
            "num_operators" => mc.num_operators, "draws" => ctx.rng.pos)
        S.make_vertex_list!(mc.vertex_list, mc.operators, mc.sse_data.bonds)
        vl = Dict("vertices" => [collect(t) for t in vec(mc.vertex_list.vertices)],
            "v_first" => [collect(t) for t in mc.vertex_list.v_first],
            "v_last" => [collect(t) for t in mc.vertex_list.v_last])
        S.worm_update(mc, ctx)
        after_worm = Dict("operators" => [op.code for op in mc.operators], "state" => Int.(mc.state),
            "num_worms" => mc.num_worms, "avg_worm_length" => mc.avg_worm_length, "draws" => ctx.rng.pos)
        push!(steps, Dict("before" => before, "after_diagonal_update" => after_diag, "vertex_list" => vl,
            "after_worm_update" => after_worm))
    end
    out = Dict("name" => name, "T" => params[:T], "tables" => tables_json(mc.model, mc.sse_data),
        "stream" => stream[1:ctx.rng.pos], "steps" => steps,
        "note" => "thermalized(ctx) was false throughout (controller active); tanh is Julia's libm tanh, the oracle uses sse_tanh: compare num_worms to 1e-12, everything else exactly")
    open(joinpath(outdir, "$(name).json"), "w") do io
        JSON.print(io, out)
    end
end

function main(outdir)
    mkpath(outdir)
    dump_case("heisenberg_4x4", Dict{Symbol,Any}(:T => 0.3, :model => S.MagnetModel, :measure => [],
            :lattice => (unitcell = S.UnitCells.square, size = (4, 4)), :J => 1.0, :thermalization => 10^9,
            :sweeps => 1, :binsize => 1), outdir)
    dump_case("spin1_dz_3x3", Dict{Symbol,Any}(:T => 0.3, :model => S.MagnetModel, :measure => [:magnetization],
            :lattice => (unitcell = S.UnitCells.honeycomb, size = (3, 3)), :S => 1, :J => 1.0,
            :Dz => 0.04556 / 8.07, :thermalization => 10^9, :sweeps => 1, :binsize => 1), outdir)
    dump_case("dimer_bilayer_3x3", Dict{Symbol,Any}(:T => 0.3, :model => S.ClusterModel, :inner_model => S.MagnetModel,
            :cluster_bases => (S.ClusterBases.dimer,), :measure_quantum_numbers => [(name = Symbol(), quantum_number = 2)],
            :lattice => (unitcell = S.UnitCells.fully_frust_square_bilayer, size = (3, 3)),
            :parameter_map => (S = [:Sa, :Sb], J = vcat([:JD], repeat([:JP], 8))), :JD => 0.5, :JP => 1,
            :Sa => 1 // 2, :Sb => 1 // 2, :thermalization => 10^9, :sweeps => 1, :binsize => 1), outdir)
end

main(length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "reference_dump"))
