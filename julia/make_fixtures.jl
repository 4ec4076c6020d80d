# make_fixtures.jl — golden vectors from the STOCK reference (sisl/MPOPIS) for tests/test_julia_fixtures.py.
#
# The reference ships no tests and Julia is not installed in the build image, so the CPU oracle (oracle/) is pinned
# only to published formulas ("parity unpinned", DESIGN.md §1). A maintainer with a Julia install closes that gap by
# running this script once from the reference checkout and committing the file it writes:
#
#     julia --project=/path/to/MPOPIS julia/make_fixtures.jl tests/golden/julia_v1.json
#
# Every record stores the inputs together with the reference's outputs, so the Python side needs no Julia RNG: the
# noise E is generated here with a fixed seed and written out (SURVEY App. G; third-party boundaries of SURVEY §8c:
# CovarianceEstimation, StatsBase.mean_and_cov, LinearAlgebra Σ^-0.5, Base.sortperm, RLEnvs MountainCar).
# Nothing in this file is used by the product or the oracle; it only CALLS the reference.
using MPOPIS, Random, LinearAlgebra, Statistics
import StatsBase, CovarianceEstimation, Distributions
import ReinforcementLearning: reset!, reward, state, action_space

const OUT = length(ARGS) >= 1 ? ARGS[1] : "julia_v1.json"

# --- a dependency-free JSON writer (numbers, strings, bools, vectors, matrices as column-major {"dims","data"}) ---
js(x::Bool) = x ? "true" : "false"
js(x::Integer) = string(x)
js(x::AbstractFloat) = isnan(x) ? "\"NaN\"" : isinf(x) ? (x > 0 ? "\"Inf\"" : "\"-Inf\"") : repr(Float64(x))
js(x::AbstractString) = "\"" * escape_string(x) * "\""
js(x::Symbol) = js(String(x))
js(x::AbstractVector) = "[" * join([js(v) for v in x], ",") * "]"
js(x::AbstractMatrix) = "{\"dims\":[$(size(x, 1)),$(size(x, 2))],\"data\":" * js(vec(collect(x))) * "}"
js(x::AbstractArray{<:Any,3}) = "{\"dims\":[$(size(x, 1)),$(size(x, 2)),$(size(x, 3))],\"data\":" * js(vec(collect(x))) * "}"
js(x::Dict) = "{" * join([js(String(k)) * ":" * js(v) for (k, v) in sort(collect(x), by = p -> String(p[1]))], ",") * "}"
js(x::NamedTuple) = js(Dict(pairs(x)))
js(x::Tuple) = js(collect(x))

fx = Dict{String,Any}("julia_version" => string(VERSION), "generator" => "julia/make_fixtures.jl v1")
rng = MersenneTwister(20261017)

# 1. _step! known answers (CAR:282-344): states x actions -> state after 1 and after 50 steps
let recs = Any[]
    env = CarRacingEnv()
    s0s = [[0.0, 0.0, pi / 2, 10.0, 0.0, 0.0, 0.0, 0.0], [20.0, -5.0, 3.1, 2.0, 0.5, 0.4, 0.1, 0.0],
           [20.0, 0.0, -3.12, -3.0, 0.2, -0.5, 0.0, 0.0], [50.0, 10.0, 0.3, 30.0, -1.0, 0.2, -0.3, 0.0]]
    acts = [[a, b] for a in (-1.0, -0.3, 0.0, 0.7, 1.0) for b in (-1.0, 0.0, 0.5, 1.0)]
    for s0 in s0s, a in acts
        reset!(env); env.state = copy(s0)
        env(a); s1 = copy(env.state); r1 = reward(env)
        for _ in 2:50; env(a); end
        push!(recs, Dict("state0" => s0, "action" => a, "state1" => s1, "reward1" => r1, "state50" => copy(env.state),
                         "reward50" => reward(env)))
    end
    fx["car_step"] = recs
end

# 2. within_track known answers (TRK:68-92): integer columns are bit-exact targets
let env = CarRacingEnv(), tr = env.track
    pos = [[x, y] for x in range(1.0, 255.0, length = 60) for y in range(-156.0, 141.0, length = 60)]
    append!(pos, [[tr.x′[i] + 1e-9, tr.y′[i] - 1e-9] for i in eachindex(tr.x′)])
    within = Bool[]; dist = Float64[]; idx = Int[]
    for p in pos
        w = MPOPIS.within_track(tr, p)
        push!(within, w.within); push!(dist, w.dist)
        push!(idx, argmin((tr.x′ .- p[1]) .^ 2 .+ (tr.y′ .- p[2]) .^ 2) - 1)   # 0-based, TRK:71-73
    end
    fx["within_track"] = Dict("track_x" => tr.x′, "track_y" => tr.y′, "track_w" => tr.lane_width′,
                              "pos" => reduce(hcat, pos), "within" => within, "dist" => dist, "min_idx0" => idx)
end

# 3./4. simulate_model + compute_weights (POL:261-278, UTL:79-86) for 1 car and 3 cars on a written-out E
for (name, mk, K) in (("car1", () -> CarRacingEnv(), 64), ("car3", () -> MultiCarRacingEnv(3), 48))
    env = mk()
    pol = CEMPPI_Policy(env; num_samples = K, horizon = 50, λ = 10.0, α = 1.0, U₀ = zeros(MPOPIS.action_space_size(action_space(env))),
                        cov_mat = block_diagm([0.0625, 0.1], name == "car1" ? 1 : 3), opt_its = 10, ce_elite_threshold = 0.8,
                        Σ_est = :ss, rng = MersenneTwister(1))
    cs = pol.params.cs
    E = randn(rng, cs, K) .* repeat([0.25, sqrt(0.1)], cs ÷ 2)
    costs = MPOPIS.simulate_model(pol, env, E, inv(Matrix(pol.Σ)), copy(pol.U))
    fx["simulate_model_" * name] = Dict("state" => copy(env.state), "U" => copy(pol.U), "E" => E, "costs" => costs,
        "weights" => Dict(string(l) => MPOPIS.compute_weights(MPOPIS.Information_Theoretic(l), costs) for l in (0.1, 10.0, 20.0)),
        "sortperm0" => sortperm(costs) .- 1)
end

# 5./6. covariance estimators on elite-like (n = 30) and tall (n = 819) sets, p = 100; weighted moments; Σ^-0.5
let p = 100
    for n in (30, 819)
        A = randn(rng, p, p) ./ 10
        X = A * randn(rng, p, n) .+ 0.1 .* randn(rng, p)          # p x n, columns = observations (= `elite`)
        rec = Dict{String,Any}("X" => X)
        for (sym, est) in ((:mle, CovarianceEstimation.SimpleCovariance()),
                           (:lw, CovarianceEstimation.LinearShrinkage(CovarianceEstimation.DiagonalUnequalVariance(), :lw)),
                           (:ss, CovarianceEstimation.LinearShrinkage(CovarianceEstimation.DiagonalUnequalVariance(), :ss)),
                           (:rblw, CovarianceEstimation.LinearShrinkage(CovarianceEstimation.DiagonalCommonVariance(), :rblw)),
                           (:oas, CovarianceEstimation.LinearShrinkage(CovarianceEstimation.DiagonalCommonVariance(), :oas)))
            rec["cov_" * String(sym)] = Matrix(cov(est, X'))             # POL:464 (without the 10e-9 ridge)
        end
        w = rand(rng, n); w ./= sum(w)
        μw, Σw = StatsBase.mean_and_cov(X, StatsBase.ProbabilityWeights(w), 2)   # POL:364,662,732
        μu, Σu = StatsBase.mean_and_cov(X, 2)                                    # POL:807
        rec["w"] = w; rec["mean_w"] = vec(μw); rec["cov_w"] = Matrix(Σw); rec["mean_u"] = vec(μu); rec["cov_u"] = Matrix(Σu)
        fx["cov_n$(n)"] = rec
    end
    S = let B = randn(rng, 40, 40); Symmetric(B * B' ./ 40 + 0.1I) end
    fx["inv_sqrt"] = Dict("A" => Matrix(S), "C" => Matrix(Matrix(S)^-0.5),                 # POL:580
                          "L" => Matrix(cholesky(S).L), "invcov" => Matrix(Distributions.invcov(Distributions.MvNormal(Matrix(S)))))
end

# 7./8. full control steps with the noise INTERCEPTED. A more specific method of simulate_model (env::CarRacingEnv) records
# the E and Σ⁻¹ every AIS iteration hands to it (POL:452) and then invokes the stock method, so the reference's own code
# path runs unchanged. The Python side injects exactly these draws: Z_n = L_n⁻¹ E_n with L_n = chol(inv(Σ⁻¹_n)), and
# checks per-iteration costs, the final control and the rolled U of two consecutive steps (pins SURVEY App. B-1/B-2).
# :mppi keeps its rollouts inside calculate_trajectory_costs (POL:186-216, no simulate_model call) and :pmcmppi draws
# its resampling indices from Julia's alias sampler between iterations, so those two are pinned through records 3-6.
const E_TAP = Matrix{Float64}[]
const SINV_TAP = Matrix{Float64}[]
const COST_TAP = Vector{Float64}[]
function MPOPIS.simulate_model(pol::MPOPIS.AbstractGMPPI_Policy, env::CarRacingEnv, E::Matrix{Float64},
                               Σ_inv::Matrix{Float64}, U_orig::Vector{Float64})
    c = invoke(MPOPIS.simulate_model,
               Tuple{MPOPIS.AbstractGMPPI_Policy,MPOPIS.AbstractEnv,Matrix{Float64},Matrix{Float64},Vector{Float64}},
               pol, env, E, Σ_inv, U_orig)
    push!(E_TAP, copy(E)); push!(SINV_TAP, copy(Σ_inv)); push!(COST_TAP, copy(c))
    return c
end
for sym in (:gmppi, :imppi, :cemppi, :cmamppi, :μaismppi, :μΣaismppi)
    env = CarRacingEnv()
    # get_policy(policy_type, env, num_samples, horizon, λ, α, U₀, cov_mat, pol_log, ais_its, λ_ais, ce_elite_threshold,
    #            ce_Σ_est, cma_σ, cma_elite_threshold)   example_utils.jl:12-19
    pol = MPOPIS.get_policy(sym, env, 64, 20, 10.0, 1.0, zeros(2), block_diagm([0.0625, 0.1], 1), true, 4, 20.0, 0.8, :ss,
                            0.75, 0.8)
    seed!(pol, 7)
    steps = Any[]
    for t in 1:2
        empty!(E_TAP); empty!(SINV_TAP); empty!(COST_TAP)
        U_before = copy(pol.U); state_before = copy(env.state)
        act = pol(env)
        push!(steps, Dict("state" => state_before, "U_before" => U_before, "control" => vec(collect(act)),
                          "U_after" => copy(pol.U), "its" => length(E_TAP),
                          "E" => [copy(e) for e in E_TAP], "Sigma_inv" => [copy(s) for s in SINV_TAP],
                          "costs" => [copy(c) for c in COST_TAP]))
        env(act)
    end
    fx["policy_" * String(sym)] = Dict("steps" => steps, "state_final" => copy(env.state), "K" => 64, "T" => 20, "N" => 4)
end

# 9. MountainCar (RLEnvs dynamics + EXM:10-22 reward): 200 steps under a fixed action sequence
let env = MountainCarEnv(continuous = true, max_steps = 200, rng = MersenneTwister(3))
    reset!(env); env.state = [-0.5, 0.0]
    acts = [sin(0.13 * t) for t in 1:200]; xs = Float64[]; vs = Float64[]; rs = Float64[]
    for a in acts
        env([a]); push!(xs, env.state[1]); push!(vs, env.state[2]); push!(rs, reward(env))
    end
    fx["mountaincar"] = Dict("actions" => acts, "x" => xs, "v" => vs, "reward" => rs)
end

# 10. documentation only: the first normals of this Julia version's MersenneTwister (stream parity is out of scope)
fx["randn_first16"] = randn(MersenneTwister(1), 16)

open(OUT, "w") do io
    write(io, js(fx))
end
println("wrote ", OUT)
