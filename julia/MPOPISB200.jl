# MPOPISB200.jl — Julia shim that plugs libmpopis_b200.so into an UNMODIFIED MPOPIS.jl.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image (SURVEY.md §0.2). It is
# the reference-side binding a maintainer adds; tests exercise the same C-ABI calls, in the same order,
# through the Python ctypes mirror (mpopis_b200/engine.py, policies.py).
#
# Mechanism: the same one MPOPIS already uses for its EnvPool backend — more specific methods of the
# policy functor / simulate_model for concrete env types (mppi_mpopi_policies.jl:148,240; utils.jl:103).
# Nothing in MPOPIS is edited; `using MPOPISB200` after `using MPOPIS` is enough:
#
#     using MPOPIS, MPOPISB200
#     MPOPISB200.enable!()                       # route CarRacing / MultiCar / MountainCar policies to the GPU
#     simulate_car_racing(policy_type=:cemppi, num_samples=65536)
#
module MPOPISB200

using MPOPIS
using Random
import MPOPIS: AbstractGMPPI_Policy, MPPI_Policy, AbstractPathIntegralPolicy, CarRacingEnv, MultiCarRacingEnv,
               simulate_model
import ReinforcementLearning: MountainCarEnv

const LIB = Ref{String}(get(ENV, "MPOPIS_B200_LIB", "libmpopis_b200.so"))
const ABI_VERSION = 1

# ---- include/mpopis_b200.h -------------------------------------------------------------------------------
struct Cfg                       # mpopis_cfg_t (field order and types must match the header)
    abi_version::Int32; policy::Int32; env::Int32; n_cars::Int32
    num_samples::Int64; horizon::Int64; opt_its::Int64
    lambda::Float64; alpha::Float64; lambda_ais::Float64; ce_elite_threshold::Float64
    sigma_est::Int32; early_stop::Int32; log_trajectories::Int32
    device::Int32; rank::Int32; world_size::Int32
    ext_action_size::Int32           # MPOPIS_ENV_EXTERNAL only
    reserved::NTuple{3,Int32}
end
struct Cma                       # mpopis_cma_t
    sigma::Float64; m_elite::Int64; mu_eff::Float64; c_sigma::Float64; d_sigma::Float64
    c_Sigma::Float64; c1::Float64; c_mu::Float64; E_norm::Float64
end

last_error() = unsafe_string(ccall((:mpopis_b200_last_error, LIB[]), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("mpopis_b200 ($rc): $(last_error())")   # MPOPIS-style error(...)

const POLICY = Dict(MPOPIS.MPPI_Policy => 0, MPOPIS.GMPPI_Policy => 1, MPOPIS.IMPPI_Policy => 2,
                    MPOPIS.CEMPPI_Policy => 3, MPOPIS.CMAMPPI_Policy => 4, MPOPIS.μAISMPPI_Policy => 5,
                    MPOPIS.μΣAISMPPI_Policy => 6, MPOPIS.PMCMPPI_Policy => 7)
policy_code(pol) = POLICY[Base.typename(typeof(pol)).wrapper]
# policies the engine does not implement (NESMPPI_Policy: unreachable from get_policy) keep the stock Julia path
supported(pol) = haskey(POLICY, Base.typename(typeof(pol)).wrapper)

function sigma_est_code(pol)
    pol isa MPOPIS.CEMPPI_Policy || return Int32(0)
    m = pol.Σ_estimation_method
    m isa MPOPIS.SimpleCovariance && return Int32(0)
    s = m.shrinkage                                   # :lw, :ss, :rblw, :oas (mppi_mpopi_policies.jl:414-426)
    return Int32(Dict(:lw => 1, :ss => 2, :rblw => 3, :oas => 4)[s])
end

# ---- handles, one per policy object ------------------------------------------------------------------------
# Weak keys: the table must not keep a policy alive (the reference examples build a new policy per trial), otherwise
# the finalizer below never runs and device memory / streams / events leak. `close!(pol)` frees a handle eagerly.
mutable struct Handle
    ptr::Ptr{Cvoid}
end
const HANDLES = WeakKeyDict{Any,Handle}()
function destroy!(hd::Handle)
    if hd.ptr != C_NULL
        ccall((:mpopis_b200_destroy, LIB[]), Cint, (Ptr{Cvoid},), hd.ptr)
        hd.ptr = C_NULL
    end
    nothing
end
"Release the engine handle of `pol` now (otherwise it is released when `pol` is garbage-collected)."
close!(pol) = haskey(HANDLES, pol) ? (destroy!(HANDLES[pol]); delete!(HANDLES, pol); nothing) : nothing

env_code(::CarRacingEnv) = (Int32(0), Int32(1))
env_code(e::MultiCarRacingEnv) = (Int32(0), Int32(e.N))
env_code(::MountainCarEnv) = (Int32(1), Int32(0))

car_params(e::CarRacingEnv) = Float64[getfield(e.params, f) for f in fieldnames(typeof(e.params))]  # 18, declaration order

function set_env!(h, e::CarRacingEnv)
    p = car_params(e); t = e.track
    check(ccall((:mpopis_b200_set_car_env, LIB[]), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
        h, 1, p, e.dt, e.δt, t.x′, t.y′, t.lane_width′, length(t.x′)))
end
function set_env!(h, e::MultiCarRacingEnv)
    # Every sub-env owns a Track built from the same file and sample factor (multi-car_racing.jl:37-45), and the engine
    # takes ONE track per handle. Refuse anything else rather than silently using car 1's centre line for all cars.
    t = e.envs[1].track
    for sub in e.envs[2:end]
        (sub.track.x′ == t.x′ && sub.track.y′ == t.y′ && sub.track.lane_width′ == t.lane_width′) ||
            error("MPOPISB200: MultiCarRacingEnv sub-envs with different tracks are not supported by the engine")
    end
    p = reduce(vcat, car_params.(e.envs))
    check(ccall((:mpopis_b200_set_car_env, LIB[]), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
        h, e.N, p, e.dt, e.δt, t.x′, t.y′, t.lane_width′, length(t.x′)))
end
function set_env!(h, e::MountainCarEnv)
    q = e.params
    p = Float64[q.min_pos, q.max_pos, q.max_speed, q.goal_pos, q.goal_velocity, q.power, q.gravity]
    check(ccall((:mpopis_b200_set_mountaincar_env, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h, p, q.max_steps))
end

# Sharded policies (BASELINE config 5: K = 2^20 over the 8 GPUs of a box): one Julia process per GPU, e.g. under
# MPI.jl / Distributed, each with `MPOPISB200.shard!(rank, world, device, nccl_id)` BEFORE the first control step;
# `nccl_id = comm_id()` on rank 0, broadcast by the host program. Every rank then calls pol(env) with the same state
# and gets the same control (the engine all-gathers the costs and all-reduces the moments over NCCL).
const SHARD = Ref{Any}(nothing)     # (rank, world, device, id::Vector{UInt8}) or nothing
function comm_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:mpopis_b200_comm_id, LIB[]), Cint, (Ptr{UInt8},), id))
    return id
end
shard!(rank::Integer, world::Integer, device::Integer, id::Vector{UInt8}) = (SHARD[] = (Int32(rank), Int32(world), Int32(device), id); nothing)

function handle(pol::AbstractPathIntegralPolicy, env; device=0)
    hd = get!(HANDLES, pol) do
        ecode, ncars = env_code(env)
        rank, world = SHARD[] === nothing ? (Int32(0), Int32(1)) : (SHARD[][1], SHARD[][2])
        device = SHARD[] === nothing ? device : SHARD[][3]
        P = pol.params
        cfg = Cfg(ABI_VERSION, policy_code(pol), ecode, ncars, P.num_samples, P.horizon,
                  hasproperty(pol, :opt_its) ? pol.opt_its : 1, P.λ, P.α,
                  hasproperty(pol, :λ_ais) ? pol.λ_ais : 20.0,
                  hasproperty(pol, :ce_elite_threshold) ? pol.ce_elite_threshold : 0.8,
                  sigma_est_code(pol), 1, P.log ? 1 : 0, device, rank, world,
                  ecode == 2 ? Int32(P.as) : Int32(0), (Int32(0), Int32(0), Int32(0)))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mpopis_b200_create, LIB[]), Cint, (Ref{Cfg}, Ref{Ptr{Cvoid}}), cfg, out))
        h = out[]
        SHARD[] === nothing || check(ccall((:mpopis_b200_comm_init, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h, SHARD[][4]))
        set_env!(h, env)
        Σ = Matrix{Float64}(pol.Σ)                   # as x as for :mppi, cs x cs otherwise; column-major as is
        check(ccall((:mpopis_b200_set_sigma, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h, Σ, size(Σ, 1)))
        if pol isa MPOPIS.CMAMPPI_Policy
            cma = Cma(pol.σ, pol.m_elite, pol.μ_eff, pol.cσ, pol.dσ, pol.cΣ, pol.c1, pol.cμ, pol.E)
            check(ccall((:mpopis_b200_set_cma, LIB[]), Cint, (Ptr{Cvoid}, Ref{Cma}, Ptr{Float64}, Int64),
                        h, cma, pol.ws, length(pol.ws)))
        end
        # Random.seed!(pol, s) seeds pol.rng (MPOPIS.jl:54); draw the engine's Philox key from that stream so
        # that `seed!(pol, seed + k)` keeps controlling reproducibility.
        check(ccall((:mpopis_b200_seed, LIB[]), Cint, (Ptr{Cvoid}, UInt64), h, rand(pol.rng, UInt64)))
        hd = Handle(h)
        finalizer(destroy!, hd)       # runs when the table's weak key (the policy) is collected and the Handle dies
        hd
    end
    return hd.ptr
end

# Peer-memory collectives on one NVLink node (include/mpopis_b200.h, csrc/comm.cu): every rank exports 128 bytes, the
# host program all-gathers them in rank order (e.g. MPI.Allgather) and every rank attaches the world x 128 bytes; the
# per-iteration exchanges then run as single kernels over NVLink instead of NCCL calls.
function peer_export(pol::AbstractPathIntegralPolicy, env)
    blob = Vector{UInt8}(undef, 128)
    check(ccall((:mpopis_b200_comm_peer_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}), handle(pol, env), blob))
    return blob
end
peer_attach!(pol::AbstractPathIntegralPolicy, env, blobs::Vector{UInt8}) =
    check(ccall((:mpopis_b200_comm_peer_attach, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), handle(pol, env), blobs, length(blobs)))

env_t(e) = Int64(hasproperty(e, :t) ? e.t : 0)

# ---- depth (iii): the whole functor, mppi_mpopi_policies.jl:121-146 and 221-238 ------------------------------
function plan!(pol::AbstractPathIntegralPolicy, env)
    h = handle(pol, env)
    control = Vector{Float64}(undef, pol.params.as)
    its = Ref{Int32}(0)
    s = Vector{Float64}(env.state)
    GC.@preserve s control begin
        check(ccall((:mpopis_b200_plan, LIB[]), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
            h, s, env_t(env), pol.U, control, its))      # pol.U is rolled in place (aliases params.U₀, App. B-2)
    end
    if pol.params.log
        K, T, ss = pol.params.num_samples, pol.params.horizon, pol.params.ss
        traj = Array{Float64}(undef, T, ss, K)
        check(ccall((:mpopis_b200_fetch, LIB[]), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, pol.logger.traj_costs, pol.logger.traj_weights, C_NULL, traj))
        for k in 1:K
            pol.logger.trajectories[k] .= @view traj[:, :, k]
        end
    end
    # get_model_controls returns a Vector for as == 1 and an as x 1 Matrix otherwise (utils.jl:63-66)
    return pol.params.as == 1 ? control : reshape(control, :, 1)
end

# ---- depth (i): simulate_model with Julia's own noise E (exact same-seed drop-in, small K) --------------------
function simulate_model_b200(pol::AbstractGMPPI_Policy, env, E::Matrix{Float64}, Σ_inv::Matrix{Float64},
                             U_orig::Vector{Float64})
    h = handle(pol, env)
    costs = Vector{Float64}(undef, pol.params.num_samples)
    s = Vector{Float64}(env.state)
    check(ccall((:mpopis_b200_rollout_costs, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h, s, env_t(env), pol.U, U_orig, E, Σ_inv, costs))
    return costs
end

# ---- the EnvpoolEnv seam (MPOPIS_ENV_EXTERNAL): MPOPIS keeps the simulator, the engine does the rest -----------------
# Replaces calculate_trajectory_costs(pol::MPPI_Policy, env::EnvpoolEnv) (mppi_mpopi_policies.jl:148-184) and
# simulate_model(pol::AbstractGMPPI_Policy, env::EnvpoolEnv, ...) (:240-259): the callback receives the K x as x T
# array get_model_controls builds (utils.jl:42-53) and answers with rollout_model's trajectory costs (utils.jl:103-121).
env_code(::MPOPIS.EnvpoolEnv) = (Int32(2), Int32(1))
const CURRENT = Ref{Any}(nothing)                      # (pol, env) of the running plan_external
function rollout_cb(::Ptr{Cvoid}, controls::Ptr{Float64}, K::Int64, as::Int64, T::Int64, out::Ptr{Float64})::Cint
    try
        pol, env = CURRENT[]
        cost = MPOPIS.rollout_model(env, Int(T), unsafe_wrap(Array, controls, (Int(K), Int(as), Int(T))), pol)
        unsafe_copyto!(out, pointer(cost), K)
        return Cint(0)
    catch
        return Cint(1)                                 # never unwind through the C frames
    end
end
function set_env!(h, env::MPOPIS.EnvpoolEnv)
    lo = Vector{Float64}(leftendpoint(action_space(env))); hi = Vector{Float64}(rightendpoint(action_space(env)))
    check(ccall((:mpopis_b200_set_external_env, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, lo, hi))
end
function plan_external!(pol, env::MPOPIS.EnvpoolEnv)
    h = handle(pol, env)
    cb = @cfunction(rollout_cb, Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}))
    control = Vector{Float64}(undef, pol.params.as); its = Ref{Int32}(0)
    CURRENT[] = (pol, env)
    try
        check(ccall((:mpopis_b200_plan_external, LIB[]), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
            h, pol.U, cb, C_NULL, C_NULL, C_NULL, control, its))
    finally
        CURRENT[] = nothing
    end
    return pol.params.as == 1 ? control : reshape(control, :, 1)
end
"Route `pol(env::EnvpoolEnv)` through the engine (sampling / adaptation / weights on the GPU, rollouts in EnvPool)."
function enable_external!()
    @eval (pol::AbstractGMPPI_Policy)(env::MPOPIS.EnvpoolEnv) = plan_external!(pol, env)
    @eval (pol::MPPI_Policy)(env::MPOPIS.EnvpoolEnv) = plan_external!(pol, env)
    nothing
end

# ---- dispatch: enable!() defines the more specific methods -----------------------------------------------------
"""
    enable!(; depth = :functor)

`depth = :functor` overrides `(pol)(env)` for CarRacingEnv / MultiCarRacingEnv / MountainCarEnv (whole control step
on the GPU, engine RNG). `depth = :simulate_model` overrides only `simulate_model` (Julia keeps drawing `E` with
`pol.rng`: bit-for-bit the same sampling as stock MPOPIS, costs from the GPU).
"""
function enable!(; depth::Symbol=:functor)
    for Env in (CarRacingEnv, MultiCarRacingEnv, MountainCarEnv)
        if depth == :functor
            # unsupported policy types (e.g. NESMPPI_Policy) fall through to the stock method of the abstract env type
            @eval (pol::AbstractGMPPI_Policy)(env::$Env) =
                supported(pol) ? plan!(pol, env) : invoke(pol, Tuple{MPOPIS.AbstractEnv}, env)
            @eval (pol::MPPI_Policy)(env::$Env) = plan!(pol, env)
        elseif depth == :simulate_model
            @eval MPOPIS.simulate_model(pol::AbstractGMPPI_Policy, env::$Env, E::Matrix{Float64},
                                        Σ_inv::Matrix{Float64}, U_orig::Vector{Float64}) =
                supported(pol) ? simulate_model_b200(pol, env, E, Σ_inv, U_orig) :
                invoke(MPOPIS.simulate_model,
                       Tuple{AbstractGMPPI_Policy,MPOPIS.AbstractEnv,Matrix{Float64},Matrix{Float64},Vector{Float64}},
                       pol, env, E, Σ_inv, U_orig)
        else
            error("depth must be :functor or :simulate_model")
        end
    end
    nothing
end

end # module
