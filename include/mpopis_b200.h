/*
 * mpopis_b200.h — C-ABI of the B200-native MPPI/MPOPI sampling engine.
 *
 * This is the drop-in boundary for the one data-parallel hot path of sisl/MPOPIS
 * (reference files, relative to the reference root):
 *   POL = src/mppi_mpopi_policies.jl   UTL = src/utils.jl
 *   CAR = src/envs/car_racing.jl       TRK = src/envs/car_racing_tracks/car_racing_tracks.jl
 *   MCR = src/envs/multi-car_racing.jl EXM = src/examples/mountaincar_example.jl
 *
 * The reference has no FFI; its plugin seam is Julia multiple dispatch on the env type
 * (the EnvpoolEnv backend overrides POL:148, POL:240, UTL:103). A Julia package binds the
 * entry points below with `ccall` and adds more specific methods of the same three
 * functions. Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - C linkage, plain pointers and sizes; no C++ types or exceptions cross the boundary.
 *   - All arrays are dense column-major Float64 (Julia `Matrix` layout) unless stated.
 *     A control vector of length cs = as*T is ordered [a_1(t=1)..a_as(t=1), a_1(t=2), ...]
 *     (POL:59-63, UTL:59). The noise matrix E is cs x K: E[r,k] at r + cs*k (POL:271).
 *   - The caller owns every host buffer; the library touches it only during the call.
 *   - Return value 0 = OK, negative = error (see mpopis_status_t);
 *     mpopis_b200_last_error() returns a thread-local description.
 *   - There is NO CPU fallback: without a CUDA device of compute capability 10.x
 *     mpopis_b200_create() fails with MPOPIS_ERR_NO_DEVICE.
 *   - A handle is not thread-safe; different handles may be used from different threads.
 */
#ifndef MPOPIS_B200_H
#define MPOPIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPOPIS_B200_ABI_VERSION 1
#define MPOPIS_MAX_CARS 8
#define MPOPIS_CAR_NPARAMS 18 /* CAR:2-21, declaration order */
#define MPOPIS_MC_NPARAMS 7   /* min_pos,max_pos,max_speed,goal_pos,goal_velocity,power,gravity */

typedef enum {
  MPOPIS_OK = 0,
  MPOPIS_ERR_BAD_ARG = -1, /* mirrors the reference's error(...) on bad config: POL:55,64,73,79,89 */
  MPOPIS_ERR_CUDA = -2,
  MPOPIS_ERR_NCCL = -3,
  MPOPIS_ERR_NOT_PD = -4,   /* Cholesky failed: mirrors Julia's PosDefException inside MvNormal(Σ) */
  MPOPIS_ERR_NO_DEVICE = -5 /* no sm_100 device: the engine never falls back to the CPU */
} mpopis_status_t;

/* get_policy symbols, src/examples/example_utils.jl:20-128 */
typedef enum {
  MPOPIS_POLICY_MPPI = 0,        /* :mppi       POL:116-146,186-216 */
  MPOPIS_POLICY_GMPPI = 1,       /* :gmppi      POL:298-315 */
  MPOPIS_POLICY_IMPPI = 2,       /* :imppi      POL:337-373 */
  MPOPIS_POLICY_CEMPPI = 3,      /* :cemppi     POL:407-472 */
  MPOPIS_POLICY_CMAMPPI = 4,     /* :cmamppi    POL:506-606 */
  MPOPIS_POLICY_MUAISMPPI = 5,   /* :μaismppi   POL:630-671 */
  MPOPIS_POLICY_MUSIGMAAISMPPI = 6, /* :μΣaismppi POL:695-742 */
  MPOPIS_POLICY_PMCMPPI = 7      /* :pmcmppi    POL:766-817 */
} mpopis_policy_t;

typedef enum {
  MPOPIS_ENV_CAR_RACING = 0,  /* CarRacingEnv (n_cars = 1, CAR) or MultiCarRacingEnv (n_cars > 1, MCR) */
  MPOPIS_ENV_MOUNTAIN_CAR = 1, /* RLEnvs MountainCarEnv(continuous=true) + EXM:4-22 */
  MPOPIS_ENV_EXTERNAL = 2      /* the caller's own batched simulator behind the EnvpoolEnv seam (POL:148,240; UTL:103):
                                  the engine samples / adapts / weights, a callback rolls the K controls out */
} mpopis_env_t;

/* CEMPPI_Policy Σ_est, POL:414-426 */
typedef enum {
  MPOPIS_SIGMA_MLE = 0,  /* SimpleCovariance() */
  MPOPIS_SIGMA_LW = 1,   /* LinearShrinkage(DiagonalUnequalVariance(), :lw) */
  MPOPIS_SIGMA_SS = 2,   /* LinearShrinkage(DiagonalUnequalVariance(), :ss) */
  MPOPIS_SIGMA_RBLW = 3, /* LinearShrinkage(DiagonalCommonVariance(), :rblw) */
  MPOPIS_SIGMA_OAS = 4   /* LinearShrinkage(DiagonalCommonVariance(), :oas) */
} mpopis_sigma_est_t;

/* Policy + problem description: the fields of MPPI_Policy_Params (POL:8-19) and of the
 * per-policy structs (POL:107-114, 321-329, 379-389, 478-496, 612-621, 677-686, 748-757). */
typedef struct mpopis_cfg {
  int32_t abi_version;        /* = MPOPIS_B200_ABI_VERSION */
  int32_t policy;             /* mpopis_policy_t */
  int32_t env;                /* mpopis_env_t */
  int32_t n_cars;             /* 1..MPOPIS_MAX_CARS for CAR_RACING, ignored otherwise */
  int64_t num_samples;        /* K  (pol.params.num_samples) */
  int64_t horizon;            /* T  (pol.params.horizon) */
  int64_t opt_its;            /* N  (pol.opt_its; 1 for :mppi/:gmppi) */
  double lambda;              /* λ  (pol.params.λ) */
  double alpha;               /* α  (pol.params.α); γ = λ(1-α), POL:266 */
  double lambda_ais;          /* λ_ais (POL:618,683,754) */
  double ce_elite_threshold;  /* POL:385; m_elite = round(Int, K*(1-thr)), POL:437 */
  int32_t sigma_est;          /* mpopis_sigma_est_t, POL:386 */
  int32_t early_stop;         /* 1 = reference behaviour (POL:459-461, 567-569); 0 = always run N iterations */
  int32_t log_trajectories;   /* pol.params.log: keep K x T x ss states for fetch() (UTL:139-141) */
  int32_t device;             /* CUDA device ordinal */
  int32_t rank;               /* shard index of this handle, 0 <= rank < world_size */
  int32_t world_size;         /* number of handles sharing the K samples (1 = unsharded) */
  int32_t ext_action_size;    /* MPOPIS_ENV_EXTERNAL only: as = action_space_size(action_space(env)), UTL:2-7 */
  int32_t reserved[3];
} mpopis_cfg_t;

/* CMA-ES constants computed by the CMAMPPI_Policy constructor (POL:513-525). The host side
 * (Julia shim / Python mirror) evaluates those formulas and passes the results. */
typedef struct mpopis_cma {
  double sigma;     /* pol.σ   */
  int64_t m_elite;  /* pol.m_elite */
  double mu_eff;    /* pol.μ_eff */
  double c_sigma;   /* pol.cσ */
  double d_sigma;   /* pol.dσ */
  double c_Sigma;   /* pol.cΣ */
  double c1;        /* pol.c1 */
  double c_mu;      /* pol.cμ */
  double E_norm;    /* pol.E  */
} mpopis_cma_t;

typedef struct mpopis_handle mpopis_t;

int mpopis_b200_abi_version(void);
const char *mpopis_b200_last_error(void);

/* Replaces the policy constructors (POL:116,298,337,407,506,630,695,766): allocates all
 * device state for the given sizes. Σ defaults to the identity until set_sigma(). */
int mpopis_b200_create(const mpopis_cfg_t *cfg, mpopis_t **out);
int mpopis_b200_destroy(mpopis_t *h);

/* Join `world_size` handles (one per process/GPU) into one sharded policy. `nccl_id` is the
 * 128-byte ncclUniqueId produced by mpopis_b200_comm_id() on rank 0 and broadcast by the host. */
int mpopis_b200_comm_id(void *nccl_id_out128);
int mpopis_b200_comm_init(mpopis_t *h, const void *nccl_id128);

/* Loop-back communicator: `world` VIRTUAL ranks on one device in one process — handles created with world_size =
 * world, rank = 0..world-1 and the same device, attached to one group, each then driven from its own host thread (a
 * collective blocks until every virtual rank has entered it). It exists so that the sharded code path (all-gathered
 * costs -> identical elite selection on every rank, ownership-compacted elite moments, fixed-order reductions:
 * SURVEY §8e; the reference itself has no multi-device path, POL:269 is its only parallel loop) can be verified on a
 * single-GPU box against the unsharded engine and the oracle. Production sharding uses comm_init (NCCL). */
int mpopis_b200_loopback_create(int32_t world, void **group_out);
int mpopis_b200_loopback_destroy(void *group);
int mpopis_b200_comm_init_loopback(mpopis_t *h, void *group);

/* Peer-memory collectives (one node, NVLink / NVSwitch): after comm_init, every rank exports 128 bytes (two CUDA IPC
 * handles: its control region and its cost vector), the host all-gathers the blobs in rank order (any transport) and
 * every rank attaches them. From then on the per-iteration exchanges (all-gather of 8 B x K_loc costs, all-reduce of
 * the 2cs+1 / cs²+1 moment sums) are single kernels that store into the peers' memory and spin on arrival flags
 * (csrc/comm.cu) instead of NCCL calls; NCCL keeps serving the oversize all-reduce of :cmamppi. The reference has no
 * counterpart (its only parallel loop is Threads.@threads over the samples, POL:269). comm_peer_loopback does the
 * same between the virtual ranks of a loop-back group (collective: call it from every rank's thread; the process must
 * run with CUDA_MODULE_LOADING=EAGER because the virtual ranks share one CUDA context — see csrc/comm.cu). */
int mpopis_b200_comm_peer_export(mpopis_t *h, void *handle_out128);
int mpopis_b200_comm_peer_attach(mpopis_t *h, const void *handles, int64_t n_bytes);
int mpopis_b200_comm_peer_loopback(mpopis_t *h);

/* env.params (CAR:2-21, 18 doubles per car in declaration order), env.dt, env.δt (CAR:33-34),
 * env.track.x′, y′, lane_width′ (TRK:6-8). n_cars > 1 is MultiCarRacingEnv (MCR:2-12). */
int mpopis_b200_set_car_env(mpopis_t *h, int32_t n_cars, const double *params18_per_car, double dt,
                            double ddt, const double *trk_x, const double *trk_y,
                            const double *trk_w, int64_t n_trk);
/* RLEnvs MountainCarEnv params {min_pos,max_pos,max_speed,goal_pos,goal_velocity,power,gravity}
 * and max_steps. */
int mpopis_b200_set_mountaincar_env(mpopis_t *h, const double *params7, int64_t max_steps);

/* MPOPIS_ENV_EXTERNAL: the action bounds leftendpoint/rightendpoint(action_space(env)) (as doubles each) that
 * get_model_controls clamps to (UTL:31-32,42-43). */
int mpopis_b200_set_external_env(mpopis_t *h, const double *action_lo, const double *action_hi);

/* pol.Σ: n = as (expanded with block_diagm over the horizon, UTL:9-21, POL:76-78) or n = cs. */
int mpopis_b200_set_sigma(mpopis_t *h, const double *Sigma, int64_t n);
/* pol.ws (K doubles) and the scalar CMA constants (POL:478-496). */
int mpopis_b200_set_cma(mpopis_t *h, const mpopis_cma_t *cma, const double *ws, int64_t n_ws);
/* Random.seed!(pol, seed) (src/MPOPIS.jl:54). The engine's generator is counter-based
 * Philox4x32-10 keyed by `seed`; Julia's MersenneTwister stream is NOT reproduced. */
int mpopis_b200_seed(mpopis_t *h, uint64_t seed);

/* Depth (iii): the whole functor (pol::AbstractGMPPI_Policy)(env) POL:221-238 /
 * (pol::MPPI_Policy)(env) POL:121-146, i.e. calculate_trajectory_costs + weighted noise +
 * get_controls_roll_U! (UTL:88-101). state: ss doubles (env.state); env_t: env.t;
 * U_inout: pol.U (cs), rolled in place; control_out: as doubles; its_run_out: AIS iterations executed. */
int mpopis_b200_plan(mpopis_t *h, const double *state, int64_t env_t, double *U_inout,
                     double *control_out, int32_t *its_run_out);
/* Same, with the standard-normal draws injected: Z is cs x K x N column-major (iteration n uses
 * Z[:,:,n] where the reference calls rand(pol.rng, P, K)); resample_u is K x (N-1) uniforms in
 * [0,1) for the :pmcmppi categorical draws (POL:805), NULL otherwise. Parity surface. */
int mpopis_b200_plan_with_noise(mpopis_t *h, const double *state, int64_t env_t, double *U_inout,
                                const double *Z, const double *resample_u, double *control_out,
                                int32_t *its_run_out);
/* The EnvpoolEnv seam. `rollout` replaces rollout_model(env::EnvpoolEnv, T, model_controls, pol) UTL:103-121: it
 * receives the clamped model controls of all K samples exactly as get_model_controls(action_space, Vₖ, T) UTL:42-53
 * shapes them — a K x as x T column-major array, controls[k + K*(r + as*t)] — steps its own batched simulator
 * from the current real state, writes traj_cost[k] = −Σ_t reward (UTL:110-111), restores the simulator
 * (reset!(env; restore=true), UTL:119) and returns 0 (non-zero aborts the plan with MPOPIS_ERR_BAD_ARG).
 * It is called on the calling thread, once per executed AIS iteration, between device phases. */
typedef int (*mpopis_rollout_fn)(void *user, const double *controls, int64_t K, int64_t as, int64_t T,
                                 double *traj_cost_out);
/* pol(env::EnvpoolEnv): simulate_model(pol, env::EnvpoolEnv, E, Σ_inv, U_orig) POL:240-259 (and the :mppi method
 * POL:148-184) inside the policy's calculate_trajectory_costs, then the weighted update and get_controls_roll_U!.
 * Sampling, the control cost γ U_origᵀ Σ⁻¹ (Vₖ − U_orig) (POL:248), adaptation, weights and the control run on
 * the device; only the K x cs controls (D2H) and the K costs (H2D) cross per iteration. Z / resample_u as in
 * plan_with_noise (NULL = the engine's Philox stream). Handle must have been created with MPOPIS_ENV_EXTERNAL. */
int mpopis_b200_plan_external(mpopis_t *h, double *U_inout, mpopis_rollout_fn rollout, void *user, const double *Z,
                              const double *resample_u, double *control_out, int32_t *its_run_out);

/* Results of the last plan(): trajectory_cost (K), weights (K), E (cs x K, shifted as in POL:468),
 * logger trajectories (K matrices T x ss, column-major, sample k at offset k*T*ss; needs
 * cfg.log_trajectories). Any pointer may be NULL. With world_size > 1 only this rank's shard
 * of E/trajectories is returned (columns rank*K/world .. ), costs and weights are global. */
int mpopis_b200_fetch(mpopis_t *h, double *costs, double *weights, double *E, double *traj);
/* Covariance / mean proposal state after the last plan() (debug + parity): Sigma' (cs x cs)
 * that the last executed iteration sampled from, and U + sum of mean shifts (cs). */
int mpopis_b200_fetch_proposal(mpopis_t *h, double *Sigma_last, double *U_last);

/* Depth (i): simulate_model(pol, env, E, Σ_inv, U_orig) POL:261-278 -> trajectory_cost (K).
 * U is pol.U at call time. Sigma_inv may be NULL when γ = λ(1-α) = 0. */
int mpopis_b200_rollout_costs(mpopis_t *h, const double *state, int64_t env_t, const double *U,
                              const double *U_orig, const double *E, const double *Sigma_inv,
                              double *costs_out);
/* compute_weights(Information_Theoretic(λ), costs) UTL:79-86. */
int mpopis_b200_weights(mpopis_t *h, const double *costs, int64_t K, double lambda, double *w_out);
/* Parity surface of the sort-free elite selection (csrc/select.cu): the elite SET order[1:m] of
 * `order = sortperm(trajectory_cost)` (POL:455-456; ascending sample ids, 0-based, restricted to the ownership window
 * [k0, k0 + kloc)) and the early-stop decision `maximum(abs.(diff(elite_traj_cost))) < 10e-3` (POL:458-461).
 * tau_out4 (nullable): {cost of the m-th smallest, its sample id, smallest cost, buckets used by the gap test}. */
int mpopis_b200_elite_select(mpopis_t *h, const double *costs, int64_t K, int64_t m, int64_t k0, int64_t kloc,
                             int32_t early_stop, int64_t *elite_ids_out, int64_t *n_out, int32_t *stop_out,
                             double *tau_out4);

/* within_track(track, pos) TRK:68-92 for n positions (pos = 2 x n column-major):
 * idx/idx2 are the 0-based min_idx / min_idx_2, dist = dist_to_pt, within = 0/1. */
int mpopis_b200_track_query(mpopis_t *h, const double *pos, int64_t n, int32_t *idx_out,
                            int32_t *idx2_out, double *dist_out, uint8_t *within_out);
/* One real environment step + reward: env(a); reward(env) (CAR:238-241,201-213; MCR:200-207,145-158;
 * EXM:4-22). state_inout ss doubles, action as doubles, env_t_inout env.t, done_out env.done
 * (MountainCar only; NULL allowed). */
int mpopis_b200_env_step(mpopis_t *h, double *state_inout, const double *action, int64_t *env_t_inout,
                         double *reward_out, uint8_t *done_out);
/* order = sortperm(costs) (POL:455, 563): stable ascending permutation, 0-based. Integer parity surface. */
int mpopis_b200_sortperm(mpopis_t *h, const double *costs, int64_t K, int64_t *perm_out);
/* reward(env) of the current state without stepping (CAR:201-213, MCR:145-158, EXM:10-22; `done` is
 * env.done, used by the MountainCar reward only). */
int mpopis_b200_env_reward(mpopis_t *h, const double *state, uint8_t done, double *reward_out);
/* The cs x K standard normals the engine's Philox generator produces for (control step `step`,
 * AIS iteration `iteration`) — RNG parity surface against oracle/ (same generator restated). */
int mpopis_b200_sample_normals(mpopis_t *h, int64_t step, int64_t iteration, double *Z_out);
/* cov(method, X) of CovarianceEstimation / StatsBase as used at POL:464 (+ no 1e-8 I):
 * X is p x n column-major (n observations of dimension p, i.e. `elite`), w NULL or n weights
 * (then the weighted mean_and_cov of POL:364,662,732; `corrected` selects the n-1 divisor of POL:807). */
int mpopis_b200_cov_estimate(mpopis_t *h, int32_t sigma_est, const double *X, int64_t p, int64_t n,
                             const double *w, int32_t corrected, double *mean_out, double *cov_out);
/* Lower Cholesky factor of an n x n SPD matrix (what MvNormal(Σ) holds). */
int mpopis_b200_cholesky(mpopis_t *h, const double *A, int64_t n, double *L_out);
/* Σ^-0.5 of a symmetric PD matrix (POL:580). */
int mpopis_b200_inv_sqrt(mpopis_t *h, const double *A, int64_t n, double *C_out);

/* λ̂ of the last shrinkage covariance estimate (:lw/:ss/:rblw/:oas), for parity tests. */
int mpopis_b200_last_shrinkage(mpopis_t *h, double *lambda_out);
/* Tuning knobs, not part of the reference API (the full table with defaults is in INTEGRATION.md):
 * "rollout_variant" (6 = automatic, the default: the warp-specialised kernel 5 while every rollout of the shard stays
 * resident, the thread-per-rollout kernel 3 beyond; 4 = warp-specialised at 96 registers; 0 = branchy fast formulation,
 * also the repair path; 1 = literal libm call sequence of CAR:299-333, the parity anchor), "rollout_spin" (hand-over of
 * the warp-specialised kernel: -1 automatic, 0 mbarrier, 1 spin counters), "rollout_block" (threads per CTA of the
 * thread-per-rollout kernel: 32, 64, 96 or 128), "rollout_profile" (1 = record per-warp cycles, see warp_cycles),
 * "rollout_stage" (noise tensor of kernel 3: 0 = register prefetch, 1 = TMA bulk copies into a shared-memory ring),
 * "graph" (1 = replay the control step as a CUDA graph, the default), "fuse_cov" (1 = shrinkage + ridge folded into
 * the Cholesky launch), "ce_small_fused" (1 = the whole :cemppi adaptation of K <= 512, cs <= 112 as one single-CTA launch),
 * "moments_small" (1 = single-CTA moment chain for <= 512 columns, the default; 0 = the multi-kernel chain).
 * get_option additionally reads "rollout_variant_used" (what 6 resolved to at the latest launch), "graph_active",
 * "ce_select" and "comm_peer". */
int mpopis_b200_set_option(mpopis_t *h, const char *key, double value);
int mpopis_b200_get_option(mpopis_t *h, const char *key, double *value_out);

/* Profiling aid: with set_option("rollout_profile", 1) every warp of the rollout kernel records the clock64() cycles
 * it spent in the kernel; this returns the values of the most recent rollout launch (ceil(K_local/32) entries at most
 * n are written) — the spread between warps shows how much of a launch is tail (repaired steps, full track scans). */
int mpopis_b200_warp_cycles(mpopis_t *h, int64_t *cycles_out, int64_t n);

/* Device-resident control loop used by bench.py's `value` leg: state and U stay in HBM, the
 * step is enqueued without host copies of inputs; the control is applied to the resident env
 * state (env(act), src/examples/car_example.jl:205-207) so that consecutive steps differ. */
int mpopis_b200_resident_reset(mpopis_t *h, const double *state, int64_t env_t, const double *U);
int mpopis_b200_resident_plan(mpopis_t *h, int32_t advance_env);
int mpopis_b200_resident_read(mpopis_t *h, double *state_out, double *U_out, double *control_out,
                              int32_t *its_run_out);
/* Σ reward(env) over the env steps of the resident loop since resident_reset (`rew += reward(env)`, car_example.jl:209;
 * mountaincar_example.jl:149) — lets many trials run as concurrent device-resident replicas without a per-step read. */
int mpopis_b200_resident_reward_sum(mpopis_t *h, double *sum_out);

/* AIS iterations executed since resident_reset() (summed over resident_plan() calls). */
int mpopis_b200_resident_total_its(mpopis_t *h, int64_t *total_its_out);
/* Measured FP64 FMA issue rate of the device (thread-level DFMA/s, dependent-chain micro-benchmark):
 * the roofline denominator of the FP64-bound rollout kernel, which MEASURED_PEAKS.json lacks. */
int mpopis_b200_measure_fp64_peak(mpopis_t *h, double *dfma_per_s_out);

/* Times the weighted-noise reduction kernel (POL:226-229: weights' * E[r,:] for every row r) alone over this
 * handle's E operand: CUDA-event ms per launch and the algorithmic bytes per launch (8·cs·K + 8·K). */
int mpopis_b200_bench_rowsum(mpopis_t *h, int32_t reps, double *ms_per_launch_out, double *bytes_per_launch_out);

/* Timing/introspection: kernels launched by this handle so far, device milliseconds spent in the
 * rollout kernel during the last plan (CUDA events on the handle's stream), the handle's
 * cudaStream_t (as void*). */
int64_t mpopis_b200_launch_count(mpopis_t *h);
int mpopis_b200_last_timing(mpopis_t *h, double *rollout_ms, double *total_ms, int32_t *rollout_launches);
void *mpopis_b200_stream(mpopis_t *h);

#ifdef __cplusplus
}
#endif
#endif /* MPOPIS_B200_H */
