// engine.cuh — shared declarations of the sm_100a MPPI/MPOPI sampling engine.
//
// Device data layout: the noise tensor E lives in HBM as [cs][ldk] doubles with the
// SAMPLE index fastest (ldk = K_local rounded up to 32), i.e. the transpose of the reference's
// cs x K column-major Julia matrix (POL:271). Thread-per-rollout kernels then read E[r][k] with
// fully coalesced 8-byte loads across a warp, and every reduction over samples (weighted sums,
// moments) streams contiguous rows. The C-ABI converts to/from the Julia layout at the boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpopis_b200.h"

namespace mpopis {

struct CarParams {  // CAR:2-21 in declaration order
  double m, Izz, h_cm, l_f, l_r, C_D0, C_D1, C_af, C_ar, mu_f, mu_r, d_max, dd_max, Fx_max, Fx_min,
      l_brake, l_drive, b_limit;
};

// Loop-invariant combinations of the car constants, evaluated once on the host (set_car_env) so that the rollout
// kernel reads them as constant-bank operands instead of holding ~30 registers of hoisted values (rollout.cu v4).
struct CarDerived {
  double cI1, cI2;   // δt·l_f/Izz, δt·l_r/Izz                       (CAR:322,326)
  double cm;         // δt/m                                         (CAR:323-324,327-328)
  double kx;         // 1 − (δt/m)·C_D1: linear part of the aero drag folded into the Vx update (CAR:308)
  double cmCD0;      // (δt/m)·C_D0
  double inv_L;      // 1/(l_f + l_r)                                (CAR:262-272)
  double wf, wr;     // m·l_r·9.81, m·l_f·9.81
  double thrC_f, thrC_r;  // 3/C_αf, 3/C_αr                          (CAR:255)
  double c2C_f, c2C_r;    // C_α²/3
  double c3C_f, c3C_r;    // C_α³/27
};

inline __host__ __device__ CarDerived derive_car(const CarParams &P, double ddt) {
  CarDerived D;
  const double cI = ddt * (1 / P.Izz);
  D.cI1 = cI * P.l_f, D.cI2 = cI * P.l_r, D.cm = ddt * (1 / P.m);
  D.kx = 1.0 - D.cm * P.C_D1, D.cmCD0 = D.cm * P.C_D0;
  D.inv_L = 1.0 / (P.l_r + P.l_f);
  D.wf = P.m * P.l_r * 9.81, D.wr = P.m * P.l_f * 9.81;
  D.thrC_f = 3.0 / P.C_af, D.thrC_r = 3.0 / P.C_ar;
  D.c2C_f = P.C_af * P.C_af / 3.0, D.c2C_r = P.C_ar * P.C_ar / 3.0;
  D.c3C_f = P.C_af * P.C_af * P.C_af / 27.0, D.c3C_r = P.C_ar * P.C_ar * P.C_ar / 27.0;
  return D;
}

struct CarEnvArgs {
  CarParams car[MPOPIS_MAX_CARS];
  CarDerived der[MPOPIS_MAX_CARS];
  double cos_blimit[MPOPIS_MAX_CARS];  // cos(β_limit) for the atan-free β test (fast path)
  double dt, ddt;
  int nsub;  // round(Int, dt/δt), CAR:299
  int n_cars;
  int n_trk;
  const double *trk;  // device: x[n_trk], y[n_trk], w[n_trk]
  // exact nearest-point pruning table (rollout.cu within_track<true>): lut_nx x lut_ny cells of
  // 1/lut_inv_c metres starting at (lut_x0, lut_y0); nullptr disables it
  const uint4 *lut;
  double lut_x0, lut_y0, lut_inv_c;
  int lut_nx, lut_ny;
};

struct McEnvArgs {  // RLEnvs MountainCarEnv params
  double min_pos, max_pos, max_speed, goal_pos, goal_vel, power, gravity;
  long long max_steps;
};

struct RolloutArgs {
  const double *E;      // [cs][ldk]
  long long ldk;        // row pitch of E in doubles
  const double *U;      // [cs] current proposal mean (pol.U inside the AIS loop)
  const double *U_orig; // [cs]
  const double *bvec;   // [cs] (γ U_orig') Σ_inv, or nullptr when γ = 0 (POL:272)
  const double *state0; // [ss] env.state
  const long long *env_t; // device scalar env.t (MountainCar max_steps term)
  double *costs;        // [K_local]
  double *traj;         // nullptr or [K_local][ss][T] (logger, UTL:139-141)
  int K;                // samples of this shard
  long long *warp_cycles; // nullptr, or one slot per warp: clock64() spent in the kernel ("rollout_profile" option)
  int T;                // horizon
};

#ifdef __CUDACC__
// Order-preserving 64-bit image of a cost under Base.isless (the `lt` of sortperm, POL:455,563): negative doubles
// reversed, −0.0 < +0.0, every NaN (either sign bit) above +Inf. Ties are broken by the sample index, which makes the
// composite (key, index) unique — any comparison sort or selection on it reproduces the stable order.
constexpr unsigned long long COST_KEY_NAN = 0xFFFFFFFFFFFFFFFEULL;  // below the padding key ~0
__device__ __forceinline__ unsigned long long cost_key(double c) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(c);
  if (c != c) return COST_KEY_NAN;
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double key_cost(unsigned long long k) {
  if (k >= COST_KEY_NAN) return __longlong_as_double(0x7ff8000000000000LL);
  k = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)k);
}
#endif

// ---- launchers (host). Every kernel of the AIS loop takes the device-side `stop` flag ----------
// rollout.cu  (variant 0 = fast math-equivalent formulation, 1 = literal libm call sequence)
void launch_rollout_car(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, int stage,
                        const int *stop, cudaStream_t s);
// rollout_split.cu ("v5", rollout_variant 4): velocity warps + pose/reward warps; returns 0 when not applicable
int rollout_split_max_cars();
int rollout_split_capacity(int n_cars, int num_sms);  // rollouts one launch keeps resident (one wave)
int launch_rollout_car_split(const CarEnvArgs &env, const RolloutArgs &a, int wide, int spin, const int *stop,
                             cudaStream_t s);
void launch_rollout_mc(const McEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t s);
void launch_track_query(const CarEnvArgs &env, const double *pos, int n, int *idx, int *idx2, double *dist,
                        unsigned char *within, int use_lut, cudaStream_t s);
void launch_env_step_car(const CarEnvArgs &env, double *state, const double *action, long long *env_t,
                         double *reward, int variant, cudaStream_t s);
void launch_env_step_mc(const McEnvArgs &env, double *state, const double *action, long long *env_t,
                        double *reward, unsigned char *done, cudaStream_t s);
void launch_env_reward_car(const CarEnvArgs &env, const double *state, double *reward, int variant, cudaStream_t s);
void launch_env_reward_mc(const McEnvArgs &env, const double *state, int done, double *reward, cudaStream_t s);
// sampling.cu
// step_dev (nullable) overrides `step`: the control-step counter in device memory (graph replay)
void launch_philox_normals(double *Z, long long ldk, int cs, int K, long long k0, uint64_t seed, uint32_t step,
                           const unsigned *step_dev, uint32_t iter, const int *stop, cudaStream_t s);
void launch_philox_uniforms(double *u, int K, uint64_t seed, uint32_t step, const unsigned *step_dev, uint32_t iter,
                            const int *stop, cudaStream_t s);
void launch_apply_L(const double *Lt, int cs, int bs, const double *Z, double *E, long long ldk, int K,
                    const int *stop, cudaStream_t s);
void launch_transpose_in(const double *colmajor, double *dev, int cs, int K, long long ldk, cudaStream_t s);
void launch_transpose_out(const double *dev, double *colmajor, int cs, int K, long long ldk, const double *shift_a,
                          const double *shift_b, cudaStream_t s);
// stats.cu
// returns the number of kernels launched; scratch: 512 doubles (nullptr forces the single-CTA kernel)
int launch_weights(const double *costs, int K, double lambda, double *w, double *scratch, const int *stop,
                   cudaStream_t s);
int rowsum_nchunks(int n);
// n_dev (nullable): device-side column count, n_eff = min(n, *n_dev); grids are sized for n
void launch_rowsum_partial(const double *X, long long ld, int rows, int n, const double *w, double *partial,
                           const int *stop, cudaStream_t s, const int *n_dev = nullptr, int sq = 0);
void launch_reduce_partials(const double *partial, int nchunks, int n, double *out, const int *stop, cudaStream_t s,
                            int stride = 0);
void launch_finalize_mean(const double *sums, int rows, double *mu, double *U, const double *scale_dev,
                          const int *stop, cudaStream_t s);
int syrk_nchunks(int n);
void launch_syrk_partial(const double *X, long long ld, int p, int n, const double *w, const double *mu, double *P,
                         const int *stop, cudaStream_t s, const int *n_dev = nullptr);
void launch_scatter_reduce(const double *P, int nchunks, int p, double *S, const int *stop, cudaStream_t s);
int shrink_q_nblocks(int n);
void launch_shrink_q_partial(const double *X, long long ld, int p, int n, const double *w, const double *mu,
                             const double *Sraw, const double *cnt_dev, int standardise, double *partial,
                             const int *stop, cudaStream_t s, const int *n_dev = nullptr,
                             const double *dinv_ext = nullptr);
void launch_cov_finalize(const double *Sraw, int p, const double *cnt_dev, int corrected, int method,
                         const double *qpart, int nq, double ridge, double *Sigma, double *lambda_out,
                         const int *stop, cudaStream_t s);
// small_adapt.cu: sort + early stop + elite moments + shrinkage + Cholesky of one :cemppi iteration in ONE single-CTA launch
// (K <= 512, m <= 128, cs <= 112: the reference's own sizes). Returns 0 when the sizes are not covered.
int launch_ce_small_adapt(const double *costs, int K, int m, int early_stop, const double *E, long long ldk, int n,
                          int method, double ridge, unsigned long long *keys_out, int *order_out, double *mu_out,
                          double *U_cur, double *sums_out, double *Sigma, double *Lt, double *lambda_out, int *info, int tag,
                          int *stop_flag, cudaStream_t s);
// single-CTA fusion of the whole moment chain for n <= MOMENTS_SMALL_MAX columns (single GPU)
constexpr int MOMENTS_SMALL_MAX = 512;
void launch_moments_small(const double *X, long long ld, int p, int n, const double *w, const int *cols, int want_cov,
                          int corrected,
                          int method, double ridge, double *mu_out, double *U, const double *scale_dev,
                          double *sums_out, double *Sraw, double *Sigma, double *lambda_out, const int *stop,
                          cudaStream_t s);
void launch_gather_cols(const double *E, long long ldk, int cs, const int *order, int m, long long k0, int Kloc,
                        double *X, long long ldx, double *mask, const int *stop, cudaStream_t s,
                        const int *m_dev = nullptr);
void launch_elite_stop(const double *sorted_costs, int m, int enabled, int *stop, cudaStream_t s);
void launch_iter_begin(const int *stop, int *its, int *total_its, cudaStream_t s);
void launch_pmc_counts(const double *wglobal, int K, const double *u, double *cdf, int *counts, long long k0,
                       int Kloc, double *wloc, const int *stop, cudaStream_t s);
void launch_ctrl_vec(const double *Sinv, int cs, const double *U_orig, double gamma, double *b, cudaStream_t s);
// step_dev (nullable): incremented by one (the Philox control-step counter)
void launch_finalize_control(const double *wsum, const double *U_orig, const double *U_cur, int cs, int as, int T,
                             double *U_next, double *control, const double *bounds, unsigned *step_dev, cudaStream_t s);
void launch_set_scalar(double *dst, double value, cudaStream_t s);
// linalg.cu
void launch_chol(const double *A, int n, const double *sigma_dev, double *Lt, double *Wglobal, int *info, int tag,
                 const int *stop, cudaStream_t s);
// cov_finalize + chol fused (n <= 160); q_dev: the shrinkage statistic (:lw / :ss). Returns 0 when not applicable.
int launch_chol_cov(const double *Sraw, int n, const double *cnt_dev, int corrected, int method, const double *q_dev,
                    double ridge, double *Sigma, double *Lt, double *lambda_out, int *info, int tag, const int *stop,
                    cudaStream_t s);
void launch_chol_solve(const double *Lt, int n, const double *u, double gamma, double *b, const int *stop,
                       cudaStream_t s);
void launch_transpose_sq(const double *in, double *out, int n, cudaStream_t s);
// cma.cu
long long launch_dfma_peak(double *out, int grid, int iters, cudaStream_t s);
int inv_sqrt_max_ctas(int num_sms);
int launch_inv_sqrt(const double *A, int n, double *Cout, double *ws, int *info, int tag, const int *stop,
                    int max_ctas, cudaStream_t s);  // returns a cudaError_t
void launch_cma_lin_gather(const double *X, long long ldx, int cs, const int *order, int K, double *dvec,
                           const int *stop, cudaStream_t s);
void launch_cma_vec(const double *dw, const double *C, const double *dvec, const double *ws, int K, int cs,
                    int n_iter, const mpopis_cma_t &c, double *psig, double *pSig, double *sigma_dev, double *U,
                    double *Sigma, const int *stop, cudaStream_t s);
// select.cu — :cemppi elite selection without a sort (radix select + bucketed early-stop test + compaction)
constexpr int SELECT_MAX_CTAS = 592;
constexpr size_t SELECT_WS_BYTES = 64 * 1024;
int select_max_ctas(int num_sms);
long long select_bucket_capacity(int m);  // entries of bmin / bmax
void launch_select_init(void *ws, unsigned long long *bmin, unsigned long long *bmax, long long nb_cap, cudaStream_t s);
// costs: the GLOBAL cost vector (Ktot); [k0, k0 + Kloc) = the samples this shard owns. Outputs: eidx[0:*m_loc] = local
// ids of the owned elites in index order, *stop_flag raised when the reference would `break` (POL:459-461);
// tau_out (nullable, 4 doubles): {τ cost, τ index, smallest cost, buckets used}. Returns a cudaError_t.
int launch_ce_select(const double *costs, int Ktot, int m, long long k0, int Kloc, int early_stop, void *ws,
                     unsigned long long *bmin, unsigned long long *bmax, long long nb_cap, int *eidx, int *m_loc,
                     double *tau_out, int *stop_flag, const int *stop, int max_ctas, cudaStream_t s);
int elite_gather_nchunks(int m_max);
// X[r][j] = E[r][eidx[j]], partial[(2c + {0,1}) * cs + r] = Σ x, Σ x² over chunk c
void launch_elite_gather_sums(const double *E, long long ldk, int cs, const int *eidx, const int *m_loc, int m_max,
                              double *X, long long ldx, double *partial, const int *stop, cudaStream_t s);
// sums = [Σx | n | Σx²]; finalize: μ, U += μ, dinv = 1/σ (standardise) or 1. nchunks = 0: sums already reduced.
void launch_ce_sums(const double *partial, int nchunks, int cs, const int *m_loc, double *sums, int finalize,
                    int standardise, double *mu, double *U, double *dinv, const int *stop, cudaStream_t s);
// sort.cu
int sort_launches(int K);
int sort_max_ctas(int num_sms);
// returns a cudaError_t (cooperative launch); stop_flag != nullptr: also run the elite early-stop test
int launch_sortperm(const double *costs, int K, unsigned long long *keys_a, unsigned long long *keys_b, int *order,
                    int *vals_b, int m, int early_stop, int *stop_flag, const int *stop, int max_ctas,
                    cudaStream_t s);

}  // namespace mpopis
