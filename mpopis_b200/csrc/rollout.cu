// rollout.cu — G2: the K-trajectory rollout + running-cost kernel (the dominant kernel).
//
// Replaces simulate_model (POL:261-278) -> rollout_model (UTL:129-144) -> _step! (CAR:282-344) +
// reward (CAR:201-213 / MCR:145-158) + within_track (TRK:68-92), and the MountainCar step/reward
// (RLEnvs, EXM:4-22). One thread integrates one rollout: T control steps x nsub Euler sub-steps,
// all FP64. The kernel is FP64-pipe bound (≈2·10³ FP64 instructions per rollout-step against
// 16 B of HBM traffic), so the work here goes into removing transcendental calls
// without changing the mathematics:
//   * slip angles: tan(atan2(y,x) − δ) is evaluated as a rotated ratio (no atan2 / tan / atan);
//     the saturation test |α| < atan(3 fy_max / C) becomes |tan α| < 3 fy_max / C on the branch
//     where |α| < π/2, and the exact libm sequence is kept for Vx <= 0 (car reversing), where the
//     un-wrapped angle matters for sign(α);
//   * tyre-force constants (fx, fz, fy_max, cubic coefficients) depend only on (pedal, sign Vx):
//     hoisted out of the sub-step loop and recomputed only when sign(Vx) flips;
//   * heading wrap atan(sin Ψ, cos Ψ) is a conditional ±2π (identity on (−π, π]);
//   * the β test |atan(Vy, Vx)| > β_limit is Vx < cos(β_limit)·‖V‖ (the speed is needed anyway).
// The integer decisions of within_track (first arg-min, neighbour choice) are computed with
// explicitly non-contracted arithmetic (__dmul_rn/__dadd_rn) so that, on identical inputs, the
// indices are bit-identical to the reference's Float64 evaluation order.
#include <math_constants.h>

#include "engine.cuh"

namespace mpopis {

__device__ __forceinline__ double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
__device__ __forceinline__ double clamp1(double v) { return fmin(fmax(v, -1.0), 1.0); }

struct TrackView {
  const double *x, *y, *w;
  int n;
};

// within_track(track, pos) TRK:68-92. Integer-exact: distances use un-fused mul/add.
__device__ __forceinline__ bool within_track(const TrackView &tr, double px, double py, int *idx_out,
                                             int *idx2_out, double *dist_out) {
  int mi = 0;
  double best = CUDART_INF;
#pragma unroll 4
  for (int i = 0; i < tr.n; ++i) {  // TRK:71,73 (findmin -> FIRST minimum: strict <)
    double dx = tr.x[i] - px, dy = tr.y[i] - py;
    double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (d < best) best = d, mi = i;
  }
  int m1 = mi == 0 ? tr.n - 1 : mi - 1, p1 = mi == tr.n - 1 ? 0 : mi + 1;  // mod1, TRK:75-76
  double ax = tr.x[m1] - px, ay = tr.y[m1] - py, bx = tr.x[p1] - px, by = tr.y[p1] - py;
  double dm1 = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));  // TRK:77
  double dp1 = __dsqrt_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)));  // TRK:78
  int m2 = dm1 <= dp1 ? m1 : p1;                                              // TRK:79
  double p1x = tr.x[mi], p1y = tr.y[mi];
  double vx = tr.x[m2] - p1x, vy = tr.y[m2] - p1y, ux = px - p1x, uy = py - p1y;
  double t = (ux * vx + uy * vy) / (vx * vx + vy * vy);  // TRK:87 (projection on the infinite line)
  double ex = p1x + t * vx - px, ey = p1y + t * vy - py; // TRK:88
  double dist = sqrt(ex * ex + ey * ey);                 // TRK:89
  if (idx_out) *idx_out = mi;
  if (idx2_out) *idx2_out = m2;
  *dist_out = dist;
  return dist < tr.w[mi];  // TRK:90
}

// Tyre-force constants that depend only on (pedal, sign(Vx)); CAR:310-318 + the invariant part
// of calc_tire_fy CAR:252-260.
struct TireConsts {
  double fxf, fxr;           // longitudinal split, CAR:315-316
  double fymax_f, fymax_r;   // sqrt(max((μ fz)² − fx², 1e-8)), CAR:253
  double thr_f, thr_r;       // 3 fy_max / C  (tan of the sliding angle), CAR:255
  double c2_f, c2_r;         // C² / (3 fy_max)
  double c3_f, c3_r;         // C³ / (27 fy_max²)
};

__device__ __forceinline__ TireConsts tire_consts(const CarParams &P, double accel, double bk,
                                                  double split, double sgnVx) {
  TireConsts c;
  double fx = accel + bk * sgnVx;  // CAR:310-312
  c.fxf = split * fx;
  c.fxr = (1 - split) * fx;
  double L = P.l_r + P.l_f;
  double fzf = (P.m * P.l_r * 9.81 - P.h_cm * fx) / L;  // calc_tire_fz 'f', CAR:262-272
  double fzr = (P.m * P.l_f * 9.81 + P.h_cm * fx) / L;  // calc_tire_fz 'r'
  c.fymax_f = sqrt(fmax((P.mu_f * fzf) * (P.mu_f * fzf) - c.fxf * c.fxf, 1e-8));
  c.fymax_r = sqrt(fmax((P.mu_r * fzr) * (P.mu_r * fzr) - c.fxr * c.fxr, 1e-8));
  c.thr_f = 3 * c.fymax_f / P.C_af;
  c.thr_r = 3 * c.fymax_r / P.C_ar;
  c.c2_f = (P.C_af * P.C_af) / (3 * c.fymax_f);
  c.c2_r = (P.C_ar * P.C_ar) / (3 * c.fymax_r);
  c.c3_f = (P.C_af * P.C_af * P.C_af) / (27 * (c.fymax_f * c.fymax_f));
  c.c3_r = (P.C_ar * P.C_ar * P.C_ar) / (27 * (c.fymax_r * c.fymax_r));
  return c;
}

// brush-tyre lateral force from tan α = num/den with |α| < π (fast path, Vx > 0)
__device__ __forceinline__ double tire_fy_ratio(double num, double den, double C, double c2, double c3,
                                                double thr, double fymax) {
  double ta = num / den;
  double cubic = -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  double sat = -fymax * jl_sign(num);
  return (den > 0.0 && fabs(ta) < thr) ? cubic : sat;
}

// literal calc_tire_fy, CAR:252-260
__device__ __forceinline__ double tire_fy_literal(double alpha, double C, double c2, double c3,
                                                  double thr, double fymax) {
  double ta = tan(alpha);
  if (fabs(alpha) < atan(thr)) return -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  return -fymax * jl_sign(alpha);
}

// _step!(env::CarRacingEnv, a), CAR:282-344. s = [x, y, Ψ, Vx, Vy, Ψ̇, δ, pedal].
template <bool FAST>
__device__ __forceinline__ void car_step(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                         double a0, double a1) {
  double x = s[0], y = s[1], psi = s[2], Vx = s[3], Vy = s[4], psid = s[5], delta = s[6];
  const double tgt = a0 * P.d_max - delta;
  const double rate = fmin(fabs(tgt) / dt, P.dd_max) * jl_sign(tgt);  // CAR:295-296
  const double pedal = a1;                                           // CAR:297
  const double accel = P.Fx_max * fmax(pedal, 0.0);                  // CAR:310
  const double bk = P.Fx_min * fmin(pedal, 0.0);                     // CAR:311 without sign(Vx)
  const double split = pedal <= 0.0 ? P.l_brake : P.l_drive;
  const double inv_Izz = 1 / P.Izz, inv_m = 1 / P.m;                 // CAR:322-324 multiply by (1/·)
  double sg = jl_sign(Vx);
  TireConsts tc = tire_consts(P, accel, bk, split, sg);
  for (int i = 0; i < nsub; ++i) {
    delta += rate * ddt;  // CAR:301
    double sg_now = jl_sign(Vx);
    if (sg_now != sg) {  // sign(Vx) flipped: brake force changes direction (rare)
      sg = sg_now;
      if (bk != 0.0) tc = tire_consts(P, accel, bk, split, sg);
    }
    double sd, cd;
    sincos(delta, &sd, &cd);
    const double yf = Vy + P.l_f * psid, yr = Vy - P.l_r * psid;
    double fyf, fyr;
    if (FAST && Vx > 0.0) {
      // tan(atan(yf, Vx) − δ) = (yf cδ − Vx sδ)/(Vx cδ + yf sδ); rear: tan α_r = yr / Vx
      fyf = tire_fy_ratio(yf * cd - Vx * sd, Vx * cd + yf * sd, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f,
                          tc.fymax_f);
      fyr = tire_fy_ratio(yr, Vx, P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
    } else {
      double a_f = atan2(yf, Vx) - delta;  // CAR:304
      double a_r = atan2(yr, Vx);          // CAR:305
      fyf = tire_fy_literal(a_f, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f, tc.fymax_f);
      fyr = tire_fy_literal(a_r, P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
    }
    const double fx_aero = (P.C_D0 + P.C_D1 * fabs(Vx)) * sg;  // CAR:308
    const double psidd = inv_Izz * (P.l_f * (tc.fxf * sd + fyf * cd) - P.l_r * fyr);        // CAR:322
    const double Vy_dot = inv_m * (fyf * cd + tc.fxf * sd + fyr) - psid * Vx;               // CAR:323
    const double Vx_dot = inv_m * (tc.fxf * cd - fyf * sd + tc.fxr - fx_aero) + psid * Vy;  // CAR:324
    psid += psidd * ddt;  // CAR:326
    Vx += Vx_dot * ddt;   // CAR:327
    Vy += Vy_dot * ddt;   // CAR:328
    psi += psid * ddt;    // CAR:329
    double sp, cp;
    if (FAST) {  // CAR:330: atan(sin Ψ, cos Ψ) == Ψ on (−π, π], otherwise Ψ − 2π·round(Ψ/2π)
      if (fabs(psi) > CUDART_PI) {
        double k = rint(psi * 0.15915494309189535);
        psi = fma(-k, 6.283185307179586, psi);
        psi = fma(-k, 2.4492935982947064e-16, psi);
      }
      sincos(psi, &sp, &cp);
    } else {
      sincos(psi, &sp, &cp);
      psi = atan2(sp, cp);
      sincos(psi, &sp, &cp);
    }
    x += (Vx * cp - Vy * sp) * ddt;  // CAR:331
    y += (Vx * sp + Vy * cp) * ddt;  // CAR:332
  }
  s[0] = x, s[1] = y, s[2] = psi, s[3] = Vx, s[4] = Vy, s[5] = psid, s[6] = delta, s[7] = pedal;
}

// reward(env::CarRacingEnv), CAR:201-213
template <bool FAST>
__device__ __forceinline__ double car_reward(const CarParams &P, double cos_bl, const TrackView &tr,
                                             const double *s) {
  double dist;
  bool within = within_track(tr, s[0], s[1], nullptr, nullptr, &dist);
  double speed = sqrt(s[3] * s[3] + s[4] * s[4]);
  bool exceed = FAST ? (s[3] < cos_bl * speed) : (fabs(atan2(s[4], s[3])) > P.b_limit);  // CAR:181-189
  double rew = 0.0;
  if (!within) rew += -1000000.0;
  if (exceed) rew += -5000.0;
  rew += -dist;
  rew += 2.0 * speed;
  return rew;
}

// (env)(a) + reward(env) for 1..N cars: CAR:238-241 / MCR:200-207, MCR:145-158
template <int NCARS, bool FAST>
__device__ __forceinline__ double cars_step_reward(const CarEnvArgs &env, const TrackView &tr, double *s,
                                                   const double *a) {
#pragma unroll
  for (int c = 0; c < NCARS; ++c)
    car_step<FAST>(env.car[c], env.dt, env.ddt, env.nsub, s + 8 * c, a[2 * c], a[2 * c + 1]);
  double rew = 0.0;
#pragma unroll
  for (int c = 0; c < NCARS; ++c) {
    rew += car_reward<FAST>(env.car[c], env.cos_blimit[c], tr, s + 8 * c);
#pragma unroll
    for (int j = c + 1; j < NCARS; ++j) {
      double dx = s[8 * j] - s[8 * c], dy = s[8 * j + 1] - s[8 * c + 1];
      double dd = sqrt(dx * dx + dy * dy);
      rew += -dd;
      if (dd <= 4.0) rew += -11000.0;  // MCR:153-155 (docstring says −7000; code is −11000)
    }
  }
  return rew;
}

__device__ __forceinline__ TrackView stage_track(const CarEnvArgs &env, double *smem) {
  // track′ (x′, y′, lane_width′) is read by every rollout at every step: stage it in shared memory
  for (int i = threadIdx.x; i < 3 * env.n_trk; i += blockDim.x) smem[i] = env.trk[i];
  __syncthreads();
  TrackView tr{smem, smem + env.n_trk, smem + 2 * env.n_trk, env.n_trk};
  return tr;
}

template <int NCARS, bool FAST>
__global__ void __launch_bounds__(128) rollout_car_kernel(const __grid_constant__ CarEnvArgs env,
                                                          const __grid_constant__ RolloutArgs a,
                                                          const int *stop) {
  extern __shared__ double smem[];
  if (stop && *stop) return;
  TrackView tr = stage_track(env, smem);
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.K) return;
  constexpr int AS = 2 * NCARS, SS = 8 * NCARS;
  double s[SS];
#pragma unroll
  for (int q = 0; q < SS; ++q) s[q] = __ldg(a.state0 + q);
  const double *Ek = a.E + k;
  double cost = 0.0, cc = 0.0;
  for (int t = 0; t < a.T; ++t) {
    double act[AS];
#pragma unroll
    for (int r = 0; r < AS; ++r) {
      const int row = t * AS + r;
      double v = __ldg(a.U + row) + Ek[(size_t)row * a.ldk];  // Vₖ = pol.U + E[:,k], POL:271
      if (a.bvec) cc += __ldg(a.bvec + row) * (v - __ldg(a.U_orig + row));  // POL:272
      act[r] = clamp1(v);                                                   // UTL:55-67
    }
    cost -= cars_step_reward<NCARS, FAST>(env, tr, s, act);  // UTL:137-138
    if (a.traj) {
#pragma unroll
      for (int q = 0; q < SS; ++q) a.traj[((size_t)k * SS + q) * a.T + t] = s[q];  // UTL:139-141
    }
  }
  a.costs[k] = cost + cc;  // POL:274-275
}

// RLEnvs MountainCarEnv(continuous=true) step + EXM:10-22 reward
__device__ __forceinline__ double mc_step_reward(const McEnvArgs &e, double &x, double &v, long long &t,
                                                 double act, bool *done_out) {
  t += 1;
  v += act * e.power + cos(3 * x) * (-e.gravity);
  v = fmin(fmax(v, -e.max_speed), e.max_speed);
  x += v;
  x = fmin(fmax(x, e.min_pos), e.max_pos);
  if (x == e.min_pos && v < 0) v = 0;
  bool done = (x >= e.goal_pos && v >= e.goal_vel) || t >= e.max_steps;
  double rew = 0.0;
  if (x >= e.goal_pos && v >= e.goal_vel) rew += 100000;
  rew += fabs(v);
  rew += done ? 0.0 : -1.0;
  if (done_out) *done_out = done;
  return rew;
}

__global__ void __launch_bounds__(128) rollout_mc_kernel(const __grid_constant__ McEnvArgs env,
                                                         const __grid_constant__ RolloutArgs a,
                                                         const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.K) return;
  double x = __ldg(a.state0), v = __ldg(a.state0 + 1);
  long long t_env = *a.env_t;
  const double *Ek = a.E + k;
  double cost = 0.0, cc = 0.0;
  for (int t = 0; t < a.T; ++t) {
    double val = __ldg(a.U + t) + Ek[(size_t)t * a.ldk];
    if (a.bvec) cc += __ldg(a.bvec + t) * (val - __ldg(a.U_orig + t));
    cost -= mc_step_reward(env, x, v, t_env, clamp1(val), nullptr);
    if (a.traj) {
      a.traj[((size_t)k * 2 + 0) * a.T + t] = x;
      a.traj[((size_t)k * 2 + 1) * a.T + t] = v;
    }
  }
  a.costs[k] = cost + cc;
}

template <bool FAST>
static void launch_rollout_car_v(const CarEnvArgs &env, const RolloutArgs &a, int block, const int *stop,
                                 cudaStream_t st) {
  const int grid = (a.K + block - 1) / block;
  const size_t smem = sizeof(double) * 3 * env.n_trk;
#define MPOPIS_LAUNCH(N)                                                                   \
  case N:                                                                                  \
    if (smem > 48 * 1024)                                                                  \
      cudaFuncSetAttribute(rollout_car_kernel<N, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)smem);                                                     \
    rollout_car_kernel<N, FAST><<<grid, block, smem, st>>>(env, a, stop);                       \
    break;
  switch (env.n_cars) {
    MPOPIS_LAUNCH(1)
    MPOPIS_LAUNCH(2)
    MPOPIS_LAUNCH(3)
    MPOPIS_LAUNCH(4)
    MPOPIS_LAUNCH(5)
    MPOPIS_LAUNCH(6)
    MPOPIS_LAUNCH(7)
    MPOPIS_LAUNCH(8)
  }
#undef MPOPIS_LAUNCH
}

void launch_rollout_car(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, const int *stop,
                        cudaStream_t st) {
  if (variant == 0) launch_rollout_car_v<true>(env, a, block, stop, st);
  else launch_rollout_car_v<false>(env, a, block, stop, st);
}

void launch_rollout_mc(const McEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t st) {
  rollout_mc_kernel<<<(a.K + block - 1) / block, block, 0, st>>>(env, a, stop);
}

// ---- parity surfaces -------------------------------------------------------------------------
__global__ void track_query_kernel(const __grid_constant__ CarEnvArgs env, const double *pos, int n, int *idx,
                                   int *idx2, double *dist, unsigned char *within) {
  extern __shared__ double smem[];
  TrackView tr = stage_track(env, smem);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a, b;
  double d;
  bool w = within_track(tr, pos[2 * i], pos[2 * i + 1], &a, &b, &d);
  if (idx) idx[i] = a;
  if (idx2) idx2[i] = b;
  if (dist) dist[i] = d;
  if (within) within[i] = w ? 1 : 0;
}

void launch_track_query(const CarEnvArgs &env, const double *pos, int n, int *idx, int *idx2,
                        double *dist, unsigned char *within, cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * env.n_trk;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(track_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  track_query_kernel<<<(n + 127) / 128, 128, smem, st>>>(env, pos, n, idx, idx2, dist, within);
}

template <bool FAST>
__global__ void env_step_car_kernel(const __grid_constant__ CarEnvArgs env, double *state,
                                    const double *action, long long *env_t, double *reward) {
  extern __shared__ double smem[];
  TrackView tr = stage_track(env, smem);
  if (threadIdx.x != 0) return;
  double rew = 0.0;
  for (int c = 0; c < env.n_cars; ++c)
    car_step<FAST>(env.car[c], env.dt, env.ddt, env.nsub, state + 8 * c, action[2 * c], action[2 * c + 1]);
  for (int c = 0; c < env.n_cars; ++c) {
    rew += car_reward<FAST>(env.car[c], env.cos_blimit[c], tr, state + 8 * c);
    for (int j = c + 1; j < env.n_cars; ++j) {
      double dx = state[8 * j] - state[8 * c], dy = state[8 * j + 1] - state[8 * c + 1];
      double dd = sqrt(dx * dx + dy * dy);
      rew += -dd;
      if (dd <= 4.0) rew += -11000.0;
    }
  }
  *env_t += 1;  // CAR:283
  if (reward) *reward = rew;
}

void launch_env_step_car(const CarEnvArgs &env, double *state, const double *action, long long *env_t,
                         double *reward, int variant, cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * env.n_trk;
  if (variant == 0) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(env_step_car_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    env_step_car_kernel<true><<<1, 32, smem, st>>>(env, state, action, env_t, reward);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(env_step_car_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    env_step_car_kernel<false><<<1, 32, smem, st>>>(env, state, action, env_t, reward);
  }
}

// reward(env) without stepping (CAR:201-213, MCR:145-158, EXM:10-22)
template <bool FAST>
__global__ void env_reward_car_kernel(const __grid_constant__ CarEnvArgs env, const double *state, double *reward) {
  extern __shared__ double smem[];
  TrackView tr = stage_track(env, smem);
  if (threadIdx.x != 0) return;
  double rew = 0.0;
  for (int c = 0; c < env.n_cars; ++c) {
    rew += car_reward<FAST>(env.car[c], env.cos_blimit[c], tr, state + 8 * c);
    for (int j = c + 1; j < env.n_cars; ++j) {
      double dx = state[8 * j] - state[8 * c], dy = state[8 * j + 1] - state[8 * c + 1];
      double dd = sqrt(dx * dx + dy * dy);
      rew += -dd;
      if (dd <= 4.0) rew += -11000.0;
    }
  }
  *reward = rew;
}

void launch_env_reward_car(const CarEnvArgs &env, const double *state, double *reward, int variant,
                           cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * env.n_trk;
  if (variant == 0) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(env_reward_car_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    env_reward_car_kernel<true><<<1, 32, smem, st>>>(env, state, reward);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(env_reward_car_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    env_reward_car_kernel<false><<<1, 32, smem, st>>>(env, state, reward);
  }
}

__global__ void env_reward_mc_kernel(const __grid_constant__ McEnvArgs e, const double *state, int done,
                                     double *reward) {
  if (threadIdx.x != 0) return;
  double rew = 0.0;
  if (state[0] >= e.goal_pos && state[1] >= e.goal_vel) rew += 100000;
  rew += fabs(state[1]);
  rew += done ? 0.0 : -1.0;
  *reward = rew;
}

void launch_env_reward_mc(const McEnvArgs &env, const double *state, int done, double *reward, cudaStream_t st) {
  env_reward_mc_kernel<<<1, 32, 0, st>>>(env, state, done, reward);
}

__global__ void env_step_mc_kernel(const __grid_constant__ McEnvArgs env, double *state, const double *action,
                                   long long *env_t, double *reward, unsigned char *done) {
  if (threadIdx.x != 0) return;
  double x = state[0], v = state[1];
  long long t = *env_t;
  bool d;
  double rew = mc_step_reward(env, x, v, t, action[0], &d);
  state[0] = x, state[1] = v, *env_t = t;
  if (reward) *reward = rew;
  if (done) *done = d ? 1 : 0;
}

void launch_env_step_mc(const McEnvArgs &env, double *state, const double *action, long long *env_t,
                        double *reward, unsigned char *done, cudaStream_t st) {
  env_step_mc_kernel<<<1, 32, 0, st>>>(env, state, action, env_t, reward, done);
}

}  // namespace mpopis
