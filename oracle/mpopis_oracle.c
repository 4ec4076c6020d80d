/*
 * mpopis_oracle.c — CPU oracle: a literal C restatement of the MPPI/MPOPI hot path of
 * sisl/MPOPIS. TEST INFRASTRUCTURE ONLY (see mpopis_oracle.h). PARITY UNPINNED (no Julia here,
 * no reference tests/golden vectors exist).
 *
 * Reference files (relative to the reference root):
 *   POL = src/mppi_mpopi_policies.jl   UTL = src/utils.jl   CAR = src/envs/car_racing.jl
 *   TRK = src/envs/car_racing_tracks/car_racing_tracks.jl   MCR = src/envs/multi-car_racing.jl
 *   EXM = src/examples/mountaincar_example.jl
 * Third-party semantics (Distributions 0.25, StatsBase 0.34, CovarianceEstimation 0.2,
 * ReinforcementLearning 0.11 — none vendored in the reference) are restated from their
 * published definitions.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp (no -ffast-math): every expression is
 * evaluated in the order Julia evaluates it, without FMA contraction.
 */
#define _GNU_SOURCE
#include "mpopis_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[512] = "";
/* threads for the embarrassingly parallel helpers (normal draws, L·Z): the reference gets these from
 * multithreaded BLAS / a fast native RNG, so the CPU baseline should not be charged for them serially */
static int g_threads = 1;
const char *orc_last_error(void) { return g_err; }
#define FAIL(code, ...)                         \
  do {                                          \
    snprintf(g_err, sizeof g_err, __VA_ARGS__); \
    return (code);                              \
  } while (0)

struct orc_handle {
  mpopis_cfg_t cfg;
  int64_t K, T, N, as, cs, ss;
  int nthreads;
  int env_set;
  int32_t n_cars;
  double car[MPOPIS_MAX_CARS][MPOPIS_CAR_NPARAMS];
  double dt, ddt;
  int64_t n_trk;
  double *tx, *ty, *tw;
  double mc[MPOPIS_MC_NPARAMS];
  int64_t mc_max_steps;
  double *Sigma; /* as x as for :mppi, cs x cs otherwise (POL:66-81) */
  int64_t sigma_n;
  mpopis_cma_t cma;
  double *ws;
  int cma_set;
  uint64_t seed;
  int64_t step;
  double *costs, *weights, *E, *traj, *Sigma_last, *U_last;
  double last_shrink;
  /* EnvpoolEnv seam (MPOPIS_ENV_EXTERNAL): bounds of the action space and the running plan's callback */
  double *ext_lo, *ext_hi;
  mpopis_rollout_fn ext_fn;
  void *ext_user;
};

/* ------------------------------------------------------------------------------------------ */
/* small dense linear algebra, column-major n x n                                              */
/* ------------------------------------------------------------------------------------------ */
#define A_(M, i, j, n) (M)[(i) + (size_t)(j) * (n)]

/* lower Cholesky factor, what PDMat(Σ) holds inside MvNormal(Σ). */
static int chol_lower(const double *A, int64_t n, double *L) {
  memset(L, 0, sizeof(double) * n * n);
  for (int64_t j = 0; j < n; ++j) {
    double d = A_(A, j, j, n);
    for (int64_t k = 0; k < j; ++k) d -= A_(L, j, k, n) * A_(L, j, k, n);
    if (!(d > 0.0)) return -1; /* PosDefException */
    double ljj = sqrt(d);
    A_(L, j, j, n) = ljj;
    for (int64_t i = j + 1; i < n; ++i) {
      double s = A_(A, i, j, n);
      for (int64_t k = 0; k < j; ++k) s -= A_(L, i, k, n) * A_(L, j, k, n);
      A_(L, i, j, n) = s / ljj;
    }
  }
  return 0;
}

/* invcov(P) = inv(Σ) through the Cholesky factor (Distributions.invcov -> inv(::PDMat)). */
static int inv_spd(const double *A, int64_t n, double *Ainv) {
  double *L = malloc(sizeof(double) * n * n), *Li = calloc(n * n, sizeof(double));
  if (chol_lower(A, n, L)) {
    free(L);
    free(Li);
    return -1;
  }
  for (int64_t j = 0; j < n; ++j) { /* Li = L^-1 (lower) by forward substitution */
    A_(Li, j, j, n) = 1.0 / A_(L, j, j, n);
    for (int64_t i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int64_t k = j; k < i; ++k) s -= A_(L, i, k, n) * A_(Li, k, j, n);
      A_(Li, i, j, n) = s / A_(L, i, i, n);
    }
  }
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j <= i; ++j) { /* Ainv = Li' * Li */
      double s = 0.0;
      for (int64_t k = i; k < n; ++k) s += A_(Li, k, i, n) * A_(Li, k, j, n);
      A_(Ainv, i, j, n) = s;
      A_(Ainv, j, i, n) = s;
    }
  free(L);
  free(Li);
  return 0;
}

/* symmetric eigendecomposition by cyclic Jacobi: A = V diag(ev) V' */
static void jacobi_eig(const double *Ain, int64_t n, double *ev, double *V) {
  double *A = malloc(sizeof(double) * n * n);
  memcpy(A, Ain, sizeof(double) * n * n);
  memset(V, 0, sizeof(double) * n * n);
  for (int64_t i = 0; i < n; ++i) A_(V, i, i, n) = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i < n; ++i) {
        if (i != j) off += A_(A, i, j, n) * A_(A, i, j, n);
        else diag += A_(A, i, j, n) * A_(A, i, j, n);
      }
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int64_t p = 0; p < n - 1; ++p)
      for (int64_t q = p + 1; q < n; ++q) {
        double apq = A_(A, p, q, n);
        if (apq == 0.0) continue;
        double app = A_(A, p, p, n), aqq = A_(A, q, q, n);
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int64_t k = 0; k < n; ++k) { /* A <- A J */
          double akp = A_(A, k, p, n), akq = A_(A, k, q, n);
          A_(A, k, p, n) = c * akp - s * akq;
          A_(A, k, q, n) = s * akp + c * akq;
        }
        for (int64_t k = 0; k < n; ++k) { /* A <- J' A */
          double apk = A_(A, p, k, n), aqk = A_(A, q, k, n);
          A_(A, p, k, n) = c * apk - s * aqk;
          A_(A, q, k, n) = s * apk + c * aqk;
        }
        for (int64_t k = 0; k < n; ++k) {
          double vkp = A_(V, k, p, n), vkq = A_(V, k, q, n);
          A_(V, k, p, n) = c * vkp - s * vkq;
          A_(V, k, q, n) = s * vkp + c * vkq;
        }
      }
  }
  for (int64_t i = 0; i < n; ++i) ev[i] = A_(A, i, i, n);
  free(A);
}

/* Σ^-0.5 for a symmetric matrix = V diag(λ^-1/2) V' (POL:580) */
static int sym_inv_sqrt(const double *A, int64_t n, double *C) {
  double *ev = malloc(sizeof(double) * n), *V = malloc(sizeof(double) * n * n);
  jacobi_eig(A, n, ev, V);
  int bad = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!(ev[i] > 0.0)) bad = 1;
  if (!bad)
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < n; ++j) {
        double s = 0.0;
        for (int64_t k = 0; k < n; ++k) s += A_(V, i, k, n) * (1.0 / sqrt(ev[k])) * A_(V, j, k, n);
        A_(C, i, j, n) = s;
      }
  free(ev);
  free(V);
  return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., Random123) + Box–Muller                                       */
/* ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

/* The engine's stream definition (DESIGN.md "RNG"): key = seed; counter =
 * (sample k, pair j, iteration | purpose<<24, control-step counter). purpose 0 = normals,
 * 1 = categorical uniforms. u = ((hi<<32|lo)>>12 + 0.5) * 2^-52, exact in double, strictly in (0,1). */
static void philox_u2(uint64_t seed, uint32_t k, uint32_t j, uint32_t it, uint32_t step, double *u1,
                      double *u2) {
  uint32_t ctr[4] = {k, j, it, step}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  orc_philox4x32_10(ctr, key, o);
  uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
  *u1 = ((double)(a >> 12) + 0.5) * 0x1.0p-52;
  *u2 = ((double)(b >> 12) + 0.5) * 0x1.0p-52;
}

static void philox_normals(uint64_t seed, int64_t step, int64_t it, int64_t cs, int64_t K, int64_t k0,
                           double *Z /* cs x K, column k holds global sample k0+k */) {
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int64_t k = 0; k < K; ++k)
    for (int64_t j = 0; 2 * j < cs; ++j) {
      double u1, u2;
      philox_u2(seed, (uint32_t)(k0 + k), (uint32_t)j, (uint32_t)it, (uint32_t)step, &u1, &u2);
      double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586 * u2;
      Z[2 * j + cs * k] = rad * cos(ang);
      if (2 * j + 1 < cs) Z[2 * j + 1 + cs * k] = rad * sin(ang);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* environments                                                                                */
/* ------------------------------------------------------------------------------------------ */
static double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }

/* calc_tire_fy, CAR:252-260 */
static double calc_tire_fy(double alpha, double mu, double C, double fzt, double fxt) {
  double fy_max = sqrt(fmax((mu * fzt) * (mu * fzt) - fxt * fxt, 1e-8));
  double ta = tan(alpha);
  if (fabs(alpha) < atan(3 * fy_max / C))
    return -C * ta + ((C * C) / (3 * fy_max)) * fabs(ta) * ta -
           ((C * C * C) / (27 * (fy_max * fy_max))) * (ta * ta * ta);
  return -fy_max * jl_sign(alpha);
}

/* calc_tire_fz, CAR:262-272 */
static double calc_tire_fz(const double *P, double fx, char tire) {
  double mass = P[0], l_t = P[3], h_cm = P[2], L = P[4] + P[3];
  if (tire == 'f') {
    l_t = P[4];
    h_cm *= -1;
  }
  return (mass * l_t * 9.81 + h_cm * fx) / L;
}

/* _step!(env::CarRacingEnv, a), CAR:282-344. P = the 18 params in declaration order CAR:2-21:
 * m Izz h_cm l_f l_r C_D0 C_D1 C_αf C_αr μ_f μ_r δ_max δ_dot_max Fx_max Fx_min λ_brake λ_drive β_limit */
static void car_step(const double *P, double dt, double ddt, double *s, const double *a) {
  double x = s[0], y = s[1], psi = s[2], Vx = s[3], Vy = s[4], psid = s[5], delta = s[6];
  const double m = P[0], Izz = P[1], l_f = P[3], l_r = P[4], C_D0 = P[5], C_D1 = P[6], C_af = P[7],
               C_ar = P[8], mu_f = P[9], mu_r = P[10], d_max = P[11], dd_max = P[12],
               Fx_max = P[13], Fx_min = P[14], l_brake = P[15], l_drive = P[16];
  double cmd_rate = fabs(a[0] * d_max - delta) / dt;                          /* CAR:295 */
  double rate = fmin(cmd_rate, dd_max) * jl_sign(a[0] * d_max - delta);      /* CAR:296 */
  double pedal = a[1];                                                        /* CAR:297 */
  long nsub = lrint(dt / ddt);                                                /* CAR:299 (ties-to-even) */
  for (long i = 0; i < nsub; ++i) {
    delta += rate * ddt;                                                      /* CAR:301 */
    double a_f = atan2(Vy + l_f * psid, Vx) - delta;                          /* CAR:304 */
    double a_r = atan2(Vy - l_r * psid, Vx);                                  /* CAR:305 */
    double fx_aero = (C_D0 + C_D1 * fabs(Vx)) * jl_sign(Vx);                  /* CAR:308 */
    double accel = Fx_max * fmax(pedal, 0.0);                                 /* CAR:310 */
    double brake = Fx_min * fmin(pedal, 0.0) * jl_sign(Vx);                   /* CAR:311 */
    double fx = accel + brake;                                                /* CAR:312 */
    double fxf = (pedal <= 0 ? l_brake : l_drive) * fx;                       /* CAR:315 */
    double fxr = (1 - (pedal <= 0 ? l_brake : l_drive)) * fx;                 /* CAR:316 */
    double fzf = calc_tire_fz(P, fx, 'f'), fzr = calc_tire_fz(P, fx, 'r');    /* CAR:317-318 */
    double fyf = calc_tire_fy(a_f, mu_f, C_af, fzf, fxf);                     /* CAR:319 */
    double fyr = calc_tire_fy(a_r, mu_r, C_ar, fzr, fxr);                     /* CAR:320 */
    double psidd = (1 / Izz) * (l_f * (fxf * sin(delta) + fyf * cos(delta)) - l_r * fyr); /* CAR:322 */
    double Vy_dot = (1 / m) * (fyf * cos(delta) + fxf * sin(delta) + fyr) - psid * Vx;    /* CAR:323 */
    double Vx_dot = (1 / m) * (fxf * cos(delta) - fyf * sin(delta) + fxr - fx_aero) + psid * Vy; /* CAR:324 */
    psid += psidd * ddt;                                                      /* CAR:326 */
    Vx += Vx_dot * ddt;                                                       /* CAR:327 */
    Vy += Vy_dot * ddt;                                                       /* CAR:328 */
    psi += psid * ddt;                                                        /* CAR:329 */
    psi = atan2(sin(psi), cos(psi));                                          /* CAR:330 */
    x += (Vx * cos(psi) - Vy * sin(psi)) * ddt;                               /* CAR:331 */
    y += (Vx * sin(psi) + Vy * cos(psi)) * ddt;                               /* CAR:332 */
  }
  s[0] = x, s[1] = y, s[2] = psi, s[3] = Vx, s[4] = Vy, s[5] = psid, s[6] = delta, s[7] = pedal;
}

/* within_track(track, pos), TRK:68-92. Returns within; 0-based indices. */
static int within_track(int64_t n, const double *tx, const double *ty, const double *tw, double px,
                        double py, int32_t *idx_out, int32_t *idx2_out, double *dist_out) {
  int64_t mi = 0;
  double best = 0.0;
  for (int64_t i = 0; i < n; ++i) { /* TRK:71,73 — findmin returns the FIRST minimum */
    double dx = tx[i] - px, dy = ty[i] - py;
    double d = dx * dx + dy * dy;
    if (i == 0 || d < best) best = d, mi = i;
  }
  int64_t m1 = (mi - 1 + n) % n, p1 = (mi + 1) % n; /* mod1, TRK:75-76 */
  double ax = tx[m1] - px, ay = ty[m1] - py, bx = tx[p1] - px, by = ty[p1] - py;
  double dist_m1 = sqrt(ax * ax + ay * ay), dist_p1 = sqrt(bx * bx + by * by); /* TRK:77-78 */
  int64_t m2 = dist_m1 <= dist_p1 ? m1 : p1;                                  /* TRK:79 */
  double p1x = tx[mi], p1y = ty[mi], p2x = tx[m2], p2y = ty[m2];
  double ux = px - p1x, uy = py - p1y, vx = p2x - p1x, vy = p2y - p1y;
  double t = (ux * vx + uy * vy) / (vx * vx + vy * vy);                       /* TRK:87 */
  double qx = p1x + t * vx, qy = p1y + t * vy;                                /* TRK:88 */
  double ex = qx - px, ey = qy - py;
  double dist = sqrt(ex * ex + ey * ey);                                      /* TRK:89 */
  if (idx_out) *idx_out = (int32_t)mi;
  if (idx2_out) *idx2_out = (int32_t)m2;
  *dist_out = dist;
  return dist < tw[mi];                                                       /* TRK:90 */
}

/* reward(env::CarRacingEnv), CAR:201-213 */
static double car_reward(const orc_t *h, const double *P, const double *s) {
  double rew = 0.0, dist;
  int within = within_track(h->n_trk, h->tx, h->ty, h->tw, s[0], s[1], NULL, NULL, &dist);
  if (!within) rew += -1000000.0;
  if (fabs(atan2(s[4], s[3])) > P[17]) rew += -5000.0; /* exceed_β, CAR:181-189 */
  rew += -dist;
  rew += 2.0 * sqrt(s[3] * s[3] + s[4] * s[4]);
  return rew;
}

/* (env::MultiCarRacingEnv)(a) MCR:200-207 and (env::CarRacingEnv)(a) CAR:238-241 */
static void cars_step(const orc_t *h, double *s, const double *a) {
  for (int i = 0; i < h->n_cars; ++i) car_step(h->car[i], h->dt, h->ddt, s + 8 * i, a + 2 * i);
}

/* reward(env::MultiCarRacingEnv), MCR:145-158 (collision penalty -11000) */
static double cars_reward(const orc_t *h, const double *s) {
  if (h->n_cars == 1) return car_reward(h, h->car[0], s);
  double rew = 0.0;
  for (int i = 0; i < h->n_cars; ++i) {
    rew += car_reward(h, h->car[i], s + 8 * i);
    for (int j = i + 1; j < h->n_cars; ++j) {
      double dx = s[8 * j] - s[8 * i], dy = s[8 * j + 1] - s[8 * i + 1];
      double dd = sqrt(dx * dx + dy * dy);
      rew += -dd;
      if (dd <= 4.0) rew += -11000.0;
    }
  }
  return rew;
}

/* RLEnvs MountainCarEnv(continuous=true) _step! */
static void mc_step(const orc_t *h, double *s, double a, int64_t *t, int *done) {
  const double min_pos = h->mc[0], max_pos = h->mc[1], max_speed = h->mc[2], goal_pos = h->mc[3],
               goal_vel = h->mc[4], power = h->mc[5], gravity = h->mc[6];
  *t += 1;
  double x = s[0], v = s[1];
  v += a * power + cos(3 * x) * (-gravity);
  v = fmin(fmax(v, -max_speed), max_speed);
  x += v;
  x = fmin(fmax(x, min_pos), max_pos);
  if (x == min_pos && v < 0) v = 0;
  *done = (x >= goal_pos && v >= goal_vel) || *t >= h->mc_max_steps;
  s[0] = x, s[1] = v;
}

/* reward(env::MountainCarEnv), EXM:10-22 */
static double mc_reward(const orc_t *h, const double *s, int done) {
  double rew = 0.0;
  if (s[0] >= h->mc[3] && s[1] >= h->mc[4]) rew += 100000;
  rew += fabs(s[1]);
  rew += done ? 0.0 : -1.0;
  return rew;
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* rollout_model(env, T, model_controls, pol, k), UTL:129-144, on a copy of the env (POL:270).
 * V = unclamped control vector (cs); action bounds are [-1,1] per component for all three
 * envs (CAR:156-159, MCR:75-84, RLEnvs continuous MountainCar). */
static double rollout(const orc_t *h, const double *state, int64_t env_t, const double *V,
                      double *traj /* T x ss col-major or NULL */) {
  double s[8 * MPOPIS_MAX_CARS], a[2 * MPOPIS_MAX_CARS];
  int64_t t_env = env_t;
  int done = 0;
  memcpy(s, state, sizeof(double) * h->ss);
  double traj_cost = 0.0;
  for (int64_t t = 0; t < h->T; ++t) {
    for (int64_t r = 0; r < h->as; ++r) a[r] = clampd(V[t * h->as + r], -1.0, 1.0); /* UTL:55-67 */
    double rew;
    if (h->cfg.env == MPOPIS_ENV_CAR_RACING) {
      cars_step(h, s, a);
      rew = cars_reward(h, s);
    } else {
      mc_step(h, s, a[0], &t_env, &done);
      rew = mc_reward(h, s, done);
    }
    traj_cost -= rew; /* UTL:138 */
    if (traj)
      for (int64_t q = 0; q < h->ss; ++q) traj[t + h->T * q] = s[q]; /* UTL:139-141 */
  }
  return traj_cost;
}

/* simulate_model(pol, env::AbstractEnv, E, Σ_inv, U_orig), POL:261-278 */
static void simulate_model(orc_t *h, const double *state, int64_t env_t, const double *U,
                           const double *U_orig, const double *E, const double *Sigma_inv,
                           double *costs, double *traj) {
  const int64_t K = h->K, cs = h->cs;
  const double gamma = h->cfg.lambda * (1 - h->cfg.alpha); /* POL:266 */
  double *row = NULL;
  if (Sigma_inv) { /* (γ * U_orig') * Σ_inv, left-associated as Julia parses POL:272 */
    row = malloc(sizeof(double) * cs);
    for (int64_t j = 0; j < cs; ++j) {
      double s = 0.0;
      for (int64_t i = 0; i < cs; ++i) s += (gamma * U_orig[i]) * A_(Sigma_inv, i, j, cs);
      row[j] = s;
    }
  }
  if (h->cfg.env == MPOPIS_ENV_EXTERNAL) { /* simulate_model(pol, env::EnvpoolEnv, …) POL:240-259 */
    const int64_t as = h->as, T = h->T;
    double *ctl = malloc(sizeof(double) * cs * K), *tc = malloc(sizeof(double) * K);
    for (int64_t k = 0; k < K; ++k) {
      double control_cost = 0.0;
      for (int64_t r = 0; r < cs; ++r) {
        const double v = U[r] + E[r + cs * k]; /* Vₖ = repeat(pol.U', K) + E', POL:252 */
        if (row) control_cost += row[r] * (v - U_orig[r]); /* POL:248 */
        /* get_model_controls(action_space, Vₖ, T) UTL:42-53: K x as x T, clamped per action component */
        ctl[k + K * r] = clampd(v, h->ext_lo[r % as], h->ext_hi[r % as]);
      }
      costs[k] = control_cost;
    }
    const int rc = h->ext_fn ? h->ext_fn(h->ext_user, ctl, K, as, T, tc) : -1; /* rollout_model UTL:103-121 */
    for (int64_t k = 0; k < K; ++k) costs[k] = rc ? NAN : tc[k] + costs[k]; /* POL:256 */
    free(ctl), free(tc), free(row);
    return;
  }
#pragma omp parallel for schedule(static) num_threads(h->nthreads)
  for (int64_t k = 0; k < K; ++k) { /* Threads.@threads for k ∈ 1:K, POL:269 */
    double V[cs];
    for (int64_t r = 0; r < cs; ++r) V[r] = U[r] + E[r + cs * k]; /* POL:271 */
    double control_cost = 0.0;
    if (row)
      for (int64_t r = 0; r < cs; ++r) control_cost += row[r] * (V[r] - U_orig[r]); /* POL:272 */
    double c = rollout(h, state, env_t, V, traj ? traj + (size_t)k * h->T * h->ss : NULL);
    costs[k] = c + control_cost; /* POL:274-275 */
  }
  free(row);
}

/* compute_weights(::Information_Theoretic, costs), UTL:79-86 */
static void compute_weights(const double *costs, int64_t K, double lambda, double *w) {
  double rho = costs[0];
  for (int64_t k = 1; k < K; ++k)
    if (costs[k] < rho) rho = costs[k];
  double eta = 0.0;
  for (int64_t k = 0; k < K; ++k) {
    w[k] = exp(-1 / lambda * (costs[k] - rho));
    eta += w[k];
  }
  for (int64_t k = 0; k < K; ++k) w[k] = w[k] / eta;
}

/* stable ascending arg-sort (Julia sortperm: merge sort, ties keep index order), POL:455,563 */
/* Base.isless for Float64: total order with -0.0 < +0.0 and NaN last (sortperm's default lt). */
static int jl_isless(double a, double b) {
  if (isnan(a)) return 0;
  if (isnan(b)) return 1;
  if (a < b) return 1;
  return a == b && signbit(a) && !signbit(b);
}
static void msort(const double *x, int64_t *a, int64_t *tmp, int64_t n) {
  if (n < 2) return;
  int64_t h = n / 2;
  msort(x, a, tmp, h);
  msort(x, a + h, tmp, n - h);
  int64_t i = 0, j = h, o = 0;
  while (i < h && j < n) tmp[o++] = jl_isless(x[a[j]], x[a[i]]) ? a[j++] : a[i++];
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, sizeof(int64_t) * n);
}
int orc_sortperm(const double *x, int64_t n, int64_t *perm) {
  int64_t *tmp = malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  msort(x, perm, tmp, n);
  free(tmp);
  return 0;
}

/* mean + covariance of the columns of X (p x n col-major) with optional weights.
 *   w == NULL, corrected = 0: cov(SimpleCovariance(), X')               (POL:464, :mle)
 *   w == NULL, corrected = 1: StatsBase.mean_and_cov(X, 2)              (POL:807)
 *   w != NULL:                StatsBase.mean_and_cov(X, ProbabilityWeights(w), 2)  (POL:364,662,732)
 *                             = Σ w (x-μ)(x-μ)' / Σ w, μ = Σ w x / Σ w   */
static void mean_and_cov(const double *X, int64_t p, int64_t n, const double *w, int corrected,
                         double *mu, double *S) {
  double sw = 0.0;
  for (int64_t k = 0; k < n; ++k) sw += w ? w[k] : 1.0;
  for (int64_t i = 0; i < p; ++i) {
    double s = 0.0;
    for (int64_t k = 0; k < n; ++k) s += (w ? w[k] : 1.0) * X[i + p * k];
    mu[i] = s / sw;
  }
  if (!S) return;
  double denom = w ? sw : (double)(n - (corrected ? 1 : 0));
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = 0; j <= i; ++j) {
      double s = 0.0;
      for (int64_t k = 0; k < n; ++k)
        s += (w ? w[k] : 1.0) * (X[i + p * k] - mu[i]) * (X[j + p * k] - mu[j]);
      A_(S, i, j, p) = A_(S, j, i, p) = s / denom;
    }
}

/* Σ_{i≠j} Var^(s_ij) of Schäfer & Strimmer (2005) p.11 for the UNCORRECTED covariance S = Z'Z/n of
 * centred data Z (n x p given as columns: Zc is p x n col-major): with w_kij = z_ki z_kj,
 * Var^(s_ij) = n/((n-1) n²) Σ_k (w_kij - w̄_ij)²   (LinearShrinkage default corrected=false). */
static double sum_var_sij(const double *Zc, const double *S, int64_t p, int64_t n) {
  double tot = 0.0;
  for (int64_t i = 0; i < p; ++i)
    for (int64_t j = 0; j < p; ++j) {
      if (i == j) continue;
      double s2 = 0.0;
      for (int64_t k = 0; k < n; ++k) {
        double wk = Zc[i + p * k] * Zc[j + p * k];
        s2 += wk * wk;
      }
      double sij = A_(S, i, j, p);
      tot += s2 - (double)n * sij * sij;
    }
  return tot * (double)n / ((double)(n - 1) * (double)n * (double)n);
}

/* cov(method, elite') of CovarianceEstimation.jl 0.2 (POL:414-426,464); SURVEY App. C-3.
 * UNPINNED: restated from the published formulas, not from the package source. */
static void cov_estimate(int method, const double *X, int64_t p, int64_t n, double *mu, double *S,
                         double *lambda_out) {
  mean_and_cov(X, p, n, NULL, 0, mu, S); /* SimpleCovariance(corrected=false) */
  double lam = 0.0;
  if (method == MPOPIS_SIGMA_MLE) {
    if (lambda_out) *lambda_out = 0.0;
    return;
  }
  if (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS) {
    double *Zc = malloc(sizeof(double) * p * n), *R = malloc(sizeof(double) * p * p);
    for (int64_t k = 0; k < n; ++k)
      for (int64_t i = 0; i < p; ++i) {
        double d = method == MPOPIS_SIGMA_SS ? 1.0 / sqrt(A_(S, i, i, p)) : 1.0;
        Zc[i + p * k] = (X[i + p * k] - mu[i]) * d;
      }
    for (int64_t i = 0; i < p; ++i)
      for (int64_t j = 0; j < p; ++j) {
        double di = method == MPOPIS_SIGMA_SS ? 1.0 / sqrt(A_(S, i, i, p)) : 1.0;
        double dj = method == MPOPIS_SIGMA_SS ? 1.0 / sqrt(A_(S, j, j, p)) : 1.0;
        A_(R, i, j, p) = A_(S, i, j, p) * di * dj;
      }
    double num = sum_var_sij(Zc, R, p, n), den = 0.0;
    for (int64_t i = 0; i < p; ++i)
      for (int64_t j = 0; j < p; ++j)
        if (i != j) den += A_(R, i, j, p) * A_(R, i, j, p);
    lam = num / den;
    lam = clampd(lam, 0.0, 1.0);
    for (int64_t i = 0; i < p; ++i) /* (1-λ)S + λ diag(S) */
      for (int64_t j = 0; j < p; ++j)
        if (i != j) A_(S, i, j, p) = (1.0 - lam) * A_(S, i, j, p);
    free(Zc);
    free(R);
  } else { /* DiagonalCommonVariance: F = tr(S)/p I; Chen, Wiesel, Eldar, Hero (2010) eq. 17/19, 23 */
    double trS = 0.0, trS2 = 0.0;
    for (int64_t i = 0; i < p; ++i) trS += A_(S, i, i, p);
    for (int64_t i = 0; i < p * p; ++i) trS2 += S[i] * S[i];
    double tr2S = trS * trS, pd = (double)p, nd = (double)n;
    if (method == MPOPIS_SIGMA_RBLW)
      lam = ((nd - 2) / nd * trS2 + tr2S) / ((nd + 2) * (trS2 - tr2S / pd));
    else
      lam = ((1.0 - 2.0 / pd) * trS2 + tr2S) / ((nd + 1.0 - 2.0 / pd) * (trS2 - tr2S / pd));
    lam = clampd(lam, 0.0, 1.0);
    double F = trS / pd;
    for (int64_t i = 0; i < p; ++i)
      for (int64_t j = 0; j < p; ++j)
        A_(S, i, j, p) = (1.0 - lam) * A_(S, i, j, p) + (i == j ? lam * F : 0.0);
  }
  if (lambda_out) *lambda_out = lam;
}

/* ------------------------------------------------------------------------------------------ */
/* handle                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static int is_g_family(int pol) { return pol != MPOPIS_POLICY_MPPI; }

int orc_create(const mpopis_cfg_t *cfg, orc_t **out) {
  if (!cfg || !out) FAIL(MPOPIS_ERR_BAD_ARG, "null argument");
  if (cfg->abi_version != MPOPIS_B200_ABI_VERSION) FAIL(MPOPIS_ERR_BAD_ARG, "abi version mismatch");
  if (cfg->policy < 0 || cfg->policy > MPOPIS_POLICY_PMCMPPI) FAIL(MPOPIS_ERR_BAD_ARG, "No policy_type");
  if (cfg->num_samples < 1 || cfg->horizon < 1 || cfg->opt_its < 1) FAIL(MPOPIS_ERR_BAD_ARG, "bad sizes");
  orc_t *h = calloc(1, sizeof *h);
  h->cfg = *cfg;
  h->K = cfg->num_samples, h->T = cfg->horizon;
  h->N = (cfg->policy == MPOPIS_POLICY_MPPI || cfg->policy == MPOPIS_POLICY_GMPPI) ? 1 : cfg->opt_its;
  if (cfg->env == MPOPIS_ENV_CAR_RACING) {
    if (cfg->n_cars < 1 || cfg->n_cars > MPOPIS_MAX_CARS) {
      free(h);
      FAIL(MPOPIS_ERR_BAD_ARG, "n_cars out of range");
    }
    h->n_cars = cfg->n_cars, h->as = 2 * cfg->n_cars, h->ss = 8 * cfg->n_cars;
  } else if (cfg->env == MPOPIS_ENV_MOUNTAIN_CAR) {
    h->n_cars = 0, h->as = 1, h->ss = 2;
  } else if (cfg->env == MPOPIS_ENV_EXTERNAL) {
    if (cfg->ext_action_size < 1 || cfg->ext_action_size > 4096 || cfg->log_trajectories) {
      free(h);
      FAIL(MPOPIS_ERR_BAD_ARG, "external env: bad ext_action_size / log_trajectories");
    }
    h->n_cars = 0, h->as = cfg->ext_action_size, h->ss = 0;
  } else {
    free(h);
    FAIL(MPOPIS_ERR_BAD_ARG, "unknown env");
  }
  h->cs = h->as * h->T; /* POL:59 */
  h->nthreads = 1;
  h->sigma_n = is_g_family(cfg->policy) ? h->cs : h->as; /* POL:66-74 */
  h->Sigma = calloc(h->sigma_n * h->sigma_n, sizeof(double));
  for (int64_t i = 0; i < h->sigma_n; ++i) A_(h->Sigma, i, i, h->sigma_n) = 1.0;
  h->costs = calloc(h->K, sizeof(double));
  h->weights = calloc(h->K, sizeof(double));
  h->E = calloc(h->cs * h->K, sizeof(double));
  h->Sigma_last = calloc(h->cs * h->cs, sizeof(double));
  h->U_last = calloc(h->cs, sizeof(double));
  if (cfg->log_trajectories) h->traj = calloc((size_t)h->K * h->T * h->ss, sizeof(double));
  *out = h;
  return 0;
}

int orc_destroy(orc_t *h) {
  if (!h) return 0;
  free(h->tx), free(h->ty), free(h->tw), free(h->Sigma), free(h->ws);
  free(h->costs), free(h->weights), free(h->E), free(h->traj), free(h->Sigma_last), free(h->U_last);
  free(h->ext_lo), free(h->ext_hi);
  free(h);
  return 0;
}

int orc_set_threads(orc_t *h, int nthreads) {
  h->nthreads = nthreads < 1 ? 1 : nthreads;
  g_threads = h->nthreads;
  return 0;
}

int orc_set_car_env(orc_t *h, int32_t n_cars, const double *params, double dt, double ddt,
                    const double *trk_x, const double *trk_y, const double *trk_w, int64_t n_trk) {
  if (h->cfg.env != MPOPIS_ENV_CAR_RACING || n_cars != h->n_cars) FAIL(MPOPIS_ERR_BAD_ARG, "env mismatch");
  if (n_trk < 2) FAIL(MPOPIS_ERR_BAD_ARG, "track needs >= 2 points");
  memcpy(h->car, params, sizeof(double) * MPOPIS_CAR_NPARAMS * n_cars);
  h->dt = dt, h->ddt = ddt, h->n_trk = n_trk;
  free(h->tx), free(h->ty), free(h->tw);
  h->tx = malloc(sizeof(double) * n_trk), h->ty = malloc(sizeof(double) * n_trk);
  h->tw = malloc(sizeof(double) * n_trk);
  memcpy(h->tx, trk_x, sizeof(double) * n_trk);
  memcpy(h->ty, trk_y, sizeof(double) * n_trk);
  memcpy(h->tw, trk_w, sizeof(double) * n_trk);
  h->env_set = 1;
  return 0;
}

int orc_set_mountaincar_env(orc_t *h, const double *params7, int64_t max_steps) {
  if (h->cfg.env != MPOPIS_ENV_MOUNTAIN_CAR) FAIL(MPOPIS_ERR_BAD_ARG, "env mismatch");
  memcpy(h->mc, params7, sizeof(double) * MPOPIS_MC_NPARAMS);
  h->mc_max_steps = max_steps;
  h->env_set = 1;
  return 0;
}

/* cov_mat handling of MPPI_Policy_Params, POL:66-81 + block_diagm UTL:9-21 */
int orc_set_sigma(orc_t *h, const double *Sigma, int64_t n) {
  const int64_t sn = h->sigma_n;
  if (n == sn) {
    memcpy(h->Sigma, Sigma, sizeof(double) * n * n);
  } else if (n == h->as) {
    memset(h->Sigma, 0, sizeof(double) * sn * sn);
    for (int64_t b = 0; b < sn; b += n)
      for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) A_(h->Sigma, b + i, b + j, sn) = A_(Sigma, i, j, n);
  } else {
    FAIL(MPOPIS_ERR_BAD_ARG, "Covariance matrix size problem");
  }
  return 0;
}

int orc_set_cma(orc_t *h, const mpopis_cma_t *cma, const double *ws, int64_t n_ws) {
  if (n_ws != h->K) FAIL(MPOPIS_ERR_BAD_ARG, "ws must have K entries");
  h->cma = *cma;
  free(h->ws);
  h->ws = malloc(sizeof(double) * n_ws);
  memcpy(h->ws, ws, sizeof(double) * n_ws);
  h->cma_set = 1;
  return 0;
}

int orc_seed(orc_t *h, uint64_t seed) {
  h->seed = seed;
  h->step = 0;
  return 0;
}

/* E = L * Z (Z cs x K), MvNormal sampling, SURVEY App. A-1 */
static void apply_L(const double *L, int64_t n, const double *Z, int64_t K, double *E) {
#pragma omp parallel for schedule(static) num_threads(g_threads) if (K > 64)
  for (int64_t k = 0; k < K; ++k)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int64_t j = 0; j <= i; ++j) s += A_(L, i, j, n) * Z[j + n * k];
      E[i + n * k] = s;
    }
}

/* get_controls_roll_U!, UTL:88-101 with the aliasing of SURVEY App. B-2: shift left by `as`,
 * the last `as` entries keep their values. */
static void controls_roll_U(const orc_t *h, const double *wc, double *U, double *control) {
  /* UTL:91: clamp to action_space(pol.env) — ±1 for the built-in envs (CAR:156-159, MCR:75-84, continuous
   * MountainCar), the caller's own bounds for the EnvpoolEnv-style external env */
  const int ext = h->cfg.env == MPOPIS_ENV_EXTERNAL && h->ext_lo && h->ext_hi;
  for (int64_t r = 0; r < h->as; ++r)
    control[r] = clampd(wc[r], ext ? h->ext_lo[r] : -1.0, ext ? h->ext_hi[r] : 1.0);
  if (h->T > 1) {
    for (int64_t r = 0; r + h->as < h->cs; ++r) U[r] = wc[r + h->as]; /* UTL:95 */
    /* UTL:96 is a self-assignment (pol.U aliases pol.params.U₀): no-op */
  } else {
    memcpy(U, wc, sizeof(double) * h->cs); /* UTL:98 */
  }
}

/* (pol::MPPI_Policy)(env) POL:121-146 with calculate_trajectory_costs POL:186-216 */
static int plan_mppi(orc_t *h, const double *state, int64_t env_t, double *U, const double *Z,
                     double *control, int32_t *its) {
  const int64_t K = h->K, T = h->T, as = h->as, cs = h->cs;
  const double gamma = h->cfg.lambda * (1 - h->cfg.alpha);
  double *L = malloc(sizeof(double) * as * as), *Sinv = malloc(sizeof(double) * as * as);
  if (chol_lower(h->Sigma, as, L) || inv_spd(h->Sigma, as, Sinv)) {
    free(L), free(Sinv);
    FAIL(MPOPIS_ERR_NOT_PD, "PosDefException: matrix is not positive definite");
  }
  /* E[k,t] = L * z  (rand(rng, P, K, T), POL:193); stored here as cs x K, r = t*as + a */
  for (int64_t k = 0; k < K; ++k)
    for (int64_t t = 0; t < T; ++t) apply_L(L, as, Z + t * as + cs * k, 1, h->E + t * as + cs * k);
  if (h->cfg.env == MPOPIS_ENV_EXTERNAL) { /* calculate_trajectory_costs(pol::MPPI_Policy, env::EnvpoolEnv) POL:148-184 */
    double *ctl = malloc(sizeof(double) * cs * K), *tc = malloc(sizeof(double) * K);
    for (int64_t k = 0; k < K; ++k) {
      double cc = 0.0;
      for (int64_t t = 0; t < T; ++t) {
        const double *Ei = h->E + t * as + cs * k, *ut = U + t * as;
        for (int64_t j = 0; j < as; ++j) { /* γ * uₜ' * Σ_inv * Eᵢ, POL:164 */
          double rj = 0.0;
          for (int64_t i = 0; i < as; ++i) rj += (gamma * ut[i]) * A_(Sinv, i, j, as);
          cc += rj * Ei[j];
        }
        for (int64_t r = 0; r < as; ++r) /* Vₜ + get_model_controls, POL:162,166 */
          ctl[k + K * (r + as * t)] = clampd(ut[r] + Ei[r], h->ext_lo[r], h->ext_hi[r]);
      }
      h->costs[k] = cc;
    }
    const int frc = h->ext_fn ? h->ext_fn(h->ext_user, ctl, K, as, T, tc) : -1;
    for (int64_t k = 0; k < K; ++k) h->costs[k] = frc ? NAN : tc[k] + h->costs[k]; /* POL:171 */
    free(ctl), free(tc);
  } else {
#pragma omp parallel for schedule(static) num_threads(h->nthreads)
  for (int64_t k = 0; k < K; ++k) { /* POL:198-214 */
    double s[8 * MPOPIS_MAX_CARS], a[2 * MPOPIS_MAX_CARS], cost = 0.0;
    int64_t t_env = env_t;
    int done = 0;
    memcpy(s, state, sizeof(double) * h->ss);
    for (int64_t t = 0; t < T; ++t) {
      const double *Ei = h->E + t * as + cs * k, *ut = U + t * as;
      double cc = 0.0;
      for (int64_t j = 0; j < as; ++j) { /* (γ * uₜ') * Σ_inv * Eᵢ, POL:204 */
        double rj = 0.0;
        for (int64_t i = 0; i < as; ++i) rj += (gamma * ut[i]) * A_(Sinv, i, j, as);
        cc += rj * Ei[j];
      }
      for (int64_t r = 0; r < as; ++r) a[r] = clampd(ut[r] + Ei[r], -1.0, 1.0); /* POL:203,205 */
      double rew;
      if (h->cfg.env == MPOPIS_ENV_CAR_RACING) {
        cars_step(h, s, a);
        rew = cars_reward(h, s);
      } else {
        mc_step(h, s, a[0], &t_env, &done);
        rew = mc_reward(h, s, done);
      }
      cost = cost - rew + cc; /* POL:208 */
      if (h->traj)
        for (int64_t q = 0; q < h->ss; ++q) h->traj[(size_t)k * T * h->ss + t + T * q] = s[q];
    }
    h->costs[k] = cost;
  }
  }
  compute_weights(h->costs, K, h->cfg.lambda, h->weights); /* POL:127 */
  double *wc = calloc(cs, sizeof(double));
  for (int64_t t = 0; t < T; ++t) /* POL:131-136 */
    for (int64_t k = 0; k < K; ++k)
      for (int64_t r = 0; r < as; ++r) wc[t * as + r] += h->weights[k] * h->E[t * as + r + cs * k];
  for (int64_t r = 0; r < cs; ++r) wc[r] = U[r] + wc[r]; /* POL:137 */
  memcpy(h->U_last, U, sizeof(double) * cs);
  memset(h->Sigma_last, 0, sizeof(double) * cs * cs);
  for (int64_t b = 0; b < cs; b += as)
    for (int64_t i = 0; i < as; ++i)
      for (int64_t j = 0; j < as; ++j) A_(h->Sigma_last, b + i, b + j, cs) = A_(h->Sigma, i, j, as);
  controls_roll_U(h, wc, U, control);
  *its = 1;
  free(wc), free(L), free(Sinv);
  return 0;
}


/* (pol::AbstractGMPPI_Policy)(env) POL:221-238 + calculate_trajectory_costs of every G-family
 * policy: :gmppi POL:303-315, :imppi POL:347-373, :cemppi POL:434-472, :cmamppi POL:532-606,
 * :μaismppi POL:644-671, :μΣaismppi POL:709-742, :pmcmppi POL:782-817. */
static int plan_g(orc_t *h, const double *state, int64_t env_t, double *U_inout, const double *Zall,
                  const double *resample_u, double *control, int32_t *its_out) {
  const int64_t K = h->K, cs = h->cs, N = h->N;
  const int pol = h->cfg.policy;
  const double gamma = h->cfg.lambda * (1 - h->cfg.alpha);
  int rc = 0, its = 0;
  double *U_orig = malloc(sizeof(double) * cs), *U = malloc(sizeof(double) * cs);
  double *Sig = malloc(sizeof(double) * cs * cs), *Sp = malloc(sizeof(double) * cs * cs);
  double *L = malloc(sizeof(double) * cs * cs), *Sinv = malloc(sizeof(double) * cs * cs);
  double *mu = malloc(sizeof(double) * cs), *ws = malloc(sizeof(double) * K);
  int64_t *order = malloc(sizeof(int64_t) * K);
  double *elite = NULL, *Cm = NULL, *psig = NULL, *pSig = NULL, *dw = NULL;
  memcpy(U_orig, U_inout, sizeof(double) * cs); /* U_orig = pol.U */
  memcpy(U, U_inout, sizeof(double) * cs);
  memcpy(Sig, h->Sigma, sizeof(double) * cs * cs); /* Σ′ = pol.Σ, POL:441,543,716,789 */
  double sigma = h->cma.sigma;                     /* POL:536 */
  int64_t m_elite = 0;
  if (pol == MPOPIS_POLICY_CEMPPI) m_elite = llrint((double)K * (1 - h->cfg.ce_elite_threshold)); /* POL:437 */
  if (pol == MPOPIS_POLICY_CMAMPPI) {
    if (!h->cma_set) {
      rc = MPOPIS_ERR_BAD_ARG;
      snprintf(g_err, sizeof g_err, "CMA constants not set");
      goto done;
    }
    m_elite = h->cma.m_elite;
    Cm = malloc(sizeof(double) * cs * cs);
    psig = calloc(cs, sizeof(double)), pSig = calloc(cs, sizeof(double)); /* POL:545 */
    dw = malloc(sizeof(double) * cs);
  }
  if ((pol == MPOPIS_POLICY_CEMPPI || pol == MPOPIS_POLICY_CMAMPPI) && N > 1 &&
      (m_elite < 2 || m_elite > K)) {
    rc = MPOPIS_ERR_BAD_ARG;
    snprintf(g_err, sizeof g_err, "m_elite out of range");
    goto done;
  }
  if (pol == MPOPIS_POLICY_CMAMPPI && N > 1 && cs * m_elite < K) {
    rc = MPOPIS_ERR_BAD_ARG; /* BoundsError in the reference: δs[order[ii]] with order[ii] > cs*m */
    snprintf(g_err, sizeof g_err, "BoundsError: CMA linear index exceeds cs*m_elite");
    goto done;
  }
  if (m_elite > 0) elite = malloc(sizeof(double) * cs * m_elite);

  for (int64_t n = 1; n <= N; ++n) {
    its = (int)n;
    /* P = MvNormal(Σ′); E = rand(rng, P, K); Σ_inv = invcov(P) */
    if (pol == MPOPIS_POLICY_CMAMPPI && N > 1) { /* POL:550-554 */
      for (int64_t i = 0; i < cs * cs; ++i) Sp[i] = sigma * sigma * Sig[i];
    } else {
      memcpy(Sp, Sig, sizeof(double) * cs * cs);
    }
    if (chol_lower(Sp, cs, L)) {
      rc = MPOPIS_ERR_NOT_PD;
      snprintf(g_err, sizeof g_err, "PosDefException: matrix is not positive definite (iteration %d)", (int)n);
      goto done;
    }
    apply_L(L, cs, Zall + (size_t)(n - 1) * cs * K, K, h->E);
    if (gamma != 0.0) inv_spd(Sp, cs, Sinv);
    memcpy(h->Sigma_last, Sp, sizeof(double) * cs * cs);
    simulate_model(h, state, env_t, U, U_orig, h->E, gamma != 0.0 ? Sinv : NULL, h->costs, h->traj);
    if (n >= N) break;
    if (pol == MPOPIS_POLICY_IMPPI || pol == MPOPIS_POLICY_MUAISMPPI) { /* POL:361-365, 659-663 */
      compute_weights(h->costs, K, pol == MPOPIS_POLICY_IMPPI ? h->cfg.lambda : h->cfg.lambda_ais, ws);
      mean_and_cov(h->E, cs, K, ws, 0, mu, NULL);
      for (int64_t r = 0; r < cs; ++r) U[r] = U[r] + mu[r];
    } else if (pol == MPOPIS_POLICY_MUSIGMAAISMPPI) { /* POL:729-734 */
      compute_weights(h->costs, K, h->cfg.lambda_ais, ws);
      mean_and_cov(h->E, cs, K, ws, 0, mu, Sig);
      for (int64_t i = 0; i < cs; ++i) A_(Sig, i, i, cs) += 10e-9;
      for (int64_t r = 0; r < cs; ++r) U[r] = U[r] + mu[r];
    } else if (pol == MPOPIS_POLICY_PMCMPPI) { /* POL:802-809 */
      compute_weights(h->costs, K, h->cfg.lambda_ais, ws);
      /* Categorical(ws) + rand(rng, cat, K): restated as an inverse-CDF draw from injected /
       * Philox uniforms (Julia's alias-table stream is not reproducible; SURVEY §7 "hard parts") */
      double *cdf = malloc(sizeof(double) * K), *Er = malloc(sizeof(double) * cs * K), acc = 0.0;
      for (int64_t k = 0; k < K; ++k) acc += ws[k], cdf[k] = acc;
      for (int64_t i = 0; i < K; ++i) {
        double u = resample_u[i + K * (n - 1)];
        int64_t lo = 0, hi = K - 1;
        while (lo < hi) {
          int64_t mid = (lo + hi) / 2;
          if (u < cdf[mid]) hi = mid;
          else lo = mid + 1;
        }
        memcpy(Er + cs * i, h->E + cs * lo, sizeof(double) * cs); /* E′ = E[:, resample_idxs] */
      }
      mean_and_cov(Er, cs, K, NULL, 1, mu, Sig);
      for (int64_t i = 0; i < cs; ++i) A_(Sig, i, i, cs) += 10e-9;
      for (int64_t r = 0; r < cs; ++r) U[r] = U[r] + mu[r];
      free(cdf), free(Er);
    } else if (pol == MPOPIS_POLICY_CEMPPI || pol == MPOPIS_POLICY_CMAMPPI) {
      orc_sortperm(h->costs, K, order); /* POL:455,563 */
      for (int64_t j = 0; j < m_elite; ++j)
        memcpy(elite + cs * j, h->E + cs * order[j], sizeof(double) * cs); /* POL:456,564 */
      double maxdiff = -INFINITY; /* maximum(abs.(diff(elite_traj_cost))) < 10e-3, POL:458-461 */
      for (int64_t j = 0; j + 1 < m_elite; ++j) { /* Julia's maximum propagates NaN (then NaN < 10e-3 is false) */
        double d = fabs(h->costs[order[j + 1]] - h->costs[order[j]]);
        if (isnan(d) || isnan(maxdiff)) maxdiff = NAN;
        else if (d > maxdiff) maxdiff = d;
      }
      if (h->cfg.early_stop && maxdiff < 10e-3) break;
      if (pol == MPOPIS_POLICY_CEMPPI) { /* POL:464-465 */
        cov_estimate(h->cfg.sigma_est, elite, cs, m_elite, mu, Sig, &h->last_shrink);
        for (int64_t i = 0; i < cs; ++i) A_(Sig, i, i, cs) += 10e-9;
        for (int64_t r = 0; r < cs; ++r) U[r] = U[r] + mu[r];
      } else { /* CMA, POL:571-599; quirks of SURVEY App. B-1 reproduced literally */
        const mpopis_cma_t *c = &h->cma;
        const double sigma_ds = sigma; /* δs = elite_E / σ uses σ before POL:582 updates it */
        for (int64_t r = 0; r < cs; ++r) { /* POL:573-576 */
          double s = 0.0;
          for (int64_t j = 0; j < m_elite; ++j) s += h->ws[j] * elite[r + cs * j];
          dw[r] = s;
        }
        for (int64_t r = 0; r < cs; ++r) U[r] += sigma * dw[r]; /* POL:577 */
        if (sym_inv_sqrt(Sig, cs, Cm)) {                      /* POL:580 */
          rc = MPOPIS_ERR_NOT_PD;
          snprintf(g_err, sizeof g_err, "Σ^-0.5 of a non-PD matrix (iteration %d)", (int)n);
          goto done;
        }
        double cf = sqrt(c->c_sigma * (2 - c->c_sigma) * c->mu_eff), nps = 0.0, normC = 0.0;
        for (int64_t i = 0; i < cs; ++i) { /* POL:581 */
          double s = 0.0;
          for (int64_t j = 0; j < cs; ++j) s += A_(Cm, i, j, cs) * dw[j];
          psig[i] = (1 - c->c_sigma) * psig[i] + cf * s;
          nps += psig[i] * psig[i];
        }
        nps = sqrt(nps);
        for (int64_t i = 0; i < cs * cs; ++i) normC += Cm[i] * Cm[i];
        normC = sqrt(normC);                                           /* ‖C‖_F */
        sigma *= exp(c->c_sigma / c->d_sigma * (nps / c->E_norm - 1)); /* POL:582 */
        int hs = nps / sqrt(1 - pow(1 - c->c_sigma, 2.0 * (double)n)) <
                 (1.4 + 2.0 / ((double)cs + 1)) * c->E_norm;           /* POL:585 */
        double cg = hs * sqrt(c->c_Sigma * (2 - c->c_Sigma) * c->mu_eff);
        for (int64_t i = 0; i < cs; ++i) pSig[i] = (1 - c->c_Sigma) * pSig[i] + cg * dw[i]; /* POL:586 */
        double temp_sum = 0.0; /* POL:588-596: δs[order[ii]] is a LINEAR (1-based) index into the
                                  cs x m matrix δs; 0-based linear index = order[ii] as stored here */
        for (int64_t ii = 0; ii < K; ++ii) {
          double d = elite[order[ii]] / sigma_ds, w0;
          if (h->ws[ii] >= 0) {
            w0 = h->ws[ii];
          } else {
            double nrm = fabs(d) * normC; /* norm(C * scalar) = |scalar| ‖C‖_F */
            w0 = (double)n * h->ws[ii] / (nrm * nrm); /* n = AIS iteration counter, POL:593 */
          }
          temp_sum += w0 * d * d;
        }
        /* POL:598: Σ = (1-c1-cμ)Σ + c1 (pΣ pΣ' + (1-hσ) cΣ (2-cΣ) Σ) .+ cμ temp_sum ; POL:599 symmetrise */
        for (int64_t j = 0; j < cs; ++j)
          for (int64_t i = 0; i <= j; ++i) {
            double sij = A_(Sig, i, j, cs);
            double v = (1 - c->c1 - c->c_mu) * sij +
                       c->c1 * (pSig[i] * pSig[j] + (1 - hs) * c->c_Sigma * (2 - c->c_Sigma) * sij) +
                       c->c_mu * temp_sum;
            A_(Sig, i, j, cs) = v;
          }
        for (int64_t j = 0; j < cs; ++j)
          for (int64_t i = j + 1; i < cs; ++i) A_(Sig, i, j, cs) = A_(Sig, j, i, cs); /* triu + triu' */
      }
    }
  }
  /* E = E .+ (pol.U - U_orig); pol.U = U_orig (POL:370-371,468-469,602-603,...) */
  if (N > 1)
    for (int64_t k = 0; k < K; ++k)
      for (int64_t r = 0; r < cs; ++r) h->E[r + cs * k] = h->E[r + cs * k] + (U[r] - U_orig[r]);
  memcpy(h->U_last, U, sizeof(double) * cs);
  compute_weights(h->costs, K, h->cfg.lambda, h->weights); /* final weights always use λ */
  {
    double *wc = malloc(sizeof(double) * cs);
    for (int64_t r = 0; r < cs; ++r) { /* POL:226-230 */
      double s = 0.0;
      for (int64_t k = 0; k < K; ++k) s += h->weights[k] * h->E[r + cs * k];
      wc[r] = U_orig[r] + s;
    }
    memcpy(U_inout, U_orig, sizeof(double) * cs);
    controls_roll_U(h, wc, U_inout, control); /* POL:231 */
    free(wc);
  }
  *its_out = its;
done:
  free(U_orig), free(U), free(Sig), free(Sp), free(L), free(Sinv), free(mu), free(ws), free(order);
  free(elite), free(Cm), free(psig), free(pSig), free(dw);
  return rc;
}

int orc_plan_with_noise(orc_t *h, const double *state, int64_t env_t, double *U_inout,
                        const double *Z, const double *resample_u, double *control_out,
                        int32_t *its_run_out) {
  if (!h || !state || !U_inout || !Z || !control_out) FAIL(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!h->env_set) FAIL(MPOPIS_ERR_BAD_ARG, "environment not set");
  if (h->cfg.policy == MPOPIS_POLICY_PMCMPPI && h->N > 1 && !resample_u)
    FAIL(MPOPIS_ERR_BAD_ARG, "pmcmppi needs resample_u");
  int32_t its = 0;
  int rc = h->cfg.policy == MPOPIS_POLICY_MPPI
               ? plan_mppi(h, state, env_t, U_inout, Z, control_out, &its)
               : plan_g(h, state, env_t, U_inout, Z, resample_u, control_out, &its);
  if (its_run_out) *its_run_out = its;
  return rc;
}

int orc_plan(orc_t *h, const double *state, int64_t env_t, double *U_inout, double *control_out,
             int32_t *its_run_out) {
  const int64_t K = h->K, cs = h->cs, N = h->N;
  double *Z = malloc(sizeof(double) * cs * K * N), *u = NULL;
  for (int64_t n = 0; n < N; ++n) philox_normals(h->seed, h->step, n, cs, K, 0, Z + (size_t)n * cs * K);
  if (h->cfg.policy == MPOPIS_POLICY_PMCMPPI && N > 1) {
    u = malloc(sizeof(double) * K * (N - 1));
    for (int64_t n = 0; n + 1 < N; ++n)
      for (int64_t i = 0; i < K; ++i) {
        double u2;
        philox_u2(h->seed, (uint32_t)i, 0u, (uint32_t)n | (1u << 24), (uint32_t)h->step, &u[i + K * n], &u2);
      }
  }
  int rc = orc_plan_with_noise(h, state, env_t, U_inout, Z, u, control_out, its_run_out);
  h->step += 1;
  free(Z), free(u);
  return rc;
}

int orc_set_external_env(orc_t *h, const double *action_lo, const double *action_hi) {
  if (!h || !action_lo || !action_hi) FAIL(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_EXTERNAL) FAIL(MPOPIS_ERR_BAD_ARG, "handle was created for a different environment");
  free(h->ext_lo), free(h->ext_hi);
  h->ext_lo = malloc(sizeof(double) * h->as), h->ext_hi = malloc(sizeof(double) * h->as);
  memcpy(h->ext_lo, action_lo, sizeof(double) * h->as);
  memcpy(h->ext_hi, action_hi, sizeof(double) * h->as);
  h->env_set = 1;
  return 0;
}

/* pol(env::EnvpoolEnv): the same plans with the rollouts delegated to the caller's simulator */
int orc_plan_external(orc_t *h, double *U_inout, mpopis_rollout_fn rollout, void *user, const double *Z,
                      const double *resample_u, double *control_out, int32_t *its_run_out) {
  if (!h || !U_inout || !rollout || !control_out) FAIL(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_EXTERNAL) FAIL(MPOPIS_ERR_BAD_ARG, "handle was not created with MPOPIS_ENV_EXTERNAL");
  const double dummy_state = 0.0;
  h->ext_fn = rollout, h->ext_user = user;
  const int rc = Z ? orc_plan_with_noise(h, &dummy_state, 0, U_inout, Z, resample_u, control_out, its_run_out)
                   : orc_plan(h, &dummy_state, 0, U_inout, control_out, its_run_out);
  h->ext_fn = NULL, h->ext_user = NULL;
  return rc;
}

int orc_fetch(orc_t *h, double *costs, double *weights, double *E, double *traj) {
  if (costs) memcpy(costs, h->costs, sizeof(double) * h->K);
  if (weights) memcpy(weights, h->weights, sizeof(double) * h->K);
  if (E) memcpy(E, h->E, sizeof(double) * h->cs * h->K);
  if (traj) {
    if (!h->traj) FAIL(MPOPIS_ERR_BAD_ARG, "log_trajectories was not enabled");
    memcpy(traj, h->traj, sizeof(double) * (size_t)h->K * h->T * h->ss);
  }
  return 0;
}

int orc_fetch_proposal(orc_t *h, double *Sigma_last, double *U_last) {
  if (Sigma_last) memcpy(Sigma_last, h->Sigma_last, sizeof(double) * h->cs * h->cs);
  if (U_last) memcpy(U_last, h->U_last, sizeof(double) * h->cs);
  return 0;
}

int orc_rollout_costs(orc_t *h, const double *state, int64_t env_t, const double *U,
                      const double *U_orig, const double *E, const double *Sigma_inv,
                      double *costs_out) {
  if (!h->env_set) FAIL(MPOPIS_ERR_BAD_ARG, "environment not set");
  const double gamma = h->cfg.lambda * (1 - h->cfg.alpha);
  if (gamma != 0.0 && !Sigma_inv) FAIL(MPOPIS_ERR_BAD_ARG, "Sigma_inv required when γ != 0");
  simulate_model(h, state, env_t, U, U_orig, E, gamma != 0.0 ? Sigma_inv : NULL, costs_out, h->traj);
  return 0;
}

int orc_weights(orc_t *h, const double *costs, int64_t K, double lambda, double *w_out) {
  (void)h;
  compute_weights(costs, K, lambda, w_out);
  return 0;
}

int orc_track_query(orc_t *h, const double *pos, int64_t n, int32_t *idx_out, int32_t *idx2_out,
                    double *dist_out, uint8_t *within_out) {
  if (!h->env_set || h->cfg.env != MPOPIS_ENV_CAR_RACING) FAIL(MPOPIS_ERR_BAD_ARG, "car env not set");
  for (int64_t i = 0; i < n; ++i) {
    int32_t a, b;
    double d;
    int w = within_track(h->n_trk, h->tx, h->ty, h->tw, pos[2 * i], pos[2 * i + 1], &a, &b, &d);
    if (idx_out) idx_out[i] = a;
    if (idx2_out) idx2_out[i] = b;
    if (dist_out) dist_out[i] = d;
    if (within_out) within_out[i] = (uint8_t)w;
  }
  return 0;
}

int orc_env_step(orc_t *h, double *state, const double *action, int64_t *env_t, double *reward_out,
                 uint8_t *done_out) {
  if (!h->env_set) FAIL(MPOPIS_ERR_BAD_ARG, "environment not set");
  int done = 0;
  double rew;
  if (h->cfg.env == MPOPIS_ENV_CAR_RACING) {
    cars_step(h, state, action);
    *env_t += 1; /* CAR:283 */
    rew = cars_reward(h, state);
  } else {
    mc_step(h, state, action[0], env_t, &done);
    rew = mc_reward(h, state, done);
  }
  if (reward_out) *reward_out = rew;
  if (done_out) *done_out = (uint8_t)done;
  return 0;
}

int orc_env_reward(orc_t *h, const double *state, uint8_t done, double *reward_out) {
  if (!h->env_set) FAIL(MPOPIS_ERR_BAD_ARG, "environment not set");
  *reward_out = h->cfg.env == MPOPIS_ENV_CAR_RACING ? cars_reward(h, state) : mc_reward(h, state, done);
  return 0;
}

int orc_sample_normals(orc_t *h, int64_t step, int64_t iteration, double *Z_out) {
  philox_normals(h->seed, step, iteration, h->cs, h->K, 0, Z_out);
  return 0;
}

int orc_cov_estimate(orc_t *h, int32_t sigma_est, const double *X, int64_t p, int64_t n,
                     const double *w, int32_t corrected, double *mean_out, double *cov_out) {
  double *mu = mean_out ? mean_out : malloc(sizeof(double) * p);
  if (w || corrected) mean_and_cov(X, p, n, w, corrected, mu, cov_out);
  else cov_estimate(sigma_est, X, p, n, mu, cov_out, &h->last_shrink);
  if (!mean_out) free(mu);
  return 0;
}

int orc_cholesky(orc_t *h, const double *A, int64_t n, double *L_out) {
  (void)h;
  if (chol_lower(A, n, L_out)) FAIL(MPOPIS_ERR_NOT_PD, "PosDefException: matrix is not positive definite");
  return 0;
}

int orc_inv_sqrt(orc_t *h, const double *A, int64_t n, double *C_out) {
  (void)h;
  if (sym_inv_sqrt(A, n, C_out)) FAIL(MPOPIS_ERR_NOT_PD, "matrix is not positive definite");
  return 0;
}

int orc_last_shrinkage(orc_t *h, double *lambda_out) {
  *lambda_out = h->last_shrink;
  return 0;
}
