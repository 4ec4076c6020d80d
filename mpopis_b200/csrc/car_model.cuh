// car_model.cuh — device arithmetic of the CarRacing step and reward (CAR:178-344, TRK:68-92), shared by the
// rollout kernels (rollout.cu) and by the host-compiled checker tools/host_step_check.cu, which runs the very same
// functions on the CPU (MPOPIS_HD = __host__ __device__) to compare the fast formulations with the literal one
// without a GPU. See rollout.cu for the description of the variants (MODE 0..3).
#pragma once
#include <math.h>
#include <string.h>
#include <math_constants.h>

#include "engine.cuh"

#define MPOPIS_HD __host__ __device__ __forceinline__

namespace mpopis {

// fdlibm __kernel_sin / __kernel_cos minimax coefficients (|x| <= π/4, error < 2^-57)
#define MPOPIS_SINCOS_TABLE                                                                               \
  {-1.66666666666666324348e-01, 8.33333333332248946124e-03,  -1.98412698298579493134e-04,                \
   2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10,                 \
   4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,                 \
   -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11}
__constant__ double kSinCos[12] = MPOPIS_SINCOS_TABLE;  // constant bank: DFMAs read them as c[bank][off] operands
static const double kSinCosHost[12] = MPOPIS_SINCOS_TABLE;

// Device intrinsics with host stand-ins (host = the CPU checker only; compiled with -ffp-contract=off)
#ifdef __CUDA_ARCH__
#define MPOPIS_SC(i) kSinCos[i]
MPOPIS_HD double rcp_seed(double d) {  // 2^-23 reciprocal seed (MUFU.RCP64H)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}
MPOPIS_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
MPOPIS_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
MPOPIS_HD double sqrt_rn(double a) { return __dsqrt_rn(a); }
MPOPIS_HD double rsqrt_f64(double a) { return rsqrt(a); }
MPOPIS_HD uint4 ldg_u4(const uint4 *p) { return __ldg(p); }
MPOPIS_HD double inf_f64() { return CUDART_INF; }
#else
#define MPOPIS_SC(i) kSinCosHost[i]
MPOPIS_HD double rcp_seed(double d) { return (double)(float)(1.0 / d); }
MPOPIS_HD double mul_rn(double a, double b) { return a * b; }
MPOPIS_HD double add_rn(double a, double b) { return a + b; }
MPOPIS_HD double sqrt_rn(double a) { return sqrt(a); }
MPOPIS_HD double rsqrt_f64(double a) { return 1.0 / sqrt(a); }
MPOPIS_HD uint4 ldg_u4(const uint4 *p) { return *p; }
MPOPIS_HD double inf_f64() { return HUGE_VAL; }
#endif

MPOPIS_HD double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
MPOPIS_HD double clamp1(double v) { return fmin(fmax(v, -1.0), 1.0); }

// sin/cos on |x| <= ~π/4 (no range reduction)
MPOPIS_HD void sincos_kernel(double x, double *s, double *c) {
  const double z = x * x;
  double ps = fma(z, MPOPIS_SC(5), MPOPIS_SC(4));
  ps = fma(z, ps, MPOPIS_SC(3));
  ps = fma(z, ps, MPOPIS_SC(2));
  ps = fma(z, ps, MPOPIS_SC(1));
  ps = fma(z, ps, MPOPIS_SC(0));
  *s = fma(x * z, ps, x);
  double pc = fma(z, MPOPIS_SC(11), MPOPIS_SC(10));
  pc = fma(z, pc, MPOPIS_SC(9));
  pc = fma(z, pc, MPOPIS_SC(8));
  pc = fma(z, pc, MPOPIS_SC(7));
  pc = fma(z, pc, MPOPIS_SC(6));
  *c = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// sin/cos for |x| <= π(1+ε): quadrant reduction with a two-term π/2, then the kernels
MPOPIS_HD void sincos_pi(double x, double *s, double *c) {
  const double kf = rint(x * 0.63661977236758138);
  double r = fma(-kf, 1.5707963267948966, x);
  r = fma(-kf, 6.123233995736766e-17, r);
  const int k = (int)kf;
  double sr, cr;
  sincos_kernel(r, &sr, &cr);
  const double s1 = (k & 1) ? cr : sr, c1 = (k & 1) ? -sr : cr;
  *s = (k & 2) ? -s1 : s1;
  *c = (k & 2) ? -c1 : c1;
}

// n/d without the IEEE slow path: reciprocal seed (2^-23), two Newton steps, one residual correction
MPOPIS_HD double fast_div(double n, double d) {
  double r = rcp_seed(d);
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
}

// sqrt without the IEEE slow-path subroutine: reciprocal-square-root seed (MUFU.RSQ64H, 2^-23), one coupled Newton step
// for (sqrt, 1/(2 sqrt)) and a final residual correction — the result is the correctly rounded square root except for
// rare last-bit ties (≤ 1 ulp). Used where the reference takes norm(...) / sqrt of a sum of squares that feeds a smooth
// cost term; a threshold comparison can flip only if the exact value sits within an ulp of the threshold.
MPOPIS_HD double sqrt_fast(double x) {
#ifdef __CUDA_ARCH__
  if (!(x > 0.0) || !(x < 1.0e300)) return sqrt_rn(x);  // 0, negative, Inf, NaN: the exact routine (rare)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double y = x * r, h = 0.5 * r;
  const double e = fma(-y, h, 0.5);
  y = fma(y, e, y), h = fma(h, e, h);       // 2^-46
  y = fma(fma(-y, y, x), h, y);             // 2^-92 before rounding
  return fma(fma(-y, y, x), h, y);
#else
  return sqrt(x);
#endif
}

// 1/sqrt(x) for x > 0 without libdevice's special-case branch (its slow-path CALL splits the basic block and keeps
// ptxas from interleaving the per-step constants with the sub-step recurrence): reciprocal-square-root seed (2^-23) and
// one third-order correction r(1 + e/2 + 3e²/8), e = 1 − x r²: relative error e³ = 2^-69 before the final rounding.
MPOPIS_HD double rsqrt_pos(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-(x * r), r, 1.0);
  return fma(r, e * fma(0.375, e, 0.5), r);
#else
  return 1.0 / sqrt(x);
#endif
}

struct TrackView {
  const double *x, *y, *w;
  int n;
  // exact pruning table (nullptr = always scan): cell -> {count, up to 7 candidate indices} as 8 x u16
  const uint4 *lut;
  double x0, y0, inv_c;
  int nx, ny;
};

// within_track(track, pos) TRK:68-92. Integer-exact: distances use un-fused mul/add.
template <bool USE_LUT, bool FAST = false>
MPOPIS_HD bool within_track(const TrackView &tr, double px, double py, int *idx_out,
                                             int *idx2_out, double *dist_out) {
  int mi = 0;
  double best = inf_f64();
  bool done = false;
  if (USE_LUT && tr.lut) {
    const double fx = (px - tr.x0) * tr.inv_c, fy = (py - tr.y0) * tr.inv_c;
    if (fx >= 0.0 && fy >= 0.0 && fx < (double)tr.nx && fy < (double)tr.ny) {
      const uint4 cell = ldg_u4(tr.lut + (int)fy * tr.nx + (int)fx);
      const unsigned cnt = cell.x & 0xffffu;
      if (cnt <= 7u) {
        const unsigned long long w0 = ((unsigned long long)cell.y << 32) | cell.x;
        const unsigned long long w1 = ((unsigned long long)cell.w << 32) | cell.z;
        for (unsigned q = 0; q < cnt; ++q) {  // candidates are stored in ascending index order
          const int i = (int)(((q < 3u) ? (w0 >> (16u * q + 16u)) : (w1 >> (16u * (q - 3u)))) & 0xffffu);
          const double dx = tr.x[i] - px, dy = tr.y[i] - py;
          const double d = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
          if (d < best) best = d, mi = i;
        }
        done = true;
      }
    }
  }
  if (!done) {
#pragma unroll 4
    for (int i = 0; i < tr.n; ++i) {  // TRK:71,73 (findmin -> FIRST minimum: strict <)
      const double dx = tr.x[i] - px, dy = tr.y[i] - py;
      const double d = add_rn(mul_rn(dx, dx), mul_rn(dy, dy));
      if (d < best) best = d, mi = i;
    }
  }
  const int m1 = mi == 0 ? tr.n - 1 : mi - 1, p1 = mi == tr.n - 1 ? 0 : mi + 1;  // mod1, TRK:75-76
  const double ax = tr.x[m1] - px, ay = tr.y[m1] - py, bx = tr.x[p1] - px, by = tr.y[p1] - py;
  // TRK:77-79 compares norms: sqrt is monotone and correctly rounded, so the squared distances decide
  // unless they are within a few ulps of each other — only then are the two square roots taken.
  const double qa = add_rn(mul_rn(ax, ax), mul_rn(ay, ay));
  const double qb = add_rn(mul_rn(bx, bx), mul_rn(by, by));
  int m2;
  if (USE_LUT && fabs(qa - qb) > 1e-14 * fmax(qa, qb)) m2 = qa < qb ? m1 : p1;
  else m2 = sqrt_rn(qa) <= sqrt_rn(qb) ? m1 : p1;
  const double p1x = tr.x[mi], p1y = tr.y[mi];
  const double vx = tr.x[m2] - p1x, vy = tr.y[m2] - p1y, ux = px - p1x, uy = py - p1y;
  // TRK:87 (projection on the infinite line); FAST: reciprocal-Newton division / rsqrt-Newton square root instead of
  // the IEEE subroutines (≈ 35 instructions and a CALL each) — same value up to a last-bit tie
  const double t = FAST ? fast_div(ux * vx + uy * vy, vx * vx + vy * vy) : (ux * vx + uy * vy) / (vx * vx + vy * vy);
  const double ex = p1x + t * vx - px, ey = p1y + t * vy - py; // TRK:88
  const double dist = FAST ? sqrt_fast(ex * ex + ey * ey) : sqrt(ex * ex + ey * ey);  // TRK:89
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

MPOPIS_HD TireConsts tire_consts(const CarParams &P, double accel, double bk,
                                                  double split, double sgnVx) {
  TireConsts c;
  const double fx = accel + bk * sgnVx;  // CAR:310-312
  c.fxf = split * fx;
  c.fxr = (1 - split) * fx;
  const double L = P.l_r + P.l_f;
  const double fzf = (P.m * P.l_r * 9.81 - P.h_cm * fx) / L;  // calc_tire_fz 'f', CAR:262-272
  const double fzr = (P.m * P.l_f * 9.81 + P.h_cm * fx) / L;  // calc_tire_fz 'r'
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

// Same constants with one rsqrt per tyre instead of a sqrt and three divisions (fast v3): with
// v = max((μ fz)² − fx², 1e-8): fy_max = v·rsqrt(v), 1/fy_max = rsqrt(v); divisions by the car's constants
// become multiplications by their reciprocals (loop-invariant, hoisted by the compiler).
MPOPIS_HD TireConsts tire_consts_fast(const CarParams &P, double accel, double bk,
                                                       double split, double sgnVx) {
  TireConsts c;
  const double fx = accel + bk * sgnVx;  // CAR:310-312
  c.fxf = split * fx;
  c.fxr = (1 - split) * fx;
  const double invL = 1.0 / (P.l_r + P.l_f);
  const double fzf = (P.m * P.l_r * 9.81 - P.h_cm * fx) * invL;
  const double fzr = (P.m * P.l_f * 9.81 + P.h_cm * fx) * invL;
  const double vf = fmax((P.mu_f * fzf) * (P.mu_f * fzf) - c.fxf * c.fxf, 1e-8);
  const double vr = fmax((P.mu_r * fzr) * (P.mu_r * fzr) - c.fxr * c.fxr, 1e-8);
  const double rf = rsqrt_f64(vf), rr = rsqrt_f64(vr);
  c.fymax_f = vf * rf;
  c.fymax_r = vr * rr;
  c.thr_f = c.fymax_f * (3.0 / P.C_af);
  c.thr_r = c.fymax_r * (3.0 / P.C_ar);
  c.c2_f = (P.C_af * P.C_af / 3.0) * rf;
  c.c2_r = (P.C_ar * P.C_ar / 3.0) * rr;
  c.c3_f = (P.C_af * P.C_af * P.C_af / 27.0) * (rf * rf);
  c.c3_r = (P.C_ar * P.C_ar * P.C_ar / 27.0) * (rr * rr);
  return c;
}

// tire_consts_fast with the car-only factors taken from CarDerived (constant bank)
MPOPIS_HD TireConsts tire_consts_der(const CarParams &P, const CarDerived &D, double accel, double bk, double split,
                                     double sgnVx) {
  TireConsts c;
  const double fx = accel + bk * sgnVx;  // CAR:310-312
  c.fxf = split * fx;
  c.fxr = (1 - split) * fx;
  const double fzf = (D.wf - P.h_cm * fx) * D.inv_L;
  const double fzr = (D.wr + P.h_cm * fx) * D.inv_L;
  const double vf = fmax((P.mu_f * fzf) * (P.mu_f * fzf) - c.fxf * c.fxf, 1e-8);
  const double vr = fmax((P.mu_r * fzr) * (P.mu_r * fzr) - c.fxr * c.fxr, 1e-8);
  const double rf = rsqrt_pos(vf), rr = rsqrt_pos(vr);  // vf, vr >= 1e-8: no special cases, no branch
  c.fymax_f = vf * rf;
  c.fymax_r = vr * rr;
  c.thr_f = c.fymax_f * D.thrC_f;
  c.thr_r = c.fymax_r * D.thrC_r;
  c.c2_f = D.c2C_f * rf;
  c.c2_r = D.c2C_r * rr;
  c.c3_f = D.c3C_f * (rf * rf);
  c.c3_r = D.c3C_r * (rr * rr);
  return c;
}

// brush-tyre lateral force from tan α = num/den with |α| < π (fast paths, Vx > 0)
template <int MODE>
MPOPIS_HD double tire_fy_ratio(double num, double den, double C, double c2, double c3,
                                                double thr, double fymax) {
  const double ta = MODE == 0 ? fast_div(num, den) : num / den;
  const double cubic = -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  const double sat = -fymax * jl_sign(num);
  return (den > 0.0 && fabs(ta) < thr) ? cubic : sat;
}

// literal calc_tire_fy, CAR:252-260
MPOPIS_HD double tire_fy_literal(double alpha, double C, double c2, double c3,
                                                  double thr, double fymax) {
  const double ta = tan(alpha);
  if (fabs(alpha) < atan(thr)) return -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  return -fymax * jl_sign(alpha);
}

// _step!(env::CarRacingEnv, a), CAR:282-344. s = [x, y, Ψ, Vx, Vy, Ψ̇, δ, pedal].
MPOPIS_HD void car_step_fast(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                              double a0, double a1);
MPOPIS_HD bool car_step_spec(const CarParams &P, const CarDerived &D, double dt, double ddt, int nsub,
                             const double *s, double *o, double a0, double a1, double *trig, bool resync);

template <int MODE>
MPOPIS_HD void car_step(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                         double a0, double a1) {
  if constexpr (MODE == 0) {
    car_step_fast(P, dt, ddt, nsub, s, a0, a1);
    return;
  }
  if constexpr (MODE == 3) {
    double o[8];
    const CarDerived D = derive_car(P, ddt);  // env_step / host checker; the rollout kernel passes env.der[c]
    double trig[4];
    if (car_step_spec(P, D, dt, ddt, nsub, s, o, a0, a1, trig, true)) {
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] = o[q];
    } else {
      car_step_fast(P, dt, ddt, nsub, s, a0, a1);
    }
    return;
  }
  constexpr bool FAST = MODE != 1;
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
    const double sg_now = jl_sign(Vx);
    if (sg_now != sg) {  // sign(Vx) flipped: brake force changes direction (rare)
      sg = sg_now;
      if (bk != 0.0) tc = tire_consts(P, accel, bk, split, sg);
    }
    double sd, cd;
    if (MODE == 0 && fabs(delta) <= 0.8) sincos_kernel(delta, &sd, &cd);
    else sincos(delta, &sd, &cd);
    const double yf = Vy + P.l_f * psid, yr = Vy - P.l_r * psid;
    double fyf, fyr;
    if (FAST && Vx > 0.0) {
      // tan(atan(yf, Vx) − δ) = (yf cδ − Vx sδ)/(Vx cδ + yf sδ); rear: tan α_r = yr / Vx
      fyf = tire_fy_ratio<MODE>(yf * cd - Vx * sd, Vx * cd + yf * sd, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f,
                                tc.fymax_f);
      fyr = tire_fy_ratio<MODE>(yr, Vx, P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
    } else {
      const double a_f = atan2(yf, Vx) - delta;  // CAR:304
      const double a_r = atan2(yr, Vx);          // CAR:305
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
        const double k = rint(psi * 0.15915494309189535);
        psi = fma(-k, 6.283185307179586, psi);
        psi = fma(-k, 2.4492935982947064e-16, psi);
      }
      if (MODE == 0) sincos_pi(psi, &sp, &cp);
      else sincos(psi, &sp, &cp);
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

// Short sin/cos for a small rotation increment (|x| <= 0.03: truncation error < 1e-18 relative)
MPOPIS_HD void sincos_tiny(double x, double *s, double *c) {
  const double z = x * x;
  double ps = fma(z, -1.9841269841269841e-04, 8.3333333333333332e-03);
  ps = fma(z, ps, -1.6666666666666666e-01);
  *s = fma(x * z, ps, x);
  const double pc = fma(z, -1.3888888888888889e-03, 4.1666666666666664e-02);
  *c = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// MODE 0 ("fast v3") implementation of _step! (CAR:282-344). On top of the v1/v2 reformulations:
//   * sin/cos of δ by the angle-addition recurrence (δ advances by the constant rate·δt inside a control
//     step and never overshoots its target, so |δ| <= max(|δ₀|, |a₁ δ_max|));
//   * sin/cos of Ψ by rotating (sin Ψ, cos Ψ) with the per-sub-step increment Ψ̇·δt; both recurrences are
//     re-synchronised with a full evaluation at every control step, so drift is bounded by nsub roundings;
//     Ψ itself is accumulated and wrapped once per step (the wrap is the identity modulo 2π);
//   * forward motion (Vx > 0, the case for every realistic rollout) needs no sign bookkeeping and shares
//     ONE reciprocal between the front and rear slip ratios; anything else takes the general path.
MPOPIS_HD void car_step_fast(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                              double a0, double a1) {
  double x = s[0], y = s[1], psi = s[2], Vx = s[3], Vy = s[4], psid = s[5], delta = s[6];
  const double tgt = a0 * P.d_max - delta;
  const double rate = fmin(fabs(tgt) / dt, P.dd_max) * jl_sign(tgt);  // CAR:295-296
  const double pedal = a1;                                           // CAR:297
  const double accel = P.Fx_max * fmax(pedal, 0.0);                  // CAR:310
  const double bk = P.Fx_min * fmin(pedal, 0.0);                     // CAR:311 without sign(Vx)
  const double split = pedal <= 0.0 ? P.l_brake : P.l_drive;
  const double inv_Izz = 1 / P.Izz, inv_m = 1 / P.m;
  double sg = jl_sign(Vx);
  TireConsts tc = tire_consts_fast(P, accel, bk, split, sg);
  const double dlt = rate * ddt;
  const bool small = fmax(fabs(delta), fabs(a0 * P.d_max)) <= 0.78;
  double sd, cd, sdl = 0.0, cdl = 1.0;
  if (small) {
    sincos_kernel(delta, &sd, &cd);
    sincos_kernel(dlt, &sdl, &cdl);
  }
  if (fabs(psi) > CUDART_PI) {  // callers may hand in any heading; CAR:330 keeps it in (−π, π] afterwards
    const double k = rint(psi * 0.15915494309189535);
    psi = fma(-k, 6.283185307179586, psi);
    psi = fma(-k, 2.4492935982947064e-16, psi);
  }
  double sp, cp;
  sincos_pi(psi, &sp, &cp);
  // ncu: the dominant stall is "wait" (fixed-latency FP64 dependencies) at ~3 warps per scheduler; unrolling
  // by two lets the scheduler overlap the position/heading tail of one sub-step with the next tyre chain.
#pragma unroll 2
  for (int i = 0; i < nsub; ++i) {
    delta += dlt;  // CAR:301
    if (small) {
      const double ns = fma(sd, cdl, cd * sdl);
      cd = fma(cd, cdl, -(sd * sdl));
      sd = ns;
    } else {
      sincos(delta, &sd, &cd);
    }
    const double yf = Vy + P.l_f * psid, yr = Vy - P.l_r * psid;
    double fyf, fyr, fx_aero;
    if (Vx > 0.0 && sg > 0.0) {
      // tan(atan(yf, Vx) − δ) = num/den, tan α_r = yr / Vx (CAR:304-305 without atan/tan)
      const double num = yf * cd - Vx * sd, den = Vx * cd + yf * sd;
      double ta_r;
      if (den > 0.0) {
        const double dv = den * Vx;
        double r = rcp_seed(dv);
        const double e = fma(-dv, r, 1.0);
        r = fma(r, fma(e, e, e), r);                             // r(1 + e + e²): error e³ = 2^-69
        const double ta = (num * Vx) * r;
        ta_r = (yr * den) * r;
        // −C t + c2 |t| t − c3 t³ = t·(−C + |t|·(c2 − c3 |t|))   (CAR:256, Horner form)
        const double at = fabs(ta);
        const double cubic = ta * fma(at, fma(-tc.c3_f, at, tc.c2_f), -P.C_af);
        fyf = at < tc.thr_f ? cubic : copysign(tc.fymax_f, -num);  // CAR:255-259
      } else {  // |α_f| >= 90°: saturated
        fyf = copysign(tc.fymax_f, -num);
        ta_r = fast_div(yr, Vx);
      }
      const double atr = fabs(ta_r);
      const double cubic_r = ta_r * fma(atr, fma(-tc.c3_r, atr, tc.c2_r), -P.C_ar);
      fyr = atr < tc.thr_r ? cubic_r : copysign(tc.fymax_r, -yr);
      fx_aero = P.C_D0 + P.C_D1 * Vx;  // CAR:308 with sign(Vx) = 1
    } else {
      const double sg_now = jl_sign(Vx);
      if (sg_now != sg) {  // sign(Vx) flipped: brake force changes direction
        sg = sg_now;
        if (bk != 0.0) tc = tire_consts_fast(P, accel, bk, split, sg);
      }
      if (Vx > 0.0) {
        fyf = tire_fy_ratio<0>(yf * cd - Vx * sd, Vx * cd + yf * sd, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f, tc.fymax_f);
        fyr = tire_fy_ratio<0>(yr, Vx, P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
      } else {  // reversing / standstill: the un-wrapped slip angle matters, keep the libm sequence
        fyf = tire_fy_literal(atan2(yf, Vx) - delta, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f, tc.fymax_f);
        fyr = tire_fy_literal(atan2(yr, Vx), P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
      }
      fx_aero = (P.C_D0 + P.C_D1 * fabs(Vx)) * sg;
    }
    const double psidd = inv_Izz * (P.l_f * (tc.fxf * sd + fyf * cd) - P.l_r * fyr);        // CAR:322
    const double Vy_dot = inv_m * (fyf * cd + tc.fxf * sd + fyr) - psid * Vx;               // CAR:323
    const double Vx_dot = inv_m * (tc.fxf * cd - fyf * sd + tc.fxr - fx_aero) + psid * Vy;  // CAR:324
    psid += psidd * ddt;  // CAR:326
    Vx += Vx_dot * ddt;   // CAR:327
    Vy += Vy_dot * ddt;   // CAR:328
    const double dpsi = psid * ddt;
    psi += dpsi;  // CAR:329 (wrapped once per step below)
    double sdp, cdp;
    if (fabs(dpsi) <= 0.03) sincos_tiny(dpsi, &sdp, &cdp);
    else if (fabs(dpsi) <= 0.8) sincos_kernel(dpsi, &sdp, &cdp);  // spinning car: still no libdevice call
    else sincos(dpsi, &sdp, &cdp);
    const double nsp = fma(sp, cdp, cp * sdp);
    cp = fma(cp, cdp, -(sp * sdp));
    sp = nsp;
    x += (Vx * cp - Vy * sp) * ddt;  // CAR:331
    y += (Vx * sp + Vy * cp) * ddt;  // CAR:332
  }
  if (fabs(psi) > CUDART_PI) {  // CAR:330
    const double k = rint(psi * 0.15915494309189535);
    psi = fma(-k, 6.283185307179586, psi);
    psi = fma(-k, 2.4492935982947064e-16, psi);
  }
  s[0] = x, s[1] = y, s[2] = psi, s[3] = Vx, s[4] = Vy, s[5] = psid, s[6] = delta, s[7] = pedal;
}

// MODE 3 ("fast v4"): one control step integrated SPECULATIVELY as straight-line code, repaired by v3 when invalid.
//
// What the profiles said (profiles/README.md, round 1): v3 executes ≈10 BRA and 6 BSSY/BSYNC pairs per sub-step, yet
// removing them alone changed nothing (305 vs 297 µs): at K = 65 536 the four SM sub-partitions hold {4,4,3,3} warps
// and the loaded ones run the FP64 pipe at ≈80-90 % — the kernel is bound by the NUMBER OF FP64-PIPE INSTRUCTIONS
// (DFMA/DMUL/DADD and DSETP alike, 2 issue cycles each). So this variant spends its effort there:
//   * the brush-tyre force is written for every Vx != 0 (not only Vx > 0): tan α is the rotated ratio in any
//     quadrant (tan has period π); |α| < atan(thr) <=> cos α > 0 and |tan α| < thr, with cos α_f ∝ den,
//     cos α_r ∝ Vx (|δ| <= 0.78 keeps the un-wrapped α_f inside (−3π/2, 3π/2)); in saturation sign(α) is
//     sign(sin α) for Vx > 0 and the sign bit of atan2's numerator for Vx < 0 (atan2 ∈ ±(π/2, π]). On the bench
//     workload up to 10 % of the rollouts slide backwards at some point; v3 sends their whole warp through the
//     libm atan2/tan sequence, here they stay on the straight-line path;
//   * sign tests (Vx > 0, den > 0), the validity tests (den·Vx normal, |Ψ̇ δt| small) and the ±fy_max selection
//     are integer operations on the high words (ALU pipe) instead of DSETP / DADD-negations (FP64 pipe);
//   * the Euler updates are folded: Ψ̇ += c₁A − c₂F_yr, Vy += c_m(A + F_yr) − (Ψ̇δt)Vx, Vx = k_x Vx + c_m(B + F_xr ∓ C_D0)
//     + (Ψ̇δt)Vy with A = F_yf cos δ + F_xf sin δ, B = F_xf cos δ − F_yf sin δ, k_x = 1 − c_m C_D1 (the aero term's
//     linear part), re-using the previous sub-step's Ψ̇δt: 60 FP64-pipe instructions per sub-step instead of 75.
//     This re-associates CAR:322-328 (differences of a few ulp per sub-step; tests/test_host_step.py bounds the
//     deviation from the literal step);
//   * conditions under which this is NOT the reference's arithmetic — den·Vx zero/denormal, a sign change of Vx
//     while braking (the tyre constants hold sign(Vx), CAR:311), a yaw increment beyond the short sincos polynomial,
//     |δ| > 0.78 — are accumulated in a predicate; the caller then discards the speculative state and integrates that
//     control step again with car_step_fast (v3, all branches).
// Returns true when the speculative result `o` is valid.
MPOPIS_HD int hi32(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  long long b;
  memcpy(&b, &x, sizeof b);
  return (int)(b >> 32);
#endif
}
// magnitude of `mag` (>= 0) with the sign opposite to the sign bit carried by the high word `shi`
MPOPIS_HD double with_opposite_sign(double mag, int shi) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(mag) | (~shi & (int)0x80000000), __double2loint(mag));
#else
  return (shi < 0) ? mag : -mag;
#endif
}

// `trig` = {sin δ, cos δ, sin Ψ, cos Ψ} carried from one control step to the next by the rollout kernel: with
// resync == false the recurrences simply continue (their drift is one rounding per sub-step; the kernel
// re-evaluates them from δ and Ψ every few control steps and after every repaired step).
MPOPIS_HD bool car_step_spec(const CarParams &P, const CarDerived &D, double dt, double ddt, int nsub,
                             const double *s, double *o, double a0, double a1, double *trig, bool resync) {
  double x = s[0], y = s[1], psi = s[2], Vx = s[3], Vy = s[4], psid = s[5], delta = s[6];
  const double tgt = a0 * P.d_max - delta;
  const double rate = fmin(fast_div(fabs(tgt), dt), P.dd_max) * jl_sign(tgt);  // CAR:295-296
  const double pedal = a1;                                                    // CAR:297
  const double accel = P.Fx_max * fmax(pedal, 0.0);                           // CAR:310
  const double bk = P.Fx_min * fmin(pedal, 0.0);                              // CAR:311 without sign(Vx)
  const double split = pedal <= 0.0 ? P.l_brake : P.l_drive;
  const double sg = jl_sign(Vx);
  const TireConsts tc = tire_consts_der(P, D, accel, bk, split, sg);
  const double dlt = rate * ddt;
  // validity is accumulated in the SIGN BIT of an integer (ALU pipe, no predicates): a sign change of Vx while
  // braking, den·Vx zero/denormal, |Ψ̇ δt| > 0.03 (or NaN)
  const int hvx0 = hi32(Vx), brake_mask = bk != 0.0 ? (int)0x80000000 : 0;
  int bad = 0;
  const bool pre_ok = (fmax(fabs(delta), fabs(a0 * P.d_max)) <= 0.78) & (Vx != 0.0) & (fabs(dlt) <= 0.03);
  double sd, cd, sp, cp, sdl, cdl;
  sincos_tiny(dlt, &sdl, &cdl);  // |rate·δt| <= δ̇_max·δt = 0.0157 for the default car
  if (resync) {
    sincos_kernel(delta, &sd, &cd);
    if (fabs(psi) > CUDART_PI) {
      const double k = rint(psi * 0.15915494309189535);
      psi = fma(-k, 6.283185307179586, psi);
      psi = fma(-k, 2.4492935982947064e-16, psi);
    }
    sincos_pi(psi, &sp, &cp);
  } else {
    sd = trig[0], cd = trig[1], sp = trig[2], cp = trig[3];
  }
  double dpsi = psid * ddt;  // Ψ̇·δt of the CURRENT Ψ̇: the −Ψ̇Vx / +Ψ̇Vy terms of this sub-step, the heading step of the last
#pragma unroll 2
  for (int i = 0; i < nsub; ++i) {
    const double ns = fma(sd, cdl, cd * sdl);  // sin/cos(δ + rate·δt), CAR:301
    cd = fma(cd, cdl, -(sd * sdl));
    sd = ns;
    const double yf = fma(P.l_f, psid, Vy), yr = fma(-P.l_r, psid, Vy);
    const double num = fma(yf, cd, -(Vx * sd)), den = fma(Vx, cd, yf * sd);  // tan α_f = num/den, tan α_r = yr/Vx
    const double dv = den * Vx;
    const int hvx = hi32(Vx);
    const bool fwd = hvx >= 0;
    bad |= ((hvx ^ hvx0) & brake_mask) | ((hi32(dv) & 0x7ff00000) - 0x00100000);
    double r = rcp_seed(dv);  // 2^-23 seed
    const double e = fma(-dv, r, 1.0);
    r = fma(r, fma(e, e, e), r);  // r(1 + e + e²): error e³ = 2^-69
    const double ta = (num * Vx) * r, ta_r = (yr * den) * r;
    const double at = fabs(ta), atr = fabs(ta_r);
    const double cubic = ta * fma(at, fma(-tc.c3_f, at, tc.c2_f), -P.C_af);  // CAR:256, Horner form
    const double cubic_r = ta_r * fma(atr, fma(-tc.c3_r, atr, tc.c2_r), -P.C_ar);
    const double fyf = ((hi32(den) >= 0) & (at < tc.thr_f)) ? cubic  // CAR:255-259
                                                             : with_opposite_sign(tc.fymax_f, fwd ? hi32(num) : hi32(yf));
    const double fyr = (fwd & (atr < tc.thr_r)) ? cubic_r : with_opposite_sign(tc.fymax_r, hi32(yr));
    const double A = fma(fyf, cd, tc.fxf * sd), B = fma(-fyf, sd, tc.fxf * cd);
    const double psid_n = fma(D.cI1, A, fma(-D.cI2, fyr, psid));                              // CAR:322,326
    const double Vy_n = fma(D.cm, A + fyr, fma(-dpsi, Vx, Vy));                             // CAR:323,328
    // CAR:308,324,327: −c_m·fx_aero = −c_m C_D1 Vx − copysign(c_m C_D0, Vx)
    const double Vx_n = fma(D.cm, B + tc.fxr, fma(dpsi, Vy, fma(Vx, D.kx, with_opposite_sign(D.cmCD0, hvx))));
    psid = psid_n, Vx = Vx_n, Vy = Vy_n;
    dpsi = psid * ddt;
    psi += dpsi;  // CAR:329 (wrapped once per step below)
    bad |= 0x3F9EB851 - (hi32(dpsi) & 0x7fffffff);  // |Ψ̇ δt| > 0.03 or NaN
    double sdp, cdp;
    sincos_tiny(dpsi, &sdp, &cdp);
    const double nsp = fma(sp, cdp, cp * sdp);
    cp = fma(cp, cdp, -(sp * sdp));
    sp = nsp;
    x = fma(fma(Vx, cp, -(Vy * sp)), ddt, x);  // CAR:331
    y = fma(fma(Vx, sp, Vy * cp), ddt, y);     // CAR:332
  }
  delta = fma((double)nsub, dlt, delta);  // CAR:301 summed
  if (fabs(psi) > CUDART_PI) {  // CAR:330
    const double k = rint(psi * 0.15915494309189535);
    psi = fma(-k, 6.283185307179586, psi);
    psi = fma(-k, 2.4492935982947064e-16, psi);
  }
  o[0] = x, o[1] = y, o[2] = psi, o[3] = Vx, o[4] = Vy, o[5] = psid, o[6] = delta, o[7] = pedal;
  trig[0] = sd, trig[1] = cd, trig[2] = sp, trig[3] = cp;
  return pre_ok & (bad >= 0);
}

// reward(env::CarRacingEnv), CAR:201-213
template <int MODE>
MPOPIS_HD double car_reward(const CarParams &P, double cos_bl, const TrackView &tr,
                                             const double *s) {
  double dist;
  const bool within = within_track<MODE == 0 || MODE == 3, MODE == 3>(tr, s[0], s[1], nullptr, nullptr, &dist);
  const double speed = MODE == 3 ? sqrt_fast(s[3] * s[3] + s[4] * s[4]) : sqrt(s[3] * s[3] + s[4] * s[4]);
  const bool exceed = MODE != 1 ? (s[3] < cos_bl * speed) : (fabs(atan2(s[4], s[3])) > P.b_limit);  // CAR:181-189
  double rew = 0.0;
  if (!within) rew += -1000000.0;
  if (exceed) rew += -5000.0;
  rew += -dist;
  rew += 2.0 * speed;
  return rew;
}

}  // namespace mpopis
