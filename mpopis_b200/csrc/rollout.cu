// rollout.cu — G2: the K-trajectory rollout + running-cost kernel (the dominant kernel).
//
// Replaces simulate_model (POL:261-278) -> rollout_model (UTL:129-144) -> _step! (CAR:282-344) +
// reward (CAR:201-213 / MCR:145-158) + within_track (TRK:68-92), and the MountainCar step/reward
// (RLEnvs, EXM:4-22). One thread integrates one rollout: T control steps x nsub Euler sub-steps,
// all FP64. The kernel is bound by FP64 instruction issue (≈16 B of HBM traffic per rollout-step
// against thousands of FP64 instructions), so the work goes into issuing fewer
// instructions for the same mathematics. Three variants are kept so every step is A/B-measurable
// and parity-checked against the oracle (tests/test_gpu_parity.py, profiles/):
//
//   MODE 1 "literal": the reference's libm call sequence (atan2, tan, atan, sin, cos per sub-step).
//   MODE 2 "fast v1": (round-1 first cut) algebraic reformulation, libdevice sincos and IEEE division
//     * slip angles: tan(atan2(y,x) − δ) is a rotated ratio (no atan2/tan/atan); the saturation test
//       |α| < atan(3 fy_max/C) becomes |tan α| < 3 fy_max/C where |α| < π/2; the exact libm sequence is
//       kept for Vx <= 0 (car reversing), where the un-wrapped angle matters for sign(α);
//     * tyre-force constants depend only on (pedal, sign Vx): hoisted out of the sub-step loop;
//     * heading wrap atan(sin Ψ, cos Ψ) is a conditional ±2π (identity on (−π, π]);
//     * the β test |atan(Vy, Vx)| > β_limit is Vx < cos(β_limit)·‖V‖.
//   MODE 0 "fast v2" (default): v1 plus
//     * sincos from fdlibm-style minimax kernels whose coefficients sit in the constant bank, so each
//       DFMA reads them as c[bank][off] operands (ncu showed 15 % of issue slots were UMOVs
//       re-materialising libdevice's FP64 immediates) and there is no slow-path branch;
//     * division as MUFU reciprocal + 2 Newton steps + residual correction, branch-free;
//     * nearest-track-point search pruned by an exact spatial look-up table built on the host
//       (mpopis_b200.cu: build_track_lut): per 2 m cell the list of points that can be the arg-min for
//       any position in the cell; candidates are evaluated in index order with the same un-fused
//       arithmetic, so the selected indices are bit-identical to the full scan (which remains the
//       fallback outside the table).
// The integer decisions of within_track (first arg-min, neighbour choice) use explicitly
// non-contracted arithmetic (__dmul_rn/__dadd_rn) in every mode so that, on identical inputs, the
// indices equal the reference's Float64 evaluation bit for bit.
#include <math_constants.h>

#include "engine.cuh"

namespace mpopis {

// fdlibm __kernel_sin / __kernel_cos minimax coefficients (|x| <= π/4, error < 2^-57)
__constant__ double kSinCos[12] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};

__device__ __forceinline__ double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
__device__ __forceinline__ double clamp1(double v) { return fmin(fmax(v, -1.0), 1.0); }

// sin/cos on |x| <= ~π/4 (no range reduction)
__device__ __forceinline__ void sincos_kernel(double x, double *s, double *c) {
  const double z = x * x;
  double ps = fma(z, kSinCos[5], kSinCos[4]);
  ps = fma(z, ps, kSinCos[3]);
  ps = fma(z, ps, kSinCos[2]);
  ps = fma(z, ps, kSinCos[1]);
  ps = fma(z, ps, kSinCos[0]);
  *s = fma(x * z, ps, x);
  double pc = fma(z, kSinCos[11], kSinCos[10]);
  pc = fma(z, pc, kSinCos[9]);
  pc = fma(z, pc, kSinCos[8]);
  pc = fma(z, pc, kSinCos[7]);
  pc = fma(z, pc, kSinCos[6]);
  *c = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// sin/cos for |x| <= π(1+ε): quadrant reduction with a two-term π/2, then the kernels
__device__ __forceinline__ void sincos_pi(double x, double *s, double *c) {
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
__device__ __forceinline__ double fast_div(double n, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
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
template <bool USE_LUT>
__device__ __forceinline__ bool within_track(const TrackView &tr, double px, double py, int *idx_out,
                                             int *idx2_out, double *dist_out) {
  int mi = 0;
  double best = CUDART_INF;
  bool done = false;
  if (USE_LUT && tr.lut) {
    const double fx = (px - tr.x0) * tr.inv_c, fy = (py - tr.y0) * tr.inv_c;
    if (fx >= 0.0 && fy >= 0.0 && fx < (double)tr.nx && fy < (double)tr.ny) {
      const uint4 cell = __ldg(tr.lut + (int)fy * tr.nx + (int)fx);
      const unsigned cnt = cell.x & 0xffffu;
      if (cnt <= 7u) {
        const unsigned long long w0 = ((unsigned long long)cell.y << 32) | cell.x;
        const unsigned long long w1 = ((unsigned long long)cell.w << 32) | cell.z;
        for (unsigned q = 0; q < cnt; ++q) {  // candidates are stored in ascending index order
          const int i = (int)(((q < 3u) ? (w0 >> (16u * q + 16u)) : (w1 >> (16u * (q - 3u)))) & 0xffffu);
          const double dx = tr.x[i] - px, dy = tr.y[i] - py;
          const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
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
      const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      if (d < best) best = d, mi = i;
    }
  }
  const int m1 = mi == 0 ? tr.n - 1 : mi - 1, p1 = mi == tr.n - 1 ? 0 : mi + 1;  // mod1, TRK:75-76
  const double ax = tr.x[m1] - px, ay = tr.y[m1] - py, bx = tr.x[p1] - px, by = tr.y[p1] - py;
  // TRK:77-79 compares norms: sqrt is monotone and correctly rounded, so the squared distances decide
  // unless they are within a few ulps of each other — only then are the two square roots taken.
  const double qa = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
  const double qb = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
  int m2;
  if (USE_LUT && fabs(qa - qb) > 1e-14 * fmax(qa, qb)) m2 = qa < qb ? m1 : p1;
  else m2 = __dsqrt_rn(qa) <= __dsqrt_rn(qb) ? m1 : p1;
  const double p1x = tr.x[mi], p1y = tr.y[mi];
  const double vx = tr.x[m2] - p1x, vy = tr.y[m2] - p1y, ux = px - p1x, uy = py - p1y;
  const double t = (ux * vx + uy * vy) / (vx * vx + vy * vy);  // TRK:87 (projection on the infinite line)
  const double ex = p1x + t * vx - px, ey = p1y + t * vy - py; // TRK:88
  const double dist = sqrt(ex * ex + ey * ey);                 // TRK:89
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
__device__ __forceinline__ TireConsts tire_consts_fast(const CarParams &P, double accel, double bk,
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
  const double rf = rsqrt(vf), rr = rsqrt(vr);
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

// brush-tyre lateral force from tan α = num/den with |α| < π (fast paths, Vx > 0)
template <int MODE>
__device__ __forceinline__ double tire_fy_ratio(double num, double den, double C, double c2, double c3,
                                                double thr, double fymax) {
  const double ta = MODE == 0 ? fast_div(num, den) : num / den;
  const double cubic = -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  const double sat = -fymax * jl_sign(num);
  return (den > 0.0 && fabs(ta) < thr) ? cubic : sat;
}

// literal calc_tire_fy, CAR:252-260
__device__ __forceinline__ double tire_fy_literal(double alpha, double C, double c2, double c3,
                                                  double thr, double fymax) {
  const double ta = tan(alpha);
  if (fabs(alpha) < atan(thr)) return -C * ta + c2 * fabs(ta) * ta - c3 * (ta * ta * ta);
  return -fymax * jl_sign(alpha);
}

// _step!(env::CarRacingEnv, a), CAR:282-344. s = [x, y, Ψ, Vx, Vy, Ψ̇, δ, pedal].
__device__ __forceinline__ void car_step_fast(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                              double a0, double a1);

template <int MODE>
__device__ __forceinline__ void car_step(const CarParams &P, double dt, double ddt, int nsub, double *s,
                                         double a0, double a1) {
  if constexpr (MODE == 0) {
    car_step_fast(P, dt, ddt, nsub, s, a0, a1);
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
__device__ __forceinline__ void sincos_tiny(double x, double *s, double *c) {
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
__device__ __forceinline__ void car_step_fast(const CarParams &P, double dt, double ddt, int nsub, double *s,
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
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(dv));  // 2^-23 seed
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

// reward(env::CarRacingEnv), CAR:201-213
template <int MODE>
__device__ __forceinline__ double car_reward(const CarParams &P, double cos_bl, const TrackView &tr,
                                             const double *s) {
  double dist;
  const bool within = within_track<MODE == 0>(tr, s[0], s[1], nullptr, nullptr, &dist);
  const double speed = sqrt(s[3] * s[3] + s[4] * s[4]);
  const bool exceed = MODE != 1 ? (s[3] < cos_bl * speed) : (fabs(atan2(s[4], s[3])) > P.b_limit);  // CAR:181-189
  double rew = 0.0;
  if (!within) rew += -1000000.0;
  if (exceed) rew += -5000.0;
  rew += -dist;
  rew += 2.0 * speed;
  return rew;
}

// (env)(a) + reward(env) for 1..N cars: CAR:238-241 / MCR:200-207, MCR:145-158
template <int NCARS, int MODE>
__device__ __forceinline__ double cars_step_reward(const CarEnvArgs &env, const TrackView &tr, double *s,
                                                   const double *a) {
#pragma unroll
  for (int c = 0; c < NCARS; ++c)
    car_step<MODE>(env.car[c], env.dt, env.ddt, env.nsub, s + 8 * c, a[2 * c], a[2 * c + 1]);
  double rew = 0.0;
#pragma unroll
  for (int c = 0; c < NCARS; ++c) {
    rew += car_reward<MODE>(env.car[c], env.cos_blimit[c], tr, s + 8 * c);
#pragma unroll
    for (int j = c + 1; j < NCARS; ++j) {
      const double dx = s[8 * j] - s[8 * c], dy = s[8 * j + 1] - s[8 * c + 1];
      const double dd = sqrt(dx * dx + dy * dy);
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
  TrackView tr{smem, smem + env.n_trk, smem + 2 * env.n_trk, env.n_trk,
               env.lut, env.lut_x0, env.lut_y0, env.lut_inv_c, env.lut_nx, env.lut_ny};
  return tr;
}

template <int NCARS, int MODE>
__global__ void __launch_bounds__(128) rollout_car_kernel(const __grid_constant__ CarEnvArgs env,
                                                          const __grid_constant__ RolloutArgs a,
                                                          const int *stop) {
  extern __shared__ double smem[];
  if (stop && *stop) return;
  const TrackView tr = stage_track(env, smem);
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.K) return;
  constexpr int AS = 2 * NCARS, SS = 8 * NCARS;
  double s[SS];
#pragma unroll
  for (int q = 0; q < SS; ++q) s[q] = __ldg(a.state0 + q);
  const double *Ek = a.E + k;
  double cost = 0.0, cc = 0.0;
  // the noise of step t+1 is fetched while step t integrates (ncu: long-scoreboard stalls on these loads)
  double e_next[AS];
#pragma unroll
  for (int r = 0; r < AS; ++r) e_next[r] = Ek[(size_t)r * a.ldk];
  for (int t = 0; t < a.T; ++t) {
    double act[AS], e_cur[AS];
#pragma unroll
    for (int r = 0; r < AS; ++r) e_cur[r] = e_next[r];
    if (t + 1 < a.T) {
#pragma unroll
      for (int r = 0; r < AS; ++r) e_next[r] = Ek[(size_t)((t + 1) * AS + r) * a.ldk];
    }
#pragma unroll
    for (int r = 0; r < AS; ++r) {
      const int row = t * AS + r;
      const double v = __ldg(a.U + row) + e_cur[r];  // Vₖ = pol.U + E[:,k], POL:271
      if (a.bvec) cc += __ldg(a.bvec + row) * (v - __ldg(a.U_orig + row));  // POL:272
      act[r] = clamp1(v);                                                   // UTL:55-67
    }
    cost -= cars_step_reward<NCARS, MODE>(env, tr, s, act);  // UTL:137-138
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
  const bool done = (x >= e.goal_pos && v >= e.goal_vel) || t >= e.max_steps;
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
    const double val = __ldg(a.U + t) + Ek[(size_t)t * a.ldk];
    if (a.bvec) cc += __ldg(a.bvec + t) * (val - __ldg(a.U_orig + t));
    cost -= mc_step_reward(env, x, v, t_env, clamp1(val), nullptr);
    if (a.traj) {
      a.traj[((size_t)k * 2 + 0) * a.T + t] = x;
      a.traj[((size_t)k * 2 + 1) * a.T + t] = v;
    }
  }
  a.costs[k] = cost + cc;
}

template <int MODE>
static void launch_rollout_car_v(const CarEnvArgs &env, const RolloutArgs &a, int block, const int *stop,
                                 cudaStream_t st) {
  const int grid = (a.K + block - 1) / block;
  const size_t smem = sizeof(double) * 3 * env.n_trk;
#define MPOPIS_LAUNCH(N)                                                                   \
  case N:                                                                                  \
    if (smem > 48 * 1024)                                                                  \
      cudaFuncSetAttribute(rollout_car_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)smem);                                                     \
    rollout_car_kernel<N, MODE><<<grid, block, smem, st>>>(env, a, stop);                 \
    break;
  switch (env.n_cars) {
    MPOPIS_LAUNCH(1)
    MPOPIS_LAUNCH(2)
    MPOPIS_LAUNCH(3)
    MPOPIS_LAUNCH(4)
  }
  if constexpr (MODE == 0) {  // 5..8 cars: only the default variant is instantiated (build time)
    switch (env.n_cars) {
      MPOPIS_LAUNCH(5)
      MPOPIS_LAUNCH(6)
      MPOPIS_LAUNCH(7)
      MPOPIS_LAUNCH(8)
    }
  }
#undef MPOPIS_LAUNCH
}

void launch_rollout_car(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, const int *stop,
                        cudaStream_t st) {
  if (variant == 0 || env.n_cars > 4) launch_rollout_car_v<0>(env, a, block, stop, st);
  else if (variant == 1) launch_rollout_car_v<1>(env, a, block, stop, st);
  else launch_rollout_car_v<2>(env, a, block, stop, st);
}

void launch_rollout_mc(const McEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t st) {
  rollout_mc_kernel<<<(a.K + block - 1) / block, block, 0, st>>>(env, a, stop);
}

// ---- parity surfaces -------------------------------------------------------------------------
template <bool USE_LUT>
__global__ void track_query_kernel(const __grid_constant__ CarEnvArgs env, const double *pos, int n, int *idx,
                                   int *idx2, double *dist, unsigned char *within) {
  extern __shared__ double smem[];
  const TrackView tr = stage_track(env, smem);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a, b;
  double d;
  const bool w = within_track<USE_LUT>(tr, pos[2 * i], pos[2 * i + 1], &a, &b, &d);
  if (idx) idx[i] = a;
  if (idx2) idx2[i] = b;
  if (dist) dist[i] = d;
  if (within) within[i] = w ? 1 : 0;
}

void launch_track_query(const CarEnvArgs &env, const double *pos, int n, int *idx, int *idx2, double *dist,
                        unsigned char *within, int use_lut, cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * env.n_trk;
  if (use_lut) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(track_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    track_query_kernel<true><<<(n + 127) / 128, 128, smem, st>>>(env, pos, n, idx, idx2, dist, within);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(track_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    track_query_kernel<false><<<(n + 127) / 128, 128, smem, st>>>(env, pos, n, idx, idx2, dist, within);
  }
}

template <int MODE>
__global__ void env_step_car_kernel(const __grid_constant__ CarEnvArgs env, double *state,
                                    const double *action, long long *env_t, double *reward) {
  extern __shared__ double smem[];
  const TrackView tr = stage_track(env, smem);
  if (threadIdx.x != 0) return;
  double rew = 0.0;
  for (int c = 0; c < env.n_cars; ++c)
    car_step<MODE>(env.car[c], env.dt, env.ddt, env.nsub, state + 8 * c, action[2 * c], action[2 * c + 1]);
  for (int c = 0; c < env.n_cars; ++c) {
    rew += car_reward<MODE>(env.car[c], env.cos_blimit[c], tr, state + 8 * c);
    for (int j = c + 1; j < env.n_cars; ++j) {
      const double dx = state[8 * j] - state[8 * c], dy = state[8 * j + 1] - state[8 * c + 1];
      const double dd = sqrt(dx * dx + dy * dy);
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
#define MPOPIS_ES(M)                                                                                          \
  {                                                                                                           \
    if (smem > 48 * 1024)                                                                                     \
      cudaFuncSetAttribute(env_step_car_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    env_step_car_kernel<M><<<1, 32, smem, st>>>(env, state, action, env_t, reward);                          \
  }
  if (variant == 0) MPOPIS_ES(0) else if (variant == 1) MPOPIS_ES(1) else MPOPIS_ES(2)
#undef MPOPIS_ES
}

// reward(env) without stepping (CAR:201-213, MCR:145-158, EXM:10-22)
template <int MODE>
__global__ void env_reward_car_kernel(const __grid_constant__ CarEnvArgs env, const double *state, double *reward) {
  extern __shared__ double smem[];
  const TrackView tr = stage_track(env, smem);
  if (threadIdx.x != 0) return;
  double rew = 0.0;
  for (int c = 0; c < env.n_cars; ++c) {
    rew += car_reward<MODE>(env.car[c], env.cos_blimit[c], tr, state + 8 * c);
    for (int j = c + 1; j < env.n_cars; ++j) {
      const double dx = state[8 * j] - state[8 * c], dy = state[8 * j + 1] - state[8 * c + 1];
      const double dd = sqrt(dx * dx + dy * dy);
      rew += -dd;
      if (dd <= 4.0) rew += -11000.0;
    }
  }
  *reward = rew;
}

void launch_env_reward_car(const CarEnvArgs &env, const double *state, double *reward, int variant,
                           cudaStream_t st) {
  const size_t smem = sizeof(double) * 3 * env.n_trk;
#define MPOPIS_ER(M)                                                                                          \
  {                                                                                                           \
    if (smem > 48 * 1024)                                                                                     \
      cudaFuncSetAttribute(env_reward_car_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    env_reward_car_kernel<M><<<1, 32, smem, st>>>(env, state, reward);                                       \
  }
  if (variant == 0) MPOPIS_ER(0) else if (variant == 1) MPOPIS_ER(1) else MPOPIS_ER(2)
#undef MPOPIS_ER
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
  const double rew = mc_step_reward(env, x, v, t, action[0], &d);
  state[0] = x, state[1] = v, *env_t = t;
  if (reward) *reward = rew;
  if (done) *done = d ? 1 : 0;
}

void launch_env_step_mc(const McEnvArgs &env, double *state, const double *action, long long *env_t,
                        double *reward, unsigned char *done, cudaStream_t st) {
  env_step_mc_kernel<<<1, 32, 0, st>>>(env, state, action, env_t, reward, done);
}

}  // namespace mpopis
