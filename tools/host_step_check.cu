// host_step_check.cu — CPU checker for the device arithmetic of the CarRacing step (csrc/car_model.cuh).
//
// The step functions are __host__ __device__, so this program runs EXACTLY the code the rollout kernels run
// (MUFU seeds replaced by float-rounded reciprocals) on the host and compares, from identical states,
//   MODE 1 "literal"  (the reference's libm call sequence, CAR:282-344)     — the yardstick
//   MODE 0 "fast v3"  (branchy fast formulation)
//   MODE 3 "fast v4"  (speculative straight-line step + v3 repair)
// over random states that include sliding / reversing cars (Vx < 0), saturated tyres and steering at the limits.
// It prints one JSON line: max relative state error per variant after one control step, and the fraction of
// control steps on which the speculative step had to be repaired. Test infrastructure (tests/test_host_step.py);
// nothing in the product links it.
//
//   nvcc -O2 -std=c++17 -Xcompiler -ffp-contract=off -I mpopis_b200/csrc tools/host_step_check.cu -o /tmp/hsc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "car_model.cuh"

using namespace mpopis;

static CarParams default_params() {  // CAR:68-93
  CarParams P;
  P.m = 2000.0, P.Izz = 3764.0, P.h_cm = 0.3, P.l_f = 1.53, P.l_r = 1.23, P.C_D0 = 241.0, P.C_D1 = 25.1;
  P.C_af = 150000.0, P.C_ar = 280000.0, P.mu_f = 0.9, P.mu_r = 0.9;
  P.d_max = 18.0 * M_PI / 180.0, P.dd_max = 90.0 * M_PI / 180.0, P.Fx_max = 7200.0, P.Fx_min = 22500.0;
  P.l_brake = 0.6, P.l_drive = 0.0, P.b_limit = 45.0 * M_PI / 180.0;
  return P;
}

static double rel_err(const double *a, const double *b) {
  double e = 0.0;
  for (int q = 0; q < 8; ++q) {
    double d = fabs(a[q] - b[q]);
    if (q == 2) {  // heading: compare modulo 2π
      d = fabs(remainder(a[q] - b[q], 2.0 * M_PI));
    }
    const double r = d / fmax(1.0, fabs(b[q]));
    if (!(r <= e)) e = r;  // NaN propagates
  }
  return e;
}

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 400000;
  const unsigned seed = argc > 2 ? (unsigned)atol(argv[2]) : 1u;
  const CarParams P = default_params();
  const double dt = 0.1, ddt = 0.01;
  const int nsub = 10;
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> N01(0.0, 1.0);
  double e0 = 0.0, e3 = 0.0, e03 = 0.0;
  long repaired = 0, reversed = 0, repaired_fwd = 0, nan3 = 0;
  double worst3[8] = {0}, worst_ref[8] = {0}, worst_in[10] = {0};
  for (long i = 0; i < n; ++i) {
    double s[8];
    s[0] = 200.0 * (U(g) - 0.5), s[1] = 200.0 * (U(g) - 0.5);
    s[2] = (2.0 * U(g) - 1.0) * M_PI;
    const double u = U(g);
    if (u < 0.55) s[3] = 3.0 + 30.0 * U(g);            // normal driving
    else if (u < 0.75) s[3] = -(0.5 + 20.0 * U(g));    // sliding backwards / reversing
    else if (u < 0.9) s[3] = (U(g) - 0.5) * 2.0;       // about to cross Vx = 0
    else s[3] = 0.3 + U(g);                            // crawling
    s[4] = (U(g) < 0.7 ? 1.0 : 8.0) * N01(g);
    s[5] = (U(g) < 0.8 ? 0.3 : 2.0) * N01(g);
    s[6] = (2.0 * U(g) - 1.0) * P.d_max * (U(g) < 0.9 ? 1.0 : 1.2);
    s[7] = 2.0 * U(g) - 1.0;
    double a0 = 2.0 * U(g) - 1.0, a1 = 2.0 * U(g) - 1.0;
    if (U(g) < 0.15) a0 = a0 > 0 ? 1.0 : -1.0;  // steering at the limit
    if (U(g) < 0.15) a1 = a1 > 0 ? 1.0 : -1.0;  // full pedal / full brake
    if (U(g) < 0.1) a1 = 0.0;
    double lit[8], v3[8], v4[8], o[8];
    memcpy(lit, s, sizeof s), memcpy(v3, s, sizeof s), memcpy(v4, s, sizeof s);
    car_step<1>(P, dt, ddt, nsub, lit, a0, a1);
    car_step<0>(P, dt, ddt, nsub, v3, a0, a1);
    double trig1[4];
    const bool ok = car_step_spec(P, derive_car(P, ddt), dt, ddt, nsub, s, o, a0, a1, trig1, true);
    car_step<3>(P, dt, ddt, nsub, v4, a0, a1);
    if (!ok) ++repaired;
    if (!ok && s[3] > 2.0) ++repaired_fwd;
    if (s[3] < 0.0) ++reversed;
    // Vx within a hair of 0 during the step: atan2's branch is decided by rounding — not comparable
    const bool sane = fabs(lit[3]) > 1e-3 || fabs(s[3]) > 1.0;
    if (!sane) continue;
    const double r0 = rel_err(v3, lit), r3 = rel_err(v4, lit), r03 = rel_err(v4, v3);
    if (!(r3 == r3)) ++nan3;
    if (!(r0 <= e0)) e0 = r0;
    if (!(r3 <= e3)) {
      e3 = r3;
      memcpy(worst3, v4, sizeof v4), memcpy(worst_ref, lit, sizeof lit), memcpy(worst_in, s, sizeof s);
      worst_in[8] = a0, worst_in[9] = a1;
    }
    if (!(r03 <= e03)) e03 = r03;
  }
  // sequences of 25 control steps the way the rollout kernel runs them: sin/cos of δ and Ψ carried between steps,
  // re-evaluated every 5th step and after a repaired step; compared with the literal step from the same start
  double eseq = 0.0;
  long seq_repairs = 0;
  const long nseq = n / 50;
  const CarDerived D = derive_car(P, ddt);
  for (long i = 0; i < nseq; ++i) {
    double s4[8] = {200.0 * (U(g) - 0.5), 200.0 * (U(g) - 0.5), (2.0 * U(g) - 1.0) * M_PI, 8.0 + 25.0 * U(g),
                    N01(g), 0.2 * N01(g), (2.0 * U(g) - 1.0) * 0.3, 0.0};
    double lit[8], trig[4], o[8];
    memcpy(lit, s4, sizeof s4);
    bool valid = false;
    double a0 = 0.0, a1 = 0.3;
    for (int t = 0; t < 25; ++t) {
      a0 = fmin(fmax(a0 + 0.3 * N01(g), -1.0), 1.0), a1 = fmin(fmax(0.3 + 0.4 * N01(g), -1.0), 1.0);
      car_step<1>(P, dt, ddt, nsub, lit, a0, a1);
      const bool resync = !valid || (t % 5) == 0;
      if (car_step_spec(P, D, dt, ddt, nsub, s4, o, a0, a1, trig, resync)) {
        memcpy(s4, o, sizeof o), valid = true;
      } else {
        car_step_fast(P, dt, ddt, nsub, s4, a0, a1), valid = false, ++seq_repairs;
      }
    }
    const double r = rel_err(s4, lit);
    if (!(r <= eseq)) eseq = r;
  }
  printf("{\"n\": %ld, \"max_rel_err_seq25_v4_vs_literal\": %.3e, \"seq_repairs\": %ld, \"max_rel_err_v3_vs_literal\": %.3e, \"max_rel_err_v4_vs_literal\": %.3e, "
         "\"max_rel_err_v4_vs_v3\": %.3e, \"repaired_frac\": %.5f, \"repaired_forward_frac\": %.6f, "
         "\"reversed_frac\": %.4f, \"nan_v4\": %ld}\n",
         n, eseq, seq_repairs, e0, e3, e03, (double)repaired / n, (double)repaired_fwd / n, (double)reversed / n, nan3);
  if (argc > 3) {
    printf("worst in : ");
    for (int q = 0; q < 10; ++q) printf("%.17g ", worst_in[q]);
    printf("\nworst v4 : ");
    for (int q = 0; q < 8; ++q) printf("%.17g ", worst3[q]);
    printf("\nworst lit: ");
    for (int q = 0; q < 8; ++q) printf("%.17g ", worst_ref[q]);
    printf("\n");
  }
  return 0;
}
