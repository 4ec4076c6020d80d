// rollout_split.cu — G2, "v5": the rollout kernel split by DATA DEPENDENCE into two warp roles.
//
// Round 1 left the thread-per-rollout kernel (rollout_kernels.cuh, MODE 3) at ≈52 % of the FP64 issue rate and traced
// the rest to latency: a dependent FP64 instruction issues 23-25 cycles after its producer, K = 65 536 rollouts are
// only 3.46 warps per scheduler, and inside one thread the in-order issue stalls on the recurrence
//     (Vx, Vy, Ψ̇, δ)ᵢ -> slip ratios -> reciprocal -> brush-tyre cubic -> forces -> (Vx, Vy, Ψ̇)ᵢ₊₁       CAR:301-328
// whose ≈15 dependent operations per Euler sub-step ARE the latency of a rollout. Everything else a control step
// does hangs off that chain without feeding it: the pose (Ψ, x, y — CAR:329-332) only consumes (Vx, Vy, Ψ̇δt), and the
// reward (CAR:201-213) with its nearest-point search, projection, IEEE sqrt/div (TRK:68-92) only consumes the pose.
// ptxas cannot overlap them across the sub-step loop / control-step loop in one instruction stream, and at 3.46
// warps per scheduler the hardware has nothing else to issue either. So the two halves become two warps:
//
//   VELOCITY warp (2 per CTA, one rollout per lane): noise load, clamp, control cost (POL:271-272, UTL:55-67), tyre
//       constants, the velocity recurrence; publishes (Vx, Vy, Ψ̇δt) per sub-step into a shared-memory ring.
//   POSE warp (1 per CTA, TWO rollouts per lane — one from each velocity warp, two independent chains in one
//       instruction stream): heading/position integration, heading wrap, within_track + reward, running cost,
//       trajectory log; writes the cost.
//
// 65 536 rollouts are then 2048 latency-bound warps whose chain is shorter (no pose/reward work in it) plus 1024
// throughput-rich warps that fill the FP64 pipe while the former wait: 5.2 warps per scheduler instead of 3.46, with
// the same arithmetic. The ring holds 15 sub-step slots of 3 x 32 doubles per velocity warp, handed over in groups of
// 5 sub-steps with mbarriers (full[3] / empty[3], one elected lane arrives after __syncwarp): both roles run a group
// as ONE straight-line block — the velocity warp with every constant in (uniform) registers and no call or barrier
// inside, the pose warp with the 5 x 2 short sin/cos polynomials of a group overlapping. The velocity sub-step is
// written for every Vx != 0 as in MODE 3 (car_model.cuh: car_step_spec); its validity conditions are accumulated over
// the group BEFORE it is published, and a lane that leaves them (Vx changes sign while braking, den·Vx denormal,
// |δ| > 0.78, a standstill) redoes that group — and the rest of the control step — on the general path (all branches,
// libm where the un-wrapped angle matters). Nothing published is ever rolled back.
// Arithmetic: identical to MODE 3 on every sub-step both consider valid (same expressions, same fma placement);
// control steps MODE 3 would repair as a whole are here repaired from the offending group on — both are the
// reference's mathematics to ~1e-13, tests/test_gpu_parity.py pins each against the oracle.
#include <cuda/ptx>
#include <math_constants.h>

#include "car_model.cuh"
#include "engine.cuh"

namespace mpopis {

namespace {

namespace ptx = cuda::ptx;

constexpr int GROUP = 5, NGROUP = 3, RING = GROUP * NGROUP;  // sub-step slots per velocity warp, handed over in groups
constexpr int GPS = 2;                                        // groups per control step: the kernel requires nsub == 10
constexpr int VW = 2;                                         // velocity warps per CTA
constexpr int NSTAGE = 4;                                     // control steps of noise in flight per velocity warp

template <int NCARS>
struct SplitSmem {
  double ring[VW][RING][NCARS][3][32];  // (Vx, Vy, Ψ̇δt) after each sub-step
  double ext[VW][2][NCARS][3][32];      // per control step (parity-buffered): Ψ̇, δ, pedal at its end (trajectory log)
  double fin[VW][32];                   // control cost (POL:272) of the finished rollout
  double noise[VW][NSTAGE][2 * NCARS][32];  // E[:, k] of the next control steps, brought in by cp.async (no register staging)
  uint64_t full[VW][NGROUP], empty[VW][NGROUP];
  unsigned nfull[VW][NGROUP], nempty[VW][NGROUP];  // SPIN hand-over: completed fills / drains of each group
};

__device__ __forceinline__ void bar_wait(uint64_t *bar, unsigned parity, long long *waited = nullptr) {
  if (ptx::mbarrier_try_wait_parity(bar, parity)) return;
  const long long t0 = waited ? clock64() : 0;
  while (!ptx::mbarrier_try_wait_parity(bar, parity)) {
  }
  if (waited) *waited += clock64() - t0;
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
  (void)ptx::mbarrier_arrive(ptx::sem_release, ptx::scope_cta, ptx::space_shared, bar);
}
// SPIN hand-over (latency regime: at most one CTA per SM, so a polling warp takes issue slots from nobody): a counter
// per group in shared memory, bumped by one lane after the warp's stores are fenced, polled by every lane of the peer.
// mbarrier.try_wait parks the warp and costs ≈ 1000 cycles per control step at K = 150 (profiles/r2_split_ablation.txt).
__device__ __forceinline__ void spin_wait(const unsigned *cnt, unsigned want) {
  while (*reinterpret_cast<const volatile unsigned *>(cnt) < want) {
  }
  __threadfence_block();  // acquire: the ring slots are read after the counter
}
__device__ __forceinline__ void spin_post(unsigned *cnt, unsigned value, int lane) {
  __threadfence_block();  // release: this lane's ring stores before the counter
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile unsigned *>(cnt) = value;
}

// The noise of a rollout — AS doubles per control step, coalesced 256-byte rows across the warp — is copied global ->
// shared memory asynchronously (cp.async, 8 bytes per lane and row, three control steps ahead). A register prefetch
// does not survive here: under register pressure ptxas spills the loaded value right behind the load, and on an in-order
// warp that store waits for the whole global-memory latency (ncu, K = 4096: 10 % of all stall samples sat on that one
// STL pair). cp.async involves no register, so nothing waits until the values are read two steps later.
__device__ __forceinline__ void noise_issue(double *dst_smem, const double *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void noise_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void noise_wait() {  // all but the N most recent groups of this thread have landed
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// GROUP Euler sub-steps of the velocity recurrence on the general path (car_step_fast's loop body, CAR:301-328, without
// the pose): every branch the reference's arithmetic can take. Rare (a lane that left the straight-line conditions), so
// it is kept out of line, recomputes the tyre constants instead of carrying them and writes its ring slots itself.
struct VelState {
  double Vx, Vy, psid, sg;
};
__device__ __noinline__ VelState vel_group_general(const CarParams &P, double ddt, double accel, double bk, double split,
                                                   double delta0, double dlt, int i0, VelState st, double *slot0,
                                                   int slot_stride) {
  double Vx = st.Vx, Vy = st.Vy, psid = st.psid, sg = st.sg;
  const double inv_Izz = 1 / P.Izz, inv_m = 1 / P.m;
  for (int j = 0; j < GROUP; ++j) {
    const double delta = fma((double)(i0 + j + 1), dlt, delta0);  // CAR:301
    double sd, cd;
    sincos(delta, &sd, &cd);
    if (!(Vx > 0.0 && sg > 0.0)) {
      const double sg_now = jl_sign(Vx);
      if (sg_now != sg) sg = sg_now;  // sign(Vx) flipped: the brake force changes direction (CAR:311)
    }
    const TireConsts tc = tire_consts_fast(P, accel, bk, split, sg);
    const double yf = Vy + P.l_f * psid, yr = Vy - P.l_r * psid;
    double fyf, fyr;
    if (Vx > 0.0) {
      fyf = tire_fy_ratio<0>(yf * cd - Vx * sd, Vx * cd + yf * sd, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f, tc.fymax_f);
      fyr = tire_fy_ratio<0>(yr, Vx, P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
    } else {  // reversing / standstill: the un-wrapped slip angle matters, keep the libm sequence (CAR:304-305)
      fyf = tire_fy_literal(atan2(yf, Vx) - delta, P.C_af, tc.c2_f, tc.c3_f, tc.thr_f, tc.fymax_f);
      fyr = tire_fy_literal(atan2(yr, Vx), P.C_ar, tc.c2_r, tc.c3_r, tc.thr_r, tc.fymax_r);
    }
    const double fx_aero = (P.C_D0 + P.C_D1 * fabs(Vx)) * sg;                              // CAR:308
    const double psidd = inv_Izz * (P.l_f * (tc.fxf * sd + fyf * cd) - P.l_r * fyr);        // CAR:322
    const double Vy_dot = inv_m * (fyf * cd + tc.fxf * sd + fyr) - psid * Vx;               // CAR:323
    const double Vx_dot = inv_m * (tc.fxf * cd - fyf * sd + tc.fxr - fx_aero) + psid * Vy;  // CAR:324
    psid += psidd * ddt;  // CAR:326
    Vx += Vx_dot * ddt;   // CAR:327
    Vy += Vy_dot * ddt;   // CAR:328
    double *sl = slot0 + (size_t)j * slot_stride;  // ring[vw][slot][car][{Vx, Vy, Ψ̇δt}][lane]
    sl[0] = Vx, sl[32] = Vy, sl[64] = psid * ddt;
  }
  return VelState{Vx, Vy, psid, sg};
}

// sin/cos of a heading increment that left the short polynomial's range (a spinning car): rare, out of line
struct SinCos {
  double s, c;
};
__device__ __noinline__ SinCos sincos_increment_general(double dpsi) {
  SinCos o;
  if (fabs(dpsi) <= 0.8) sincos_kernel(dpsi, &o.s, &o.c);
  else sincos(dpsi, &o.s, &o.c);
  return o;
}

struct VelCar {  // velocity-side state of one car
  double Vx, Vy, psid, delta, sd, cd;
  bool trig_valid;
};
struct PoseCar {  // pose-side state of one car
  double x, y, psi, sp, cp;
};

// Everything a control step needs that does NOT depend on the velocities: action (clamped U + noise), steering rate,
// tyre constants for an assumed sign(Vx), sin/cos of the steering increment and of δ at the step's start. δ advances
// by nsub·rate·δt, so δ(t+1) is known when step t begins: the constants of step t+1 are computed DURING step t (inside
// the straight-line block of its first sub-step group, where ptxas interleaves them with the recurrence) instead of
// in front of it — ≈ 850 of ≈ 3100 cycles per control step of a lone velocity warp were this prologue (ablations:
// tools/ablate.py, profiles/r2_split_ablation.txt).
struct StepK {  // kept small: two of these (this step's and the next one's) are live in registers at once
  TireConsts tc;
  double dlt, sdl, cdl, pedal;
  int sgp;  // sign(Vx) the tyre constants were built for: -1, 0, 1
  bool pre_ok;
};

__device__ __forceinline__ StepK make_step(const CarParams &P, const CarDerived &D, double dt, double ddt, double v0,
                                           double v1, double delta0, int sg_guess) {
  StepK k;
  const double a0 = clamp1(v0), a1 = clamp1(v1);                                  // UTL:55-67
  const double tgt = a0 * P.d_max - delta0;
  const double rate = fmin(fast_div(fabs(tgt), dt), P.dd_max) * jl_sign(tgt);     // CAR:295-296
  const double accel = P.Fx_max * fmax(a1, 0.0);                                  // CAR:310
  const double bk = P.Fx_min * fmin(a1, 0.0);                                     // CAR:311 without sign(Vx)
  const double split = a1 <= 0.0 ? P.l_brake : P.l_drive;
  k.sgp = sg_guess;
  k.tc = tire_consts_der(P, D, accel, bk, split, (double)sg_guess);
  k.dlt = rate * ddt;
  k.pre_ok = (fmax(fabs(delta0), fabs(a0 * P.d_max)) <= 0.78) & (fabs(k.dlt) <= 0.03);
  sincos_tiny(k.dlt, &k.sdl, &k.cdl);       // |rate·δt| <= δ̇_max·δt = 0.0157 for the default car
  k.pedal = a1;
  return k;
}
__device__ __forceinline__ int sign_int(double x) { return x > 0.0 ? 1 : (x < 0.0 ? -1 : 0); }

template <int NCARS, bool SPIN>
__device__ __forceinline__ void velocity_warp(const CarEnvArgs &env, const RolloutArgs &a, SplitSmem<NCARS> &sm,
                                              const double *Us, int vw, int k, int lane, long long *waited) {
  constexpr int AS = 2 * NCARS;
  const int T = a.T, cs = AS * T;
  const double ddt = env.ddt;
  const double *Uc = Us, *Uo = Us + cs, *Bv = Us + 2 * cs;  // pol.U, U_orig, (γ U_origᵀ Σ⁻¹) staged in shared memory
  VelCar car[NCARS];
#pragma unroll
  for (int c = 0; c < NCARS; ++c) {
    car[c].Vx = __ldg(a.state0 + 8 * c + 3), car[c].Vy = __ldg(a.state0 + 8 * c + 4);
    car[c].psid = __ldg(a.state0 + 8 * c + 5), car[c].delta = __ldg(a.state0 + 8 * c + 6);
    car[c].sd = 0.0, car[c].cd = 1.0, car[c].trig_valid = false;
  }
  const double *Ek = a.E + k;
  double cc = 0.0;
  auto fetch = [&](int t) {  // start the copy of control step t's noise (an empty group past the horizon keeps the count)
    if (t < T) {
#pragma unroll
      for (int r = 0; r < AS; ++r) noise_issue(&sm.noise[vw][t % NSTAGE][r][lane], Ek + (size_t)(t * AS + r) * a.ldk);
    }
    noise_commit();
  };
  fetch(0), fetch(1), fetch(2);
  noise_wait<2>();  // step 0 has landed
  StepK cur[NCARS], nxt[NCARS];
#pragma unroll
  for (int c = 0; c < NCARS; ++c) {
    const double v0 = Uc[2 * c] + sm.noise[vw][0][2 * c][lane], v1 = Uc[2 * c + 1] + sm.noise[vw][0][2 * c + 1][lane];  // Vₖ = pol.U + E[:,k], POL:271
    if (a.bvec) cc += Bv[2 * c] * (v0 - Uo[2 * c]) + Bv[2 * c + 1] * (v1 - Uo[2 * c + 1]);  // POL:272
    cur[c] = make_step(env.car[c], env.der[c], env.dt, ddt, v0, v1, car[c].delta, sign_int(car[c].Vx));
    nxt[c] = cur[c];
  }
  int gi = 0;  // global group counter of this rollout
  for (int t = 0; t < T; ++t) {
    fetch(t + 3);  // its slot held step t − 1, consumed during step t − 2
    // ---- what DOES depend on the velocities at the step's start ----
    double dpsi[NCARS], sg[NCARS];
    int hvx0[NCARS], brake_mask[NCARS];
    bool gen[NCARS];  // this lane integrates the rest of the control step on the general path
#pragma unroll
    for (int c = 0; c < NCARS; ++c) {
      const CarParams &P = env.car[c];
      const double a1 = cur[c].pedal, bk = P.Fx_min * fmin(a1, 0.0);
      sg[c] = jl_sign(car[c].Vx);
      if (bk != 0.0 && sign_int(car[c].Vx) != cur[c].sgp)  // rare: the car changed direction while the constants were in flight
        cur[c].tc = tire_consts_der(P, env.der[c], P.Fx_max * fmax(a1, 0.0), bk, a1 <= 0.0 ? P.l_brake : P.l_drive, sg[c]);
      hvx0[c] = hi32(car[c].Vx), brake_mask[c] = bk != 0.0 ? (int)0x80000000 : 0;
      gen[c] = !(cur[c].pre_ok & (car[c].Vx != 0.0));
      if (!car[c].trig_valid || (t % 5) == 0) sincos_kernel(car[c].delta, &car[c].sd, &car[c].cd);  // re-synchronise the δ recurrence
      dpsi[c] = car[c].psid * ddt;
    }
#pragma unroll
    for (int q = 0; q < GPS; ++q, ++gi) {
      const int grp = gi % NGROUP;
      if (gi >= NGROUP) {
        if (SPIN) spin_wait(&sm.nempty[vw][grp], (unsigned)(gi / NGROUP));
        else bar_wait(&sm.empty[vw][grp], (unsigned)((gi / NGROUP - 1) & 1), waited);
      }
      if (q == 0 && t + 1 < T) {  // constants of step t + 1, overlapped with this group's recurrence
        noise_wait<2>();            // steps t + 2 and t + 3 may still be in flight, step t + 1 has landed
#pragma unroll
        for (int c = 0; c < NCARS; ++c) {
          const int row = (t + 1) * AS + 2 * c;
          const double *en = &sm.noise[vw][(t + 1) % NSTAGE][2 * c][lane];
          const double v0 = Uc[row] + en[0], v1 = Uc[row + 1] + en[32];
          if (a.bvec) cc += Bv[row] * (v0 - Uo[row]) + Bv[row + 1] * (v1 - Uo[row + 1]);
          nxt[c] = make_step(env.car[c], env.der[c], env.dt, ddt, v0, v1, fma((double)env.nsub, cur[c].dlt, car[c].delta),
                             sign_int(car[c].Vx));
        }
      }
#pragma unroll
      for (int c = 0; c < NCARS; ++c) {
        const CarParams &P = env.car[c];
        const CarDerived &D = env.der[c];
        const TireConsts &tc = cur[c].tc;
        const double sdl = cur[c].sdl, cdl = cur[c].cdl;
        double Vx = car[c].Vx, Vy = car[c].Vy, psid = car[c].psid, sd = car[c].sd, cd = car[c].cd, dp = dpsi[c];
        int bad = 0;
        // --- GROUP straight-line sub-steps, valid for every Vx != 0 and den != 0. The arithmetic is car_step_spec's
        // (car_model.cuh), re-associated so that the loop-carried chain (Vx, Vy, Ψ̇) -> ... -> (Vx, Vy, Ψ̇) is short:
        //   * tan α_f = num/den and tan α_r = y_r/Vx through two reciprocal seeds (MUFU) instead of one of den·Vx — the
        //     rear one starts from Vx, available when the sub-step begins; the Newton correction is applied to the
        //     quotient, t = t₀(1 + e + e²) with t₀ = num·r₀, not to the reciprocal first (one operation fewer in series);
        //   * the brush-tyre cubic (CAR:256) as fma(t|t|, c2 − c3|t|, −C t): depth 2 instead of 3;
        //   * the Euler updates (CAR:322-328) written as ONE fma on each tyre force, everything else pre-summed:
        //       Ψ̇' = (c₁cosδ)F_yf + [c₁F_xf sinδ + Ψ̇ − c₂F_yr],  Vy' = (c_m cosδ)F_yf + [c_m(F_xf sinδ + F_yr) + Vy − ΨδVx],
        //       Vx' = −(c_m sinδ)F_yf + [c_m F_xf cosδ + ΨδVy + k_x Vx + c_m F_xr ∓ c_m C_D0].
        // A lone warp runs one such sub-step in ≈ 150 cycles (tools/chain_bench.cu: 168 with, 137 without the validity
        // bookkeeping; DFMA -> DFMA issues after ≈ 9 cycles, DMUL -> DFMA ≈ 17, MUFU.RCP64H -> DFMA ≈ 27 on B200).
        // Differences to MODE 3 are re-association roundings (1e-16 relative per operation). The statements are
        // written level by level of the dependence graph, the order ptxas largely keeps.
        const double cI1fxf = D.cI1 * tc.fxf, cmfxf = D.cm * tc.fxf, cmfxr = D.cm * tc.fxr;
        const double nddt = -ddt;
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          // level 0: the δ recurrence (independent of the velocities) and everything that needs the state only
          const double ns = fma(sd, cdl, cd * sdl);  // sin/cos(δ + rate·δt), CAR:301
          const double nc = fma(cd, cdl, -(sd * sdl));
          sd = ns, cd = nc;
          const int hvx = hi32(Vx);
          const bool fwd = hvx >= 0;
          const double rx = rcp_seed(Vx);  // 2^-23 seed of 1/Vx
          const double yf = fma(P.l_f, psid, Vy), yr = fma(-P.l_r, psid, Vy);
          const double wx = Vx * nddt, wy = Vy * ddt;                      // −δt·Vx, δt·Vy
          const double k3 = fma(Vx, D.kx, cmfxr + with_opposite_sign(D.cmCD0, hvx));
          // level 1
          const double vxsd = Vx * sd, vxcd = Vx * cd;
          const double ex = fma(-Vx, rx, 1.0);
          const double t0r = yr * rx;
          const double k2 = fma(psid, wx, Vy);                             // Vy − Ψ̇δt·Vx          (CAR:323)
          const double k3b = fma(psid, wy, k3);                            // … + Ψ̇δt·Vy          (CAR:324)
          const double k1 = fma(cI1fxf, sd, psid);                         // Ψ̇ + c₁F_xf sinδ      (CAR:322)
          const double qI = D.cI1 * cd, qy = D.cm * cd, qx = D.cm * sd;
          const double fxfsd = tc.fxf * sd;
          // level 2
          const double num = fma(yf, cd, -vxsd), den = fma(yf, sd, vxcd);  // tan α_f = num/den, tan α_r = yr/Vx
          const double px = fma(ex, ex, ex);
          const double Kx = fma(cmfxf, cd, k3b);
          // level 3
          const double r0 = rcp_seed(den);
          const double ta_r = fma(t0r, px, t0r);  // t₀(1 + e + e²): relative error e³ = 2^-69
          bad |= ((hvx ^ hvx0[c]) & brake_mask[c]) | ((hi32(den) & 0x7ff00000) - 0x00100000) |
                 ((hvx & 0x7ff00000) - 0x00100000);
          // level 4
          const double e = fma(-den, r0, 1.0), t0 = num * r0;
          const double atr = fabs(ta_r);
          const double ur = ta_r * atr, vr = fma(-tc.c3_r, atr, tc.c2_r), x1r = -P.C_ar * ta_r;
          const bool lin_r = fwd & (atr < tc.thr_r);
          // level 5
          const double pe = fma(e, e, e);
          const double cubic_r = fma(ur, vr, x1r);  // CAR:256 as fma(t|t|, c2 − c3|t|, −C t)
          // level 6
          const double ta = fma(t0, pe, t0);
          const double fyr = lin_r ? cubic_r : with_opposite_sign(tc.fymax_r, hi32(yr));  // CAR:255-259
          // level 7
          const double at = fabs(ta);
          const double u = ta * at, v = fma(-tc.c3_f, at, tc.c2_f), x1 = -P.C_af * ta;
          const bool lin_f = (hi32(den) >= 0) & (at < tc.thr_f);
          const double Kp = fma(-D.cI2, fyr, k1);
          const double Ky = fma(D.cm, fxfsd + fyr, k2);
          // level 8
          const double cubic = fma(u, v, x1);
          const double fyf = lin_f ? cubic : with_opposite_sign(tc.fymax_f, fwd ? hi32(num) : hi32(yf));
          // level 9: the new state — one fma on F_yf each (CAR:322-328)
          psid = fma(qI, fyf, Kp), Vy = fma(qy, fyf, Ky), Vx = fma(-qx, fyf, Kx);
          dp = psid * ddt;
          sm.ring[vw][grp * GROUP + j][c][0][lane] = Vx;
          sm.ring[vw][grp * GROUP + j][c][1][lane] = Vy;
          sm.ring[vw][grp * GROUP + j][c][2][lane] = dp;
        }
        gen[c] = gen[c] | (bad < 0);
        if (gen[c]) {  // rare: redo THIS group on the general path from its un-advanced state, stay there for the step
          const double a1 = cur[c].pedal;
          const VelState o = vel_group_general(P, ddt, P.Fx_max * fmax(a1, 0.0), P.Fx_min * fmin(a1, 0.0),
                                               a1 <= 0.0 ? P.l_brake : P.l_drive, car[c].delta, cur[c].dlt,
                                               q * GROUP, VelState{car[c].Vx, car[c].Vy, car[c].psid, sg[c]},
                                               &sm.ring[vw][grp * GROUP][c][0][lane], NCARS * 3 * 32);
          Vx = o.Vx, Vy = o.Vy, psid = o.psid, sg[c] = o.sg;
          dp = psid * ddt;
        }
        car[c].Vx = Vx, car[c].Vy = Vy, car[c].psid = psid, car[c].sd = sd, car[c].cd = cd, dpsi[c] = dp;
      }
      if (q + 1 == GPS) {
#pragma unroll
        for (int c = 0; c < NCARS; ++c) {
          car[c].delta = fma((double)env.nsub, cur[c].dlt, car[c].delta);  // CAR:301 summed (= nxt[c].delta0)
          car[c].trig_valid = !gen[c];  // after a general step the δ recurrence restarts from δ itself
          // end-of-step extras for the trajectory log. ext[t & 1] was last read at the end of step t − 2, which the
          // pose warp has passed: this warp is at most NGROUP = 3 groups ahead of it
          sm.ext[vw][t & 1][c][0][lane] = car[c].psid;
          sm.ext[vw][t & 1][c][1][lane] = car[c].delta;
          sm.ext[vw][t & 1][c][2][lane] = cur[c].pedal;  // CAR:297, state[8]
        }
        if (t + 1 == T) sm.fin[vw][lane] = cc;
      }
      if (SPIN) spin_post(&sm.nfull[vw][grp], (unsigned)(gi / NGROUP + 1), lane);
      else {
        __syncwarp();  // every lane's slots of the group are written: hand it over
        if (lane == 0) bar_arrive(&sm.full[vw][grp]);
      }
    }
#pragma unroll
    for (int c = 0; c < NCARS; ++c) cur[c] = nxt[c];
  }
}

// NV rollouts per lane: the pose warp serves the velocity warps v0 .. v0 + NV − 1
template <int NCARS, int NV, bool SPIN>
__device__ __forceinline__ void pose_warp(const CarEnvArgs &env, const RolloutArgs &a, SplitSmem<NCARS> &sm,
                                          const TrackView &tr, int kbase, int v0, int nvw, int lane, long long *waited) {
  constexpr int SS = 8 * NCARS;
  const int T = a.T;
  const double ddt = env.ddt;
  PoseCar pc[NV][NCARS];
  double cost[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    cost[v] = 0.0;
#pragma unroll
    for (int c = 0; c < NCARS; ++c) {
      pc[v][c].x = __ldg(a.state0 + 8 * c + 0), pc[v][c].y = __ldg(a.state0 + 8 * c + 1);
      pc[v][c].psi = __ldg(a.state0 + 8 * c + 2), pc[v][c].sp = 0.0, pc[v][c].cp = 1.0;
    }
  }
  int gi = 0;
  for (int t = 0; t < T; ++t) {
    if ((t % 5) == 0) {  // re-synchronise the heading recurrence (as MODE 3 does every 5th control step)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < NCARS; ++c) {
          double psi = pc[v][c].psi;
          if (fabs(psi) > CUDART_PI) {  // callers may hand in any heading; CAR:330 keeps it in (−π, π] afterwards
            const double kk = rint(psi * 0.15915494309189535);
            psi = fma(-kk, 6.283185307179586, psi);
            psi = fma(-kk, 2.4492935982947064e-16, psi);
            pc[v][c].psi = psi;
          }
          sincos_pi(psi, &pc[v][c].sp, &pc[v][c].cp);
        }
    }
    double vx_end[NV][NCARS], vy_end[NV][NCARS];
#pragma unroll
    for (int q = 0; q < GPS; ++q, ++gi) {
      const int grp = gi % NGROUP;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (v0 + v < nvw) {
          if (SPIN) spin_wait(&sm.nfull[v0 + v][grp], (unsigned)(gi / NGROUP + 1));
          else bar_wait(&sm.full[v0 + v][grp], (unsigned)((gi / NGROUP) & 1), waited);
        }
      // GROUP sub-steps of NV x NCARS independent poses in one straight-line block: the short sin/cos polynomials of
      // all increments overlap, only the heading rotation (two operations per sub-step) is sequential
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < NCARS; ++c) {
          PoseCar &p = pc[v][c];
          double Vx[GROUP], Vy[GROUP], dpsi[GROUP], sdp[GROUP], cdp[GROUP];
          int wide = 0;
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            Vx[j] = sm.ring[v0 + v][grp * GROUP + j][c][0][lane], Vy[j] = sm.ring[v0 + v][grp * GROUP + j][c][1][lane];
            dpsi[j] = sm.ring[v0 + v][grp * GROUP + j][c][2][lane];
            sincos_tiny(dpsi[j], &sdp[j], &cdp[j]);
            wide |= 0x3F9EB851 - (hi32(dpsi[j]) & 0x7fffffff);  // |Ψ̇δt| > 0.03 or NaN
          }
          if (wide < 0) {  // a spinning car: the general sin/cos for the increments that need it (rare)
#pragma unroll
            for (int j = 0; j < GROUP; ++j)
              if ((0x3F9EB851 - (hi32(dpsi[j]) & 0x7fffffff)) < 0) {
                const SinCos o = sincos_increment_general(dpsi[j]);
                sdp[j] = o.s, cdp[j] = o.c;
              }
          }
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            p.psi += dpsi[j];  // CAR:329 (wrapped once per step below)
            const double nsp = fma(p.sp, cdp[j], p.cp * sdp[j]);
            p.cp = fma(p.cp, cdp[j], -(p.sp * sdp[j]));
            p.sp = nsp;
            p.x = fma(fma(Vx[j], p.cp, -(Vy[j] * p.sp)), ddt, p.x);  // CAR:331
            p.y = fma(fma(Vx[j], p.sp, Vy[j] * p.cp), ddt, p.y);     // CAR:332
          }
          vx_end[v][c] = Vx[GROUP - 1], vy_end[v][c] = Vy[GROUP - 1];
        }
      if (SPIN) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (v0 + v < nvw) spin_post(&sm.nempty[v0 + v][grp], (unsigned)(gi / NGROUP + 1), lane);
      } else {
        __syncwarp();  // every lane has read the group: give it back
        if (lane == 0 && !(q + 1 == GPS && t + 1 == T)) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (v0 + v < nvw) bar_arrive(&sm.empty[v0 + v][grp]);
        }
      }
    }
    // ---- end of the control step: heading wrap (CAR:330), reward (CAR:201-213 / MCR:145-158), log ----
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s[SS];
#pragma unroll
      for (int c = 0; c < NCARS; ++c) {
        PoseCar &p = pc[v][c];
        if (fabs(p.psi) > CUDART_PI) {
          const double kk = rint(p.psi * 0.15915494309189535);
          p.psi = fma(-kk, 6.283185307179586, p.psi);
          p.psi = fma(-kk, 2.4492935982947064e-16, p.psi);
        }
        s[8 * c + 0] = p.x, s[8 * c + 1] = p.y, s[8 * c + 2] = p.psi;
        s[8 * c + 3] = vx_end[v][c], s[8 * c + 4] = vy_end[v][c];
      }
      double rew = 0.0;
#pragma unroll
      for (int c = 0; c < NCARS; ++c) {
        rew += car_reward<3>(env.car[c], env.cos_blimit[c], tr, s + 8 * c);
#pragma unroll
        for (int j = c + 1; j < NCARS; ++j) {
          const double dx = s[8 * j] - s[8 * c], dy = s[8 * j + 1] - s[8 * c + 1];
          const double dd = sqrt_fast(dx * dx + dy * dy);
          rew += -dd;
          if (dd <= 4.0) rew += -11000.0;  // MCR:153-155 (docstring says −7000; code is −11000)
        }
      }
      cost[v] -= rew;  // UTL:137-138
      const int k = kbase + (v0 + v) * 32 + lane;
      if (a.traj && v0 + v < nvw && k < a.K) {
#pragma unroll
        for (int c = 0; c < NCARS; ++c) {
          s[8 * c + 5] = sm.ext[v0 + v][t & 1][c][0][lane], s[8 * c + 6] = sm.ext[v0 + v][t & 1][c][1][lane];
          s[8 * c + 7] = sm.ext[v0 + v][t & 1][c][2][lane];
        }
#pragma unroll
        for (int q = 0; q < SS; ++q) a.traj[((size_t)k * SS + q) * T + t] = s[q];  // UTL:139-141
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int k = kbase + (v0 + v) * 32 + lane;
    if (v0 + v < nvw && k < a.K) a.costs[k] = cost[v] + sm.fin[v0 + v][lane];  // POL:274-275
  }
}

// NV: rollouts per pose lane (pose warps per CTA = VW / NV). REGS: register budget per thread — the velocity chain's
// schedule and the pose warp's unrolled groups want ≈ 150; 96 keeps every rollout of K = 65 536 resident at once.
template <int NCARS, int NV, int REGS, bool SPIN>
__global__ void __launch_bounds__(32 * (VW + VW / NV)) __maxnreg__(REGS)
    rollout_car_split_kernel(const __grid_constant__ CarEnvArgs env, const __grid_constant__ RolloutArgs a, const int *stop) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (stop && *stop) return;
  SplitSmem<NCARS> &sm = *reinterpret_cast<SplitSmem<NCARS> *>(smem_raw);
  double *trk_s = reinterpret_cast<double *>(smem_raw + sizeof(SplitSmem<NCARS>));
  double *Us = trk_s + 3 * env.n_trk;  // pol.U | U_orig | bvec, cs doubles each
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, cs = 2 * NCARS * a.T;
  const int kbase = blockIdx.x * (32 * VW);
  const int nvw = min(VW, (a.K - kbase + 31) / 32);  // velocity warps of this CTA that own at least one rollout
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < VW; ++v)
#pragma unroll
      for (int q = 0; q < NGROUP; ++q) {
        ptx::mbarrier_init(&sm.full[v][q], 1), ptx::mbarrier_init(&sm.empty[v][q], 1);
        sm.nfull[v][q] = 0, sm.nempty[v][q] = 0;
      }
    ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
  }
  const long long t_begin = a.warp_cycles ? clock64() : 0;
  for (int i = threadIdx.x; i < 3 * env.n_trk; i += blockDim.x) trk_s[i] = env.trk[i];
  for (int i = threadIdx.x; i < cs; i += blockDim.x) {
    Us[i] = __ldg(a.U + i), Us[cs + i] = __ldg(a.U_orig + i);
    Us[2 * cs + i] = a.bvec ? __ldg(a.bvec + i) : 0.0;
  }
  if (nvw < VW)  // a pose warp integrates its rollouts unconditionally: give the absent ones benign inputs
    for (int i = threadIdx.x; i < (int)(sizeof(sm.ring) / sizeof(double)); i += blockDim.x) (&sm.ring[0][0][0][0][0])[i] = 0.0;
  __syncthreads();
  const TrackView tr{trk_s, trk_s + env.n_trk, trk_s + 2 * env.n_trk, env.n_trk,
                     env.lut, env.lut_x0, env.lut_y0, env.lut_inv_c, env.lut_nx, env.lut_ny};
  long long waited = 0, *wp = a.warp_cycles ? &waited : nullptr;
  if (w < VW) {
    if (w < nvw) velocity_warp<NCARS, SPIN>(env, a, sm, Us, w, min(kbase + w * 32 + lane, a.K - 1), lane, wp);  // padding lanes copy the last rollout
  } else {
    const int v0 = (w - VW) * NV;
    if (v0 < nvw) pose_warp<NCARS, NV, SPIN>(env, a, sm, tr, kbase, v0, nvw, lane, wp);
  }
  if (a.warp_cycles && lane == 0) {  // "rollout_profile": [total | waiting on the ring] cycles per warp
    constexpr int WPC = VW + VW / NV;
    a.warp_cycles[2 * (blockIdx.x * WPC + w)] = clock64() - t_begin;
    a.warp_cycles[2 * (blockIdx.x * WPC + w) + 1] = waited;
  }
}

}  // namespace

int rollout_split_max_cars() { return 3; }

// returns 0 when the configuration is not covered (the caller falls back to the thread-per-rollout kernel)
template <int N, int NV, int REGS, bool SPIN>
static int launch_split(const CarEnvArgs &env, const RolloutArgs &a, const int *stop, cudaStream_t st) {
  const int grid = (a.K + 32 * VW - 1) / (32 * VW);
  const size_t smem = sizeof(SplitSmem<N>) + sizeof(double) * (3 * (size_t)env.n_trk + 3 * (size_t)2 * N * a.T);
  if (smem > 200 * 1024) return 0;
  cudaFuncSetAttribute(rollout_car_split_kernel<N, NV, REGS, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // all of the SM's L1/shared array as shared memory: 7 CTAs x 28 KB must be resident together at K = 65 536
  cudaFuncSetAttribute(rollout_car_split_kernel<N, NV, REGS, SPIN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  rollout_car_split_kernel<N, NV, REGS, SPIN><<<grid, 32 * (VW + VW / NV), smem, st>>>(env, a, stop);
  return 1;
}

// Largest shard for which the split kernel is the default — the latency regime, where a launch lasts as long as one
// rollout's chain: up to two CTAs (64 rollouts, 4 warps) per SM the 255-register / spin flavour, up to ≈ 2.6 per SM the
// 160-register one. Measured against the thread-per-rollout kernel (MODE 3: every rollout of K = 65 536 resident at 128
// registers), µs per launch: K = 150: 87 vs 150, 8 192: 95 vs 146, 18 944: 111 vs 146, 24 576: 147 vs 170; at K = 65 536
// the split kernel needs several waves (372-495 vs 271) — profiles/r2_ab_variants.txt.
int rollout_split_capacity(int n_cars, int num_sms) { return n_cars == 1 ? num_sms * 168 : num_sms * 64; }

// wide: 0 = 2 velocity + 1 pose warp, 96 registers (7 CTAs per SM); 1 = 2 + 2 warps, 160 registers (3 CTAs per SM)
// spin: poll shared-memory counters instead of parking on mbarriers (only sensible while a polling warp shares its
// scheduler with nobody: at most one CTA per SM)
int launch_rollout_car_split(const CarEnvArgs &env, const RolloutArgs &a, int wide, int spin, const int *stop,
                             cudaStream_t st) {
  if (env.nsub != GROUP * GPS) return 0;  // the ring hands over two groups of five sub-steps per control step
  switch (env.n_cars) {
    case 1:
      if (!wide) return launch_split<1, 2, 96, false>(env, a, stop, st);
      // at most one CTA per SM (spin): the register file is not a constraint either — no spills at 255
      return spin ? launch_split<1, 1, 255, true>(env, a, stop, st) : launch_split<1, 1, 160, false>(env, a, stop, st);
    case 2: return spin ? launch_split<2, 1, 255, true>(env, a, stop, st) : launch_split<2, 1, 255, false>(env, a, stop, st);
    case 3: return spin ? launch_split<3, 1, 255, true>(env, a, stop, st) : launch_split<3, 1, 255, false>(env, a, stop, st);
  }
  return 0;
}

}  // namespace mpopis
