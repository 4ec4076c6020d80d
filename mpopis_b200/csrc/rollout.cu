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
#include <cuda/ptx>
#include <math_constants.h>

#include "car_model.cuh"
#include "engine.cuh"

namespace mpopis {

// (env)(a) + reward(env) for 1..N cars: CAR:238-241 / MCR:200-207, MCR:145-158
template <int NCARS, int MODE>
__device__ __forceinline__ double cars_step_reward(const CarEnvArgs &env, const TrackView &tr, double *s,
                                                   const double *a, double *trig, bool resync, bool *trig_valid) {
  if constexpr (MODE == 3) {  // every car's straight-line step first (independent chains), then the rare repairs
    double o[8 * NCARS];
    bool ok[NCARS];
#pragma unroll
    for (int c = 0; c < NCARS; ++c)
      ok[c] = car_step_spec(env.car[c], env.der[c], env.dt, env.ddt, env.nsub, s + 8 * c, o + 8 * c, a[2 * c],
                            a[2 * c + 1], trig + 4 * c, resync);
    bool all_ok = true;
#pragma unroll
    for (int c = 0; c < NCARS; ++c) {
      if (ok[c]) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s[8 * c + q] = o[8 * c + q];
      } else {
        car_step_fast(env.car[c], env.dt, env.ddt, env.nsub, s + 8 * c, a[2 * c], a[2 * c + 1]);
        all_ok = false;
      }
    }
    *trig_valid = all_ok;  // a repaired step leaves no carried sin/cos: re-evaluate at the next step
  } else {
#pragma unroll
    for (int c = 0; c < NCARS; ++c)
      car_step<MODE>(env.car[c], env.dt, env.ddt, env.nsub, s + 8 * c, a[2 * c], a[2 * c + 1]);
  }
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

// STAGE 1: the noise tile of a warp — AS rows x 32 samples, 256 contiguous bytes per row — is brought into shared
// memory by the TMA bulk-copy engine (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), a few control steps
// ahead, in a per-warp ring: no CTA-wide barrier, no registers held across the step. ncu on the register prefetch of
// STAGE 0 (load E of step t+1, integrate step t): under the 128-register cap the prefetched values are spilled right
// after the load, so the warp waits for the load after all — 24 % of the per-step stall samples sit on those two
// STL instructions (≈11 % of the kernel; profiles/README.md). The kernel stays FP64-bound; this removes a stall, it
// does not turn it into a bandwidth kernel.
template <int NCARS, int MODE, int STAGE>
__global__ void __launch_bounds__(128, NCARS == 1 ? 4 : 1) rollout_car_kernel(const __grid_constant__ CarEnvArgs env,
                                                          const __grid_constant__ RolloutArgs a,
                                                          const int *stop) {
  extern __shared__ double smem[];
  constexpr int AS = 2 * NCARS, SS = 8 * NCARS;
  constexpr int D = NCARS <= 2 ? 4 : 2;  // ring depth (control steps in flight)
  __shared__ __align__(128) double Es[STAGE ? 4 : 1][STAGE ? D : 1][STAGE ? AS : 1][32];
  __shared__ __align__(8) uint64_t bars[STAGE ? 4 : 1][D];
  if (stop && *stop) return;
  const long long t_begin = a.warp_cycles ? clock64() : 0;
  const TrackView tr = stage_track(env, smem);
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long kw = (long long)blockIdx.x * blockDim.x + w * 32;  // first sample of this warp
  if constexpr (STAGE) {
    if (kw >= a.K) return;  // whole warp out of range; a partly filled warp keeps all lanes (they integrate padding)
  } else {
    if (k >= a.K) return;
  }
  auto issue = [&](int t) {  // lane 0: arm the stage's mbarrier and start the AS bulk copies of control step t
    namespace ptx = cuda::ptx;
    const int st = t % D;
    ptx::fence_proxy_async(ptx::space_shared);  // the stage was read through the generic proxy
    ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &bars[w][st], AS * 256);
#pragma unroll
    for (int r = 0; r < AS; ++r)
      ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, &Es[w][st][r][0],
                         a.E + (size_t)(t * AS + r) * a.ldk + kw, 256, &bars[w][st]);
  };
  if constexpr (STAGE) {
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < D; ++q) cuda::ptx::mbarrier_init(&bars[w][q], 1);
      cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
      for (int t = 0; t < D && t < a.T; ++t) issue(t);
    }
    __syncwarp();
  }
  double s[SS];
#pragma unroll
  for (int q = 0; q < SS; ++q) s[q] = __ldg(a.state0 + q);
  const double *Ek = a.E + k;
  double cost = 0.0, cc = 0.0;
  double trig[4 * NCARS];
  bool trig_valid = false;
  // STAGE 0: the noise of step t+1 is fetched while step t integrates
  double e_next[AS];
  if constexpr (!STAGE) {
#pragma unroll
    for (int r = 0; r < AS; ++r) e_next[r] = Ek[(size_t)r * a.ldk];
  }
  for (int t = 0; t < a.T; ++t) {
    double act[AS], e_cur[AS];
    if constexpr (STAGE) {
      const int st = t % D;
      while (!cuda::ptx::mbarrier_try_wait_parity(&bars[w][st], (unsigned)((t / D) & 1))) {
      }
#pragma unroll
      for (int r = 0; r < AS; ++r) e_cur[r] = Es[w][st][r][lane];
      __syncwarp();  // every lane has read the stage before it is refilled
      if (lane == 0 && t + D < a.T) issue(t + D);
    } else {
#pragma unroll
      for (int r = 0; r < AS; ++r) e_cur[r] = e_next[r];
      if (t + 1 < a.T) {
#pragma unroll
        for (int r = 0; r < AS; ++r) e_next[r] = Ek[(size_t)((t + 1) * AS + r) * a.ldk];
      }
    }
#pragma unroll
    for (int r = 0; r < AS; ++r) {
      const int row = t * AS + r;
      const double v = __ldg(a.U + row) + e_cur[r];  // Vₖ = pol.U + E[:,k], POL:271
      if (a.bvec) cc += __ldg(a.bvec + row) * (v - __ldg(a.U_orig + row));  // POL:272
      act[r] = clamp1(v);                                                   // UTL:55-67
    }
    // MODE 3 carries sin/cos of δ and Ψ across control steps; re-evaluated every 5th step and after a repair
    const bool resync = !trig_valid || (t % 5) == 0;
    cost -= cars_step_reward<NCARS, MODE>(env, tr, s, act, trig, resync, &trig_valid);  // UTL:137-138
    if (a.traj && k < a.K) {
#pragma unroll
      for (int q = 0; q < SS; ++q) a.traj[((size_t)k * SS + q) * a.T + t] = s[q];  // UTL:139-141
    }
  }
  if (k < a.K) a.costs[k] = cost + cc;  // POL:274-275
  if (a.warp_cycles && lane == 0) a.warp_cycles[k >> 5] = clock64() - t_begin;
}

// ---- work-queue variant ------------------------------------------------------------------------------------------
// tools/warp_cycles.py (clock64 per warp) shows what bounds a launch at K = 65 536: 2048 warps over 592 sub-partitions
// is 3.46 per scheduler, i.e. {4,4,3,3} per SM; the warp scheduler favours the OLDER warps, so on every SM the two
// warps of the last-placed CTA run at what is left and then alone — per-warp cycles: median 377 k, but 13.3 % of the
// warps (exactly 2 x 136 SMs) need 505 k, and they are the kernel's duration (540 k). The work does not come in units
// that divide evenly, so this kernel makes the units smaller and hands them out dynamically: a persistent grid of
// 3 warps per scheduler pulls (batch of 32 rollouts, unit of `unit_len` control steps) pairs from an atomic counter in
// unit-major order; the rollout state (s, cost, control cost, carried sin/cos) travels through a scratch buffer between
// units (L2-resident, 15 doubles per rollout per unit), the successor unit spins on a per-batch counter published with
// release/acquire. Whoever finishes early takes the next unit of ANY batch, so all batches advance at the same rate.
// A unit's predecessor was handed out earlier to a warp that is running, so the spin cannot deadlock whatever the
// residency.
__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NCARS, int MODE>
__global__ void __launch_bounds__(128, NCARS == 1 ? 4 : 1) rollout_car_queue_kernel(const __grid_constant__ CarEnvArgs env,
                                                                const __grid_constant__ RolloutArgs a,
                                                                const int *stop) {
  extern __shared__ double smem[];
  constexpr int AS = 2 * NCARS, SS = 8 * NCARS, NF = SS + 3 + 4 * NCARS;  // s | cost | cc | trig_valid | trig
  if (stop && *stop) return;
  const long long t_begin = a.warp_cycles ? clock64() : 0;
  const TrackView tr = stage_track(env, smem);
  const int lane = threadIdx.x & 31;
  const int nb = (a.K + 31) >> 5, nu = (a.T + a.unit_len - 1) / a.unit_len;
  const int total = nb * nu;
  int *head = a.ws_sync, *done = a.ws_sync + 1;
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(head, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= total) break;
    const int u = i / nb, b = i - u * nb;
    const int k = b * 32 + lane;
    double *wsb = a.ws + (size_t)b * NF * 32 + lane;  // field f of this lane at wsb[f * 32]
    double s[SS], cost = 0.0, cc = 0.0, trig[4 * NCARS];
    bool trig_valid = false;
    if (u == 0) {
#pragma unroll
      for (int q = 0; q < SS; ++q) s[q] = __ldg(a.state0 + q);
    } else {
      while (ld_acquire(done + b) < u) {
      }
#pragma unroll
      for (int q = 0; q < SS; ++q) s[q] = wsb[q * 32];
      cost = wsb[SS * 32], cc = wsb[(SS + 1) * 32], trig_valid = wsb[(SS + 2) * 32] != 0.0;
#pragma unroll
      for (int q = 0; q < 4 * NCARS; ++q) trig[q] = wsb[(SS + 3 + q) * 32];
    }
    const int t0 = u * a.unit_len, t1 = min(a.T, t0 + a.unit_len);
    const double *Ek = a.E + min(k, a.K - 1);  // lanes past K (last batch) integrate a copy of the last rollout
    double e_next[AS];
#pragma unroll
    for (int r = 0; r < AS; ++r) e_next[r] = Ek[(size_t)(t0 * AS + r) * a.ldk];
    for (int t = t0; t < t1; ++t) {
      double act[AS], e_cur[AS];
#pragma unroll
      for (int r = 0; r < AS; ++r) e_cur[r] = e_next[r];
      if (t + 1 < t1) {
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
      const bool resync = !trig_valid || (t % 5) == 0;
      cost -= cars_step_reward<NCARS, MODE>(env, tr, s, act, trig, resync, &trig_valid);  // UTL:137-138
      if (a.traj && k < a.K) {
#pragma unroll
        for (int q = 0; q < SS; ++q) a.traj[((size_t)k * SS + q) * a.T + t] = s[q];  // UTL:139-141
      }
    }
    if (u == nu - 1) {
      if (k < a.K) a.costs[k] = cost + cc;  // POL:274-275
    } else {
#pragma unroll
      for (int q = 0; q < SS; ++q) wsb[q * 32] = s[q];
      wsb[SS * 32] = cost, wsb[(SS + 1) * 32] = cc, wsb[(SS + 2) * 32] = trig_valid ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < 4 * NCARS; ++q) wsb[(SS + 3 + q) * 32] = trig[q];
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release(done + b, u + 1);
    }
  }
  if (a.warp_cycles && lane == 0)
    a.warp_cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = clock64() - t_begin;
}

int rollout_queue_fields(int n_cars) { return n_cars == 1 ? 8 + 3 + 4 : 0; }

// returns 1 if launched, 0 if the configuration is not covered (caller falls back to launch_rollout_car)
int launch_rollout_car_queue(const CarEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t st) {
  if (env.n_cars != 1 || !a.ws || !a.ws_sync || a.unit_len < 1 || a.queue_ctas < 1) return 0;
  const size_t smem = sizeof(double) * 3 * env.n_trk;
  if (smem > 40 * 1024)
    cudaFuncSetAttribute(rollout_car_queue_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nb = (a.K + 31) / 32;
  cudaMemsetAsync(a.ws_sync, 0, sizeof(int) * (size_t)(nb + 1), st);
  rollout_car_queue_kernel<1, 3><<<a.queue_ctas, block, smem, st>>>(env, a, stop);
  return 1;
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

template <int MODE, int STAGE>
static void launch_rollout_car_v(const CarEnvArgs &env, const RolloutArgs &a, int block, const int *stop,
                                 cudaStream_t st) {
  const int grid = (a.K + block - 1) / block;
  const size_t smem = sizeof(double) * 3 * env.n_trk;
#define MPOPIS_LAUNCH(N)                                                                               \
  case N:                                                                                              \
    if (smem > 40 * 1024)                                                                              \
      cudaFuncSetAttribute(rollout_car_kernel<N, MODE, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)smem);                                                                 \
    rollout_car_kernel<N, MODE, STAGE><<<grid, block, smem, st>>>(env, a, stop);                      \
    break;
  switch (env.n_cars) {
    MPOPIS_LAUNCH(1)
    MPOPIS_LAUNCH(2)
    MPOPIS_LAUNCH(3)
    MPOPIS_LAUNCH(4)
  }
  if constexpr ((MODE == 0 || MODE == 3) && STAGE == 0) {  // 5..8 cars: production variants, register prefetch only
    switch (env.n_cars) {
      MPOPIS_LAUNCH(5)
      MPOPIS_LAUNCH(6)
      MPOPIS_LAUNCH(7)
      MPOPIS_LAUNCH(8)
    }
  }
#undef MPOPIS_LAUNCH
}

// stage: 0 = register prefetch of the next step's noise, 1 = TMA bulk copies into a per-warp shared-memory ring
// (variant 3, up to 4 cars, the rows of E 256-byte aligned per warp — guaranteed by ldk % 32 == 0)
void launch_rollout_car(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, int stage,
                        const int *stop, cudaStream_t st) {
  if (variant == 3 && stage == 1 && env.n_cars <= 4) launch_rollout_car_v<3, 1>(env, a, block, stop, st);
  else if (variant == 3) launch_rollout_car_v<3, 0>(env, a, block, stop, st);
  else if (variant == 0 || env.n_cars > 4) launch_rollout_car_v<0, 0>(env, a, block, stop, st);
  else if (variant == 1) launch_rollout_car_v<1, 0>(env, a, block, stop, st);
  else launch_rollout_car_v<2, 0>(env, a, block, stop, st);
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
  if (variant == 0) MPOPIS_ES(0) else if (variant == 1) MPOPIS_ES(1) else if (variant == 3) MPOPIS_ES(3) else MPOPIS_ES(2)
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
  if (variant == 0) MPOPIS_ER(0) else if (variant == 1) MPOPIS_ER(1) else if (variant == 3) MPOPIS_ER(3) else MPOPIS_ER(2)
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
