// rollout_kernels.cuh — the rollout kernel template and its per-(variant, staging) launcher, shared by the two
// translation units that instantiate it (rollout.cu: the production variants 0 and 3; rollout_aux.cu: the literal and
// first-cut variants 1 and 2 plus the single-step parity kernels) so that nvcc compiles them in parallel.
#pragma once
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
      const double dd = MODE == 3 ? sqrt_fast(dx * dx + dy * dy) : sqrt(dx * dx + dy * dy);
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

template <int MODE, int STAGE>
static inline void launch_rollout_car_v(const CarEnvArgs &env, const RolloutArgs &a, int block, const int *stop,
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

// rollout_aux.cu
void launch_rollout_car_aux(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, const int *stop,
                            cudaStream_t st);

}  // namespace mpopis
