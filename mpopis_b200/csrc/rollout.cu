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
#include "rollout_kernels.cuh"

namespace mpopis {

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

// stage: 0 = register prefetch of the next step's noise, 1 = TMA bulk copies into a per-warp shared-memory ring
// (variant 3, up to 4 cars, the rows of E 256-byte aligned per warp — guaranteed by ldk % 32 == 0)
void launch_rollout_car(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, int stage,
                        const int *stop, cudaStream_t st) {
  if (variant == 3 && stage == 1 && env.n_cars <= 4) launch_rollout_car_v<3, 1>(env, a, block, stop, st);
  else if (variant == 3) launch_rollout_car_v<3, 0>(env, a, block, stop, st);
  else if (variant == 0 || env.n_cars > 4) launch_rollout_car_v<0, 0>(env, a, block, stop, st);
  else launch_rollout_car_aux(env, a, variant, block, stop, st);  // variants 1 and 2 live in rollout_aux.cu
}

void launch_rollout_mc(const McEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t st) {
  rollout_mc_kernel<<<(a.K + block - 1) / block, block, 0, st>>>(env, a, stop);
}

}  // namespace mpopis
