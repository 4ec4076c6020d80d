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
  else launch_rollout_car_aux(env, a, variant, block, stop, st);  // the literal variant lives in rollout_aux.cu
}

void launch_rollout_mc(const McEnvArgs &env, const RolloutArgs &a, int block, const int *stop, cudaStream_t st) {
  rollout_mc_kernel<<<(a.K + block - 1) / block, block, 0, st>>>(env, a, stop);
}

}  // namespace mpopis
