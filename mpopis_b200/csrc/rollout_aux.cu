// rollout_aux.cu — the non-production instantiation of the rollout kernel (MODE 1 "literal": the reference's libm
// call sequence, the parity anchor) and the single-step parity surfaces (track query, env step, reward).
// Split from rollout.cu only to halve the build's critical path; see rollout.cu for the description of the variants.
#include "rollout_kernels.cuh"

namespace mpopis {

void launch_rollout_car_aux(const CarEnvArgs &env, const RolloutArgs &a, int variant, int block, const int *stop,
                            cudaStream_t st) {
  (void)variant;  // only the literal variant lives here since the first fast cut (MODE 2) was retired
  launch_rollout_car_v<1, 0>(env, a, block, stop, st);
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
  if (variant == 0) MPOPIS_ES(0) else if (variant == 1) MPOPIS_ES(1) else MPOPIS_ES(3)
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
  if (variant == 0) MPOPIS_ER(0) else if (variant == 1) MPOPIS_ER(1) else MPOPIS_ER(3)
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
