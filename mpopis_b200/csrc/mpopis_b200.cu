// mpopis_b200.cu — handle, AIS-loop orchestration and the C-ABI of include/mpopis_b200.h.
//
// One handle = one policy object on one GPU (optionally one shard of a K-sharded policy). A control
// step (`plan`) is a fixed, stream-ordered sequence of kernel launches with no host synchronisation
// inside: the CE/CMA early stop (POL:459-461, 567-569) is a device flag every later kernel checks,
// Cholesky failure is a device flag read back with the result. Host <-> device traffic per step is
// state (ss) + U (cs) in, control (as) + U (cs) + three ints out.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "comm.cuh"
#include "engine.cuh"

using namespace mpopis;

namespace {

thread_local char g_err[1024] = "";
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(MPOPIS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <class T>
int dalloc(T **p, size_t n) {
  CU(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  CU(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
  return 0;
}

// scratch of the parity-surface entry points: freed on every return path
template <class T>
struct DevBuf {
  T *p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t n) { return dalloc(&p, n); }
  operator T *() const { return p; }
};

}  // namespace

struct mpopis_handle {
  mpopis_cfg_t cfg{};
  int K = 0, Kloc = 0, T = 0, N = 0, as = 0, cs = 0, ss = 0, m_elite = 0;
  long long k0 = 0, ldk = 0, ldm = 0;
  int dev = 0, world = 1, rank = 0;
  cudaStream_t st = nullptr, st2 = nullptr;  // main stream; side stream for the next iteration's normals
  cudaEvent_t ev_z_free = nullptr, ev_z_ready = nullptr, ev_q_fork = nullptr, ev_q_join = nullptr;
  // injected noise / uniforms travel through a two-slot pinned ring (see stage_h2d)
  double *h_pin[2] = {nullptr, nullptr};
  size_t h_pin_bytes = 0;
  cudaEvent_t ev_pin[2] = {nullptr, nullptr};
  unsigned pin_next = 0;
  Comm comm{};
  bool env_set = false, cma_set = false;
  CarEnvArgs car{};
  McEnvArgs mc{};
  double gamma = 0.0;
  uint64_t seed = 0;
  long long step = 0;
  int rollout_spin = -1;  // split kernel hand-over: -1 auto, 0 mbarrier, 1 spin on shared-memory counters
  int rollout_variant = 6, rollout_variant_used = -1, rollout_block = 64, rollout_stage = 0, coop_max = 1, sort_max = 1, sel_max = 1;
  bool moments_small = true;  // single-CTA moment chain for small n ("moments_small" option, A/B)
  int sigma_bs = 0;  // block size of the initial Σ (as => block diagonal, cs => dense)
  bool L0_valid = false;
  mpopis_cma_t cma{};
  long long launches = 0;
  // device state
  double *d_trk = nullptr, *d_state = nullptr, *d_U_orig = nullptr, *d_U_cur = nullptr, *d_U_next = nullptr,
         *d_control = nullptr, *d_Sigma0 = nullptr, *d_Sigma = nullptr, *d_Lt = nullptr, *d_Lt0 = nullptr,
         *d_cholW = nullptr, *d_bvec = nullptr, *d_Z = nullptr, *d_E = nullptr, *d_stage = nullptr,
         *d_costs = nullptr, *d_w = nullptr, *d_sorted = nullptr, *d_X = nullptr, *d_mask = nullptr,
         *d_part = nullptr, *d_sums = nullptr, *d_mu = nullptr, *d_P = nullptr, *d_Sraw = nullptr,
         *d_qpart = nullptr, *d_q = nullptr, *d_lambda = nullptr, *d_traj = nullptr, *d_u = nullptr,
         *d_cdf = nullptr, *d_wcnt = nullptr, *d_ws = nullptr, *d_sigma = nullptr, *d_psig = nullptr,
         *d_pSig = nullptr, *d_dw = nullptr, *d_C = nullptr, *d_ns = nullptr, *d_reward = nullptr,
         *d_ones = nullptr;
  int num_sms = 0;
  // CUDA graph of one control step (plan without injected noise): the launch sequence is static — early stop and
  // Cholesky failure are device flags, the Philox control-step counter lives in device memory (d_step)
  cudaGraphExec_t gexec = nullptr;
  bool graph_enabled = true, capturing = false;
  long long graph_launches = 0;
  unsigned *d_step = nullptr;
  bool Lt_ready = false, small_fused = true;  // small_adapt.cu formed the next iteration's factor already ("ce_small_fused")
  bool use_select = false, cov_pending = false, fuse_cov = true;  // fuse_cov: cov_finalize folded into the Cholesky kernel  // :cemppi: select.cu path (sharded, or K above the single-CTA sort)
  long long *d_env_t = nullptr, *d_warp_cycles = nullptr;  // d_warp_cycles: "rollout_profile" option
  unsigned long long *d_keys_a = nullptr, *d_keys_b = nullptr;
  int *d_order = nullptr, *d_vals_b = nullptr, *d_hist = nullptr, *d_counts = nullptr, *d_flags = nullptr;
  unsigned char *d_done = nullptr;
  uint4 *d_lut = nullptr;
  // :cemppi selection without a sort (select.cu): workspace, bucket min/max keys, local elite ids, debug outputs
  void *d_sel_ws = nullptr;
  unsigned long long *d_bmin = nullptr, *d_bmax = nullptr;
  long long nb_cap = 0;
  int *d_eidx = nullptr, *d_mloc = nullptr;
  double *d_tau = nullptr, *d_bvec2 = nullptr;  // d_bvec2: 1/σ_i of the :ss standardisation
  size_t part_doubles = 0;
  // d_flags: [0] stop, [1] its, [2] info
  int *stop() { return d_flags; }
  int *its() { return d_flags + 1; }
  int *info() { return d_flags + 2; }
  // external-simulator seam (MPOPIS_ENV_EXTERNAL): callback of the running plan_external, pinned staging, bounds
  mpopis_rollout_fn ext_fn = nullptr;
  void *ext_user = nullptr;
  double *h_ext_controls = nullptr, *h_ext_costs = nullptr, *d_ext_cc = nullptr, *d_ext_in = nullptr,
         *d_ext_bounds = nullptr;
  int *h_ext_stop = nullptr;
  // pinned staging
  double *h_in = nullptr, *h_out = nullptr;
  int *h_flags = nullptr;
  // timing
  std::vector<cudaEvent_t> ev;  // 2 per iteration + 2 total
  int last_its_launched = 0;
  double last_rollout_ms = 0, last_total_ms = 0;
  bool timing_valid = false;
  // optional phase tracer (MPOPIS_TRACE=1): CUDA events at phase boundaries of the last plan, printed to stderr
  bool trace = false;
  std::vector<std::pair<const char *, cudaEvent_t>> marks;
  std::vector<cudaEvent_t> mark_pool;
};

namespace {

// timing events inside a captured control step are EXTERNAL event-record nodes: they can be timed after a replay
cudaError_t record(mpopis_t *h, cudaEvent_t e) {
  return h->capturing ? cudaEventRecordWithFlags(e, h->st, cudaEventRecordExternal) : cudaEventRecord(e, h->st);
}

void mark(mpopis_t *h, const char *name) {
  if (!h->trace) return;
  cudaEvent_t e;
  if (h->marks.size() < h->mark_pool.size()) e = h->mark_pool[h->marks.size()];
  else {
    cudaEventCreate(&e);
    h->mark_pool.push_back(e);
  }
  cudaEventRecord(e, h->st);
  h->marks.emplace_back(name, e);
}

void dump_marks(mpopis_t *h) {
  if (!h->trace || h->marks.size() < 2) return;
  std::vector<std::pair<std::string, double>> agg;
  double tot = 0;
  for (size_t i = 1; i < h->marks.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->marks[i - 1].second, h->marks[i].second);
    tot += ms;
    bool found = false;
    for (auto &a : agg)
      if (a.first == h->marks[i].first) a.second += ms, found = true;
    if (!found) agg.emplace_back(h->marks[i].first, ms);
  }
  fprintf(stderr, "[mpopis trace rank %d] total %.3f ms:", h->rank, tot);
  for (auto &a : agg) fprintf(stderr, " %s=%.3f", a.first.c_str(), a.second);
  fprintf(stderr, "\n");
}

int set_device(mpopis_t *h) {
  CU(cudaSetDevice(h->dev));
  return 0;
}

// Host -> device copy of caller-owned (pageable) memory on h->st WITHOUT handing the driver a pageable pointer: a pageable
// cudaMemcpyAsync blocks the calling thread inside the driver until the stream has drained, and with virtual ranks
// sharing one context (loop-back group, peer-memory collectives) the thread it starves is the one that still has to
// launch the kernel this rank's stream is spinning on. The bytes go through a two-slot pinned ring; a slot is reused
// only after ITS previous copy has completed (an event wait on the copy alone, not on the stream).
int ensure_pin(mpopis_t *h, size_t bytes) {
  if (bytes <= h->h_pin_bytes) return 0;
  CU(cudaStreamSynchronize(h->st));
  for (int b = 0; b < 2; ++b) {
    if (h->h_pin[b]) cudaFreeHost(h->h_pin[b]), h->h_pin[b] = nullptr;
    CU(cudaMallocHost((void **)&h->h_pin[b], bytes));
    if (!h->ev_pin[b]) CU(cudaEventCreateWithFlags(&h->ev_pin[b], cudaEventDisableTiming));
  }
  h->h_pin_bytes = bytes;
  return 0;
}
int stage_h2d(mpopis_t *h, void *dst_dev, const void *src_host, size_t bytes) {
  if (int rc = ensure_pin(h, bytes)) return rc;
  const int b = (int)(h->pin_next++ & 1u);
  CU(cudaEventSynchronize(h->ev_pin[b]));  // a never-recorded event is complete
  memcpy(h->h_pin[b], src_host, bytes);
  CU(cudaMemcpyAsync(dst_dev, h->h_pin[b], bytes, cudaMemcpyHostToDevice, h->st));
  CU(cudaEventRecord(h->ev_pin[b], h->st));
  return 0;
}

bool is_g_family(int pol) { return pol != MPOPIS_POLICY_MPPI; }
bool adapts_sigma(int pol) {
  return pol == MPOPIS_POLICY_CEMPPI || pol == MPOPIS_POLICY_CMAMPPI || pol == MPOPIS_POLICY_MUSIGMAAISMPPI ||
         pol == MPOPIS_POLICY_PMCMPPI;
}

int allreduce_sum(mpopis_t *h, double *buf, size_t n) {
  if (h->world == 1) return 0;
  if (comm_allreduce_sum(h->comm, buf, n, h->st)) return fail(MPOPIS_ERR_NCCL, "all-reduce: %s", comm_error());
  return 0;
}
// in-place all-gather of the trajectory costs (8 B per sample): every rank then performs the identical selection
int allgather_costs(mpopis_t *h) {
  if (h->world == 1) return 0;
  if (comm_allgather_f64(h->comm, h->d_costs, (size_t)h->Kloc, h->st))
    return fail(MPOPIS_ERR_NCCL, "all-gather: %s", comm_error());
  return 0;
}

// (weighted) mean [+ covariance] of the columns of X ([cs][ld], n local columns) — G5.
// w: per-local-column weights or nullptr. Adds the mean to U_cur when update_U (scaled by *scale_dev).
int moments(mpopis_t *h, const double *X, long long ld, int n, const double *w, bool want_cov, int corrected,
            int method, double ridge, bool update_U, const double *scale_dev, double *Sigma_out,
            const int *n_dev = nullptr, const int *cols = nullptr) {
  const int cs = h->cs;
  const int *stop = h->stop();
  if (h->world == 1 && n <= MOMENTS_SMALL_MAX && cs <= 512 && !n_dev && h->moments_small) {
    // the reference's own problem sizes: one launch instead of eight
    launch_moments_small(X, ld, cs, n, w, cols, want_cov, corrected, method, ridge, h->d_mu, update_U ? h->d_U_cur : nullptr,
                         scale_dev, h->d_sums, h->d_Sraw, Sigma_out, h->d_lambda, stop, h->st);
    h->launches += 1;
    mark(h, "moments.small");
    return 0;
  }
  // The shrinkage estimators reach this chain only on one GPU (the sharded :cemppi update is ce_adapt below; the other
  // policies fit with SimpleCovariance / mean_and_cov), so the shrinkage statistic needs no collective here.
  const int nch = rowsum_nchunks(n);
  const bool shrink = want_cov && (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS);
  if (shrink && h->world > 1) return fail(MPOPIS_ERR_BAD_ARG, "shrinkage estimators are sharded through ce_adapt only");
  launch_rowsum_partial(X, ld, cs, n, w, h->d_part, stop, h->st, n_dev, 0);
  launch_reduce_partials(h->d_part, nch, cs + 1, h->d_sums, stop, h->st, cs + 1);
  h->launches += 2;
  mark(h, "mean.local");
  if (int rc = allreduce_sum(h, h->d_sums, cs + 1)) return rc;
  mark(h, "mean.coll");
  launch_finalize_mean(h->d_sums, cs, h->d_mu, update_U ? h->d_U_cur : nullptr, scale_dev, stop, h->st);
  h->launches += 1;
  mark(h, "mean");
  if (!want_cov) return 0;
  launch_syrk_partial(X, ld, cs, n, w, h->d_mu, h->d_P, stop, h->st, n_dev);
  launch_scatter_reduce(h->d_P, syrk_nchunks(n), cs, h->d_Sraw, stop, h->st);
  h->launches += 2;
  mark(h, "scatter.local");
  if (int rc = allreduce_sum(h, h->d_Sraw, (size_t)cs * cs)) return rc;
  mark(h, "scatter.coll");
  int nq = 0;
  if (shrink) {
    launch_shrink_q_partial(X, ld, cs, n, w, h->d_mu, h->d_Sraw, h->d_sums + cs, method == MPOPIS_SIGMA_SS, h->d_qpart,
                            stop, h->st, n_dev);
    launch_reduce_partials(h->d_qpart, shrink_q_nblocks(n), 1, h->d_q, stop, h->st);
    h->launches += 2;
    nq = 1;
    mark(h, "shrinkq");
  }
  launch_cov_finalize(h->d_Sraw, cs, h->d_sums + cs, corrected, method, h->d_q, nq, ridge, Sigma_out, h->d_lambda, stop,
                      h->st);
  h->launches += 1;
  mark(h, "covfin");
  return 0;
}

// The cross-entropy adaptation (POL:455-465) without a sort (select.cu): exact radix select of the m-th smallest
// (cost, sample id) on the (all-gathered) cost vector, the early-stop test on bucketed elite costs, the ids of the
// elites THIS shard owns, then mean / covariance of those columns. Sharded: every rank selects redundantly on identical
// input (no collective beyond the cost all-gather), the moments need two all-reduces:
//   [Σx | n | Σx²] (2cs + 1)  and  [scatter matrix | shrinkage statistic] (cs² + 1).
int ce_adapt(mpopis_t *h) {
  const int cs = h->cs, K = h->K, Kloc = h->Kloc, m = h->m_elite, method = h->cfg.sigma_est;
  int *stop = h->stop();
  cudaStream_t st = h->st;
  const cudaError_t e = (cudaError_t)launch_ce_select(h->d_costs, K, m, h->k0, Kloc, h->cfg.early_stop, h->d_sel_ws,
                                                      h->d_bmin, h->d_bmax, h->nb_cap, h->d_eidx, h->d_mloc, h->d_tau,
                                                      stop, stop, h->sel_max, st);
  if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  mark(h, "select");
  const int mmax = m < Kloc ? m : Kloc, nch = elite_gather_nchunks(mmax);
  const bool shrink = method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS, ss = method == MPOPIS_SIGMA_SS;
  launch_elite_gather_sums(h->d_E, h->ldk, cs, h->d_eidx, h->d_mloc, mmax, h->d_X, h->ldm, h->d_part, stop, st);
  h->launches += 2;
  if (h->world == 1) {
    launch_ce_sums(h->d_part, nch, cs, h->d_mloc, h->d_sums, 1, ss, h->d_mu, h->d_U_cur, h->d_bvec2, stop, st);
    h->launches += 1;
  } else {
    launch_ce_sums(h->d_part, nch, cs, h->d_mloc, h->d_sums, 0, ss, h->d_mu, h->d_U_cur, h->d_bvec2, stop, st);
    mark(h, "mean.local");
    if (int rc = allreduce_sum(h, h->d_sums, 2 * (size_t)cs + 1)) return rc;
    mark(h, "mean.coll");
    launch_ce_sums(nullptr, 0, cs, h->d_mloc, h->d_sums, 1, ss, h->d_mu, h->d_U_cur, h->d_bvec2, stop, st);
    h->launches += 2;
  }
  mark(h, "mean");
  double *qdst = h->d_Sraw + (size_t)cs * cs;  // travels with the scatter matrix
  if (shrink) {  // the shrinkage statistic needs only X, μ and 1/σ: it runs beside the scatter matrix on the side stream
    CU(cudaEventRecord(h->ev_q_fork, st));
    CU(cudaStreamWaitEvent(h->st2, h->ev_q_fork, 0));
    launch_shrink_q_partial(h->d_X, h->ldm, cs, mmax, nullptr, h->d_mu, h->d_Sraw, h->d_sums + cs, ss, h->d_qpart, stop,
                            h->st2, h->d_mloc, h->d_bvec2);
    launch_reduce_partials(h->d_qpart, shrink_q_nblocks(mmax), 1, qdst, stop, h->st2);
    CU(cudaEventRecord(h->ev_q_join, h->st2));
    h->launches += 2;
  }
  launch_syrk_partial(h->d_X, h->ldm, cs, mmax, nullptr, h->d_mu, h->d_P, stop, st, h->d_mloc);
  launch_scatter_reduce(h->d_P, syrk_nchunks(mmax), cs, h->d_Sraw, stop, st);
  h->launches += 2;
  if (shrink) CU(cudaStreamWaitEvent(st, h->ev_q_join, 0));
  mark(h, "scatter.local");
  if (int rc = allreduce_sum(h, h->d_Sraw, (size_t)cs * cs + (shrink ? 1 : 0))) return rc;
  mark(h, "scatter.coll");
  if (cs <= 160 && h->fuse_cov) {  // Σ′ = shrink(S) + 1e-8 I is formed inside the next iteration's factorisation
    h->cov_pending = true;
    return 0;
  }
  launch_cov_finalize(h->d_Sraw, cs, h->d_sums + cs, 0, method, qdst, shrink ? 1 : 0, 10e-9, h->d_Sigma, h->d_lambda,
                      stop, st);
  h->launches += 1;
  mark(h, "covfin");
  return 0;
}

// ---- external-simulator seam (EnvpoolEnv methods POL:148-184, 240-259; UTL:42-53, 103-121) ------------------
// controls[k + K*(r)] = clamp(pol.U[r] + E[r,k], lo[r % as], hi[r % as]) with r = a + as*t: the K x as x T column-major
// array get_model_controls(action_space, Vₖ, T) returns (UTL:42-53) — the device's own [cs][k] layout with pitch K.
// cc[k] = (γ U_origᵀ Σ⁻¹)·(Vₖ − U_orig) (POL:248); one thread per sample, coalesced in k.
__global__ void ext_controls_kernel(const double *__restrict__ E, long long ldk, const double *__restrict__ U_cur,
                                    const double *__restrict__ U_orig, const double *__restrict__ bvec,
                                    const double *__restrict__ bounds, int as, int cs, int K,
                                    double *__restrict__ controls, double *__restrict__ cc, const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double acc = 0.0;
  for (int r = 0; r < cs; ++r) {
    const double v = U_cur[r] + E[(size_t)r * ldk + k];  // POL:252
    if (bvec) acc += bvec[r] * (v - U_orig[r]);
    const int a = r % as;
    controls[(size_t)r * K + k] = fmin(fmax(v, bounds[a]), bounds[as + a]);
  }
  cc[k] = acc;
}

__global__ void ext_costs_kernel(const double *__restrict__ traj_cost, const double *__restrict__ cc, int K,
                                 double *__restrict__ costs, const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) costs[k] = traj_cost[k] + cc[k];  // POL:256
}

int ext_rollouts(mpopis_t *h, const double *U_cur, const double *U_orig, const double *bvec) {
  if (!h->ext_fn) return fail(MPOPIS_ERR_BAD_ARG, "external env: use mpopis_b200_plan_external");
  const int K = h->Kloc, cs = h->cs;
  cudaStream_t st = h->st;
  ext_controls_kernel<<<(K + 127) / 128, 128, 0, st>>>(h->d_E, h->ldk, U_cur, U_orig, bvec, h->d_ext_bounds, h->as, cs, K,
                                                       h->d_stage, h->d_ext_cc, h->stop());
  CU(cudaMemcpyAsync(h->h_ext_controls, h->d_stage, sizeof(double) * cs * K, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->h_ext_stop, h->stop(), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  h->launches += 1;
  if (*h->h_ext_stop) return 0;  // early stop (POL:459-461): this iteration is not executed
  if (h->ext_fn(h->ext_user, h->h_ext_controls, K, h->as, h->T, h->h_ext_costs) != 0)
    return fail(MPOPIS_ERR_BAD_ARG, "external rollout callback reported an error");
  CU(cudaMemcpyAsync(h->d_ext_in, h->h_ext_costs, sizeof(double) * K, cudaMemcpyHostToDevice, st));
  ext_costs_kernel<<<(K + 255) / 256, 256, 0, st>>>(h->d_ext_in, h->d_ext_cc, K, h->d_costs + h->k0, h->stop());
  h->launches += 1;
  CU(cudaGetLastError());
  return 0;
}

int launch_rollouts(mpopis_t *h, const double *U_cur, const double *U_orig, const double *bvec) {
  if (h->cfg.env == MPOPIS_ENV_EXTERNAL) return ext_rollouts(h, U_cur, U_orig, bvec);
  RolloutArgs a{};
  a.E = h->d_E, a.ldk = h->ldk, a.U = U_cur, a.U_orig = U_orig, a.bvec = bvec;
  a.state0 = h->d_state, a.env_t = h->d_env_t, a.costs = h->d_costs + h->k0;
  a.traj = h->cfg.log_trajectories ? h->d_traj : nullptr;
  a.K = h->Kloc, a.T = h->T;
  a.warp_cycles = h->d_warp_cycles;
  if (h->cfg.env == MPOPIS_ENV_CAR_RACING) {
    // variants 4 / 5 ("v5"): the warp-specialised kernel (rollout_split.cu; 1..3 cars, nsub = 10 — anything else falls
    // back to variant 3). 6 = automatic (the default): the wide split kernel while one launch keeps every rollout
    // resident (latency regime: K <= 28 416 one-car rollouts on 148 SMs), the thread-per-rollout kernel beyond.
    int variant = h->rollout_variant;
    if (variant == 6)
      variant = (h->cfg.n_cars <= rollout_split_max_cars() && h->Kloc <= rollout_split_capacity(h->cfg.n_cars, h->num_sms) &&
                 !h->rollout_stage) ? 5 : 3;
    const int ctas = (h->Kloc + 63) / 64;
    // default: the lone-CTA flavour (255 registers, spin hand-over) while at most two CTAs share an SM
    const int spin = h->rollout_spin < 0 ? ctas <= 2 * h->num_sms : h->rollout_spin;
    h->rollout_variant_used = variant;
    if (!(variant >= 4 && launch_rollout_car_split(h->car, a, variant == 5, spin, h->stop(), h->st)))
      h->rollout_variant_used = variant >= 4 ? 3 : variant, launch_rollout_car(h->car, a, variant >= 4 ? 3 : variant, h->rollout_block,
                         h->rollout_stage, h->stop(), h->st);
  } else
    launch_rollout_mc(h->mc, a, h->rollout_block, h->stop(), h->st);
  h->launches += 1;
  CU(cudaGetLastError());
  return 0;
}

int cma_update(mpopis_t *h, int n_iter);  // cma.cu-style section below

// factor the initial Σ once per set_sigma() — outside the (possibly captured) control step
void factor_sigma0(mpopis_t *h) {
  if (h->L0_valid) return;
  launch_chol(h->d_Sigma0, h->cs, nullptr, h->d_Lt0, h->d_cholW, h->info(), 1000, nullptr, h->st);
  h->launches += 1;
  h->L0_valid = true;
}

// The AIS loop + final control, entirely on h->st. Z_host: injected normals (cs x K x N col-major)
// or nullptr for the Philox generator; u_host: injected PMC uniforms (K x (N-1)) or nullptr.
int plan_core(mpopis_t *h, const double *Z_host, const double *u_host) {
  const int cs = h->cs, K = h->K, Kloc = h->Kloc, N = h->N, pol = h->cfg.policy;
  int *stop = h->stop();
  cudaStream_t st = h->st;
  if (!h->env_set) return fail(MPOPIS_ERR_BAD_ARG, "environment not set");
  if (pol == MPOPIS_POLICY_CMAMPPI && !h->cma_set) return fail(MPOPIS_ERR_BAD_ARG, "CMA constants not set");
  if (pol == MPOPIS_POLICY_CMAMPPI && N > 1 && (long long)cs * h->m_elite < K)  // δs[order[ii]] out of bounds
    return fail(MPOPIS_ERR_BAD_ARG, "BoundsError: CMA linear index exceeds cs*m_elite (POL:593)");
  if (pol == MPOPIS_POLICY_PMCMPPI && N > 1 && Z_host && !u_host)
    return fail(MPOPIS_ERR_BAD_ARG, "pmcmppi with injected noise needs resample_u");
  CU(record(h, h->ev[0]));
  h->Lt_ready = false;
  h->marks.clear();
  mark(h, "begin");
  CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 2, st));  // stop, its (info is sticky until read)
  CU(cudaMemcpyAsync(h->d_U_cur, h->d_U_orig, sizeof(double) * cs, cudaMemcpyDeviceToDevice, st));
  const bool adapt = adapts_sigma(pol) && N > 1;
  if (adapt) CU(cudaMemcpyAsync(h->d_Sigma, h->d_Sigma0, sizeof(double) * cs * cs, cudaMemcpyDeviceToDevice, st));
  if (pol == MPOPIS_POLICY_CMAMPPI) {
    launch_set_scalar(h->d_sigma, h->cma.sigma, st);  // a kernel, not a pageable H2D copy: graph-capturable
    CU(cudaMemsetAsync(h->d_psig, 0, sizeof(double) * cs, st));
    CU(cudaMemsetAsync(h->d_pSig, 0, sizeof(double) * cs, st));
  }
  const double *Lt = h->d_Lt0;
  int bs = h->sigma_bs;
  const double *bvec = nullptr;
  for (int n = 0; n < N; ++n) {
    launch_iter_begin(stop, h->its(), h->d_flags + 3, st);
    h->launches += 1;
    // --- proposal factor L of Σ′ (POL:447; CMA samples from σ²Σ, POL:550-554) ---
    const bool cma_scaled = pol == MPOPIS_POLICY_CMAMPPI && N > 1;
    if ((adapt && n > 0) || cma_scaled) {
      if (h->Lt_ready) {
        // the fused small-size adaptation (small_adapt.cu) has factored Σ′ already
      } else {
        // ce_adapt leaves Σ′ un-finalised when the shrinkage + ridge can be folded into the factorisation (one launch)
        if (!(h->cov_pending &&
              launch_chol_cov(h->d_Sraw, cs, h->d_sums + cs, 0, h->cfg.sigma_est, h->d_Sraw + (size_t)cs * cs, 10e-9,
                              h->d_Sigma, h->d_Lt, h->d_lambda, h->info(), n + 1, stop, st)))
          launch_chol(adapt && n > 0 ? h->d_Sigma : h->d_Sigma0, cs, cma_scaled ? h->d_sigma : nullptr, h->d_Lt,
                      h->d_cholW, h->info(), n + 1, stop, st);
        h->launches += 1;
      }
      h->cov_pending = false, h->Lt_ready = false;
      Lt = h->d_Lt;
      if (n > 0) bs = cs;
    }
    if (h->gamma != 0.0 && (n == 0 || Lt == h->d_Lt)) {  // b = γ Σ′⁻¹ U_orig (POL:449 + POL:272)
      launch_chol_solve(Lt, cs, h->d_U_orig, h->gamma, h->d_bvec, stop, st);
      h->launches += 1;
      bvec = h->d_bvec;
    }
    mark(h, "chol");
    // --- Z, E = L Z (POL:448) ---
    if (Z_host) {
      const double *src = Z_host + ((size_t)n * K + h->k0) * cs;
      if (int rc = stage_h2d(h, h->d_stage, src, sizeof(double) * cs * Kloc)) return rc;
      launch_transpose_in(h->d_stage, h->d_Z, cs, Kloc, h->ldk, st);
    } else if (n == 0) {
      launch_philox_normals(h->d_Z, h->ldk, cs, Kloc, h->k0, h->seed, 0u, h->d_step, 0u, stop, st);
    } else {
      CU(cudaStreamWaitEvent(st, h->ev_z_ready, 0));  // Z of this iteration was drawn on the side stream
    }
    launch_apply_L(Lt, cs, bs, h->d_Z, h->d_E, h->ldk, Kloc, stop, st);
    h->launches += 2;
    mark(h, "sample");
    // --- rollouts (POL:452 -> POL:261-278) ---
    CU(record(h, h->ev[2 + 2 * n]));
    if (int rc = launch_rollouts(h, h->d_U_cur, h->d_U_orig, bvec)) return rc;
    CU(record(h, h->ev[3 + 2 * n]));
    if (!Z_host && n + 1 < N) {
      // The next iteration's normals depend on nothing but (seed, step, n+1): draw them on the side stream
      // while the latency-bound adaptation kernels (sort passes, Cholesky, moment finalisation) leave most
      // SMs idle. They are released only AFTER the rollout kernel: both are FP64-issue bound, and letting
      // them overlap slowed the rollouts by more than the 36 µs it hid (measured, profiles/README.md).
      CU(cudaEventRecord(h->ev_z_free, st));
      CU(cudaStreamWaitEvent(h->st2, h->ev_z_free, 0));
      launch_philox_normals(h->d_Z, h->ldk, cs, Kloc, h->k0, h->seed, 0u, h->d_step, (uint32_t)(n + 1), stop,
                            h->st2);
      CU(cudaEventRecord(h->ev_z_ready, h->st2));
    }
    if (int rc = allgather_costs(h)) return rc;
    if (n == N - 1) break;
    mark(h, "rollout+gather");
    // --- adaptation (the `if n < N` blocks) ---
    switch (pol) {
      case MPOPIS_POLICY_IMPPI:
      case MPOPIS_POLICY_MUAISMPPI:
      case MPOPIS_POLICY_MUSIGMAAISMPPI: {  // POL:361-365, 659-663, 729-734
        const double lam = pol == MPOPIS_POLICY_IMPPI ? h->cfg.lambda : h->cfg.lambda_ais;
        h->launches += launch_weights(h->d_costs, K, lam, h->d_w, h->d_ones, stop, st);
        const bool cov = pol == MPOPIS_POLICY_MUSIGMAAISMPPI;
        if (int rc = moments(h, h->d_E, h->ldk, Kloc, h->d_w + h->k0, cov, 0, MPOPIS_SIGMA_MLE, 10e-9, true,
                             nullptr, h->d_Sigma))
          return rc;
        break;
      }
      case MPOPIS_POLICY_PMCMPPI: {  // POL:802-809
        h->launches += launch_weights(h->d_costs, K, h->cfg.lambda_ais, h->d_w, h->d_ones, stop, st) - 1;
        if (u_host) {
          if (int rc = stage_h2d(h, h->d_u, u_host + (size_t)n * K, sizeof(double) * K)) return rc;
        } else launch_philox_uniforms(h->d_u, K, h->seed, 0u, h->d_step, (uint32_t)n, stop, st);
        launch_pmc_counts(h->d_w, K, h->d_u, h->d_cdf, h->d_counts, h->k0, Kloc, h->d_wcnt, stop, st);
        h->launches += 5;
        if (int rc = moments(h, h->d_E, h->ldk, Kloc, h->d_wcnt, true, 1, MPOPIS_SIGMA_MLE, 10e-9, true, nullptr,
                             h->d_Sigma))
          return rc;
        break;
      }
      case MPOPIS_POLICY_CEMPPI:
      case MPOPIS_POLICY_CMAMPPI: {  // POL:455-465, 563-599
        const int m = h->m_elite;
        if (pol == MPOPIS_POLICY_CEMPPI && h->use_select) {  // selection instead of a sort (select.cu)
          if (int rc = ce_adapt(h)) return rc;
          break;
        }
        if (pol == MPOPIS_POLICY_CEMPPI && h->world == 1 && h->small_fused && h->moments_small &&
            launch_ce_small_adapt(h->d_costs, K, m, h->cfg.early_stop, h->d_E, h->ldk, cs, h->cfg.sigma_est, 10e-9,
                                  h->d_keys_a, h->d_order, h->d_mu, h->d_U_cur, h->d_sums, h->d_Sigma, h->d_Lt, h->d_lambda,
                                  h->info(), n + 2, stop, st)) {
          h->Lt_ready = true;  // the factor of iteration n + 1 (failure tag n + 2, as launch_chol would report it)
          h->launches += 1;
          mark(h, "adapt.small");
          break;
        }
        {  // order = sortperm(costs) + the elite early-stop test (POL:455-461, 563-569)
          const cudaError_t e = (cudaError_t)launch_sortperm(h->d_costs, K, h->d_keys_a, h->d_keys_b, h->d_order,
                                                             h->d_vals_b, m, h->cfg.early_stop, stop, stop,
                                                             h->sort_max, st);
          if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
        }
        h->launches += sort_launches(K);
        if (pol == MPOPIS_POLICY_CEMPPI && h->world == 1 && m <= MOMENTS_SMALL_MAX && cs <= 512 && h->moments_small) {
          // small elite set: the single-CTA moment kernel reads the elite columns through `order` (no gather)
          mark(h, "select");
          if (int rc = moments(h, h->d_E, h->ldk, m, nullptr, true, 0, h->cfg.sigma_est, 10e-9, true, nullptr, h->d_Sigma,
                               nullptr, h->d_order))
            return rc;
          break;
        }
        launch_gather_cols(h->d_E, h->ldk, cs, h->d_order, m, h->k0, Kloc, h->d_X, h->ldm,
                           h->world > 1 ? h->d_mask : nullptr, stop, st);
        h->launches += 1;
        mark(h, "select");
        if (pol == MPOPIS_POLICY_CEMPPI) {
          if (int rc = moments(h, h->d_X, h->ldm, m, h->world > 1 ? h->d_mask : nullptr, true, 0,
                               h->cfg.sigma_est, 10e-9, true, nullptr, h->d_Sigma))
            return rc;
        } else {
          if (int rc = cma_update(h, n + 1)) return rc;
        }
        break;
      }
      default: break;
    }
  }
  h->last_its_launched = N;
  // --- final weights (always λ: POL:313,367,470,604,665,736,811), weighted noise, control, roll ---
  h->launches += launch_weights(h->d_costs, K, h->cfg.lambda, h->d_w, h->d_ones, nullptr, st);
  launch_rowsum_partial(h->d_E, h->ldk, cs, Kloc, h->d_w + h->k0, h->d_part, nullptr, st);
  launch_reduce_partials(h->d_part, rowsum_nchunks(Kloc), cs + 1, h->d_sums, nullptr, st);
  h->launches += 2;
  if (int rc = allreduce_sum(h, h->d_sums, cs + 1)) return rc;
  launch_finalize_control(h->d_sums, h->d_U_orig, h->d_U_cur, cs, h->as, h->T, h->d_U_next, h->d_control,
                          h->cfg.env == MPOPIS_ENV_EXTERNAL ? h->d_ext_bounds : nullptr, h->d_step, st);
  h->launches += 1;
  mark(h, "final");
  CU(record(h, h->ev[1]));
  CU(cudaGetLastError());
  h->step += 1;
  h->timing_valid = false;
  return 0;
}

int finish_timing(mpopis_t *h) {
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  h->last_total_ms = ms;
  double r = 0;
  for (int n = 0; n < h->last_its_launched; ++n) {
    CU(cudaEventElapsedTime(&ms, h->ev[2 + 2 * n], h->ev[3 + 2 * n]));
    r += ms;
  }
  h->last_rollout_ms = r;
  h->timing_valid = true;
  dump_marks(h);
  return 0;
}

void drop_graph(mpopis_t *h) {
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  h->gexec = nullptr;
}

// One control step on the engine's own Philox stream: replayed from a CUDA graph when the sequence is capturable
// (no injected noise, no host callback, NCCL or no communicator, tracer off), launched kernel by kernel otherwise.
// Any capture/instantiate failure falls back to eager launches for the lifetime of the handle.
int plan_step(mpopis_t *h) {
  factor_sigma0(h);
  const bool can = h->graph_enabled && !h->trace && !h->comm.host_synchronous() && h->cfg.env != MPOPIS_ENV_EXTERNAL;
  if (!can) return plan_core(h, nullptr, nullptr);
  if (!h->gexec) {
    const long long l0 = h->launches, step0 = h->step;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      h->graph_enabled = false;
      return plan_core(h, nullptr, nullptr);
    }
    h->capturing = true;
    const int rc = plan_core(h, nullptr, nullptr);
    h->capturing = false;
    const cudaError_t e = cudaStreamEndCapture(h->st, &g);
    h->graph_launches = h->launches - l0;
    h->launches = l0, h->step = step0;  // nothing ran yet
    const bool bad = rc || e != cudaSuccess || !g || cudaGraphInstantiate(&h->gexec, g, 0) != cudaSuccess ||
                     cudaGraphUpload(h->gexec, h->st) != cudaSuccess;
    // virtual ranks share one context: instantiation / upload may synchronise the device, so nobody launches (and
    // starts spinning on a peer) before every rank of the loop-back group is through it
    if (comm_host_barrier(h->comm)) return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
    if (bad) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      if (h->gexec) cudaGraphExecDestroy(h->gexec);
      h->gexec = nullptr, h->graph_enabled = false;
      if (getenv("MPOPIS_GRAPH_VERBOSE")) fprintf(stderr, "[mpopis] graph capture failed (rc=%d, %s): eager launches\n", rc, cudaGetErrorString(e));
      return plan_core(h, nullptr, nullptr);
    }
    cudaGraphDestroy(g);
  }
  CU(cudaGraphLaunch(h->gexec, h->st));
  h->launches += h->graph_launches;
  h->last_its_launched = h->N;
  h->step += 1;
  h->timing_valid = false;
  return 0;
}

int check_info(mpopis_t *h, int info) {
  if (info == COMM_PEER_TIMEOUT)
    return fail(MPOPIS_ERR_NCCL, "a peer-memory collective waited 30 s for a rank that never arrived (rank %d of %d)",
                h->rank, h->world);
  if (info != 0)
    return fail(MPOPIS_ERR_NOT_PD, "PosDefException: matrix is not positive definite; Cholesky factorization failed (%s)",
                info >= 1000 ? "initial Σ" : (std::string("AIS iteration ") + std::to_string(info)).c_str());
  return 0;
}

// CMA-ES adaptation (POL:571-599): δw row sums, the linear-index gather for the scalar "rank-μ"
// term, Σ^-0.5 (one cooperative kernel) and the vector-sized updates (one CTA).
int cma_update(mpopis_t *h, int n_iter) {
  const int cs = h->cs, m = h->m_elite, K = h->K;
  const int *stop = h->stop();
  launch_rowsum_partial(h->d_X, h->ldm, cs, m, h->d_ws, h->d_part, stop, h->st);
  launch_reduce_partials(h->d_part, rowsum_nchunks(m), cs + 1, h->d_sums, stop, h->st);
  if (int rc = allreduce_sum(h, h->d_sums, cs + 1)) return rc;
  launch_cma_lin_gather(h->d_X, h->ldm, cs, h->d_order, K, h->d_cdf, stop, h->st);
  if (int rc = allreduce_sum(h, h->d_cdf, K)) return rc;
  cudaError_t e = (cudaError_t)launch_inv_sqrt(h->d_Sigma, cs, h->d_C, h->d_ns, h->info(), 2000 + n_iter, stop,
                                               h->coop_max, h->st);
  if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  launch_cma_vec(h->d_sums, h->d_C, h->d_cdf, h->d_ws, K, cs, n_iter, h->cma, h->d_psig, h->d_pSig, h->d_sigma,
                 h->d_U_cur, h->d_Sigma, stop, h->st);
  h->launches += 5;
  return 0;
}

// Exact pruning table for the arg-min of within_track (TRK:71-73). For every cell of a uniform grid over
// the track's bounding box (+ margin) it lists the sampled points that can be nearest to SOME position in
// the cell: point i is kept iff  dmin(cell, p_i) <= min_j dmax(cell, p_j)  (with slack for rounding and for
// positions that the kernel's own cell-index arithmetic puts one ulp across a cell edge). The true arg-min
// always satisfies that inequality, so scanning the candidates in index order with the reference's
// arithmetic returns the same index as the full scan. Cells with more than 7 candidates (or positions
// outside the table) fall back to the full scan. Layout per cell: 8 x u16 = {count, idx0..idx6}.
void build_track_lut(const double *x, const double *y, const double *w, int n, std::vector<uint16_t> &cells,
                     double &x0, double &y0, double &cell, int &nx, int &ny) {
  double xmin = x[0], xmax = x[0], ymin = y[0], ymax = y[0], wmax = w[0];
  for (int i = 1; i < n; ++i) {
    xmin = std::min(xmin, x[i]), xmax = std::max(xmax, x[i]);
    ymin = std::min(ymin, y[i]), ymax = std::max(ymax, y[i]);
    wmax = std::max(wmax, w[i]);
  }
  const double margin = wmax + 40.0;
  cell = 2.0;
  x0 = xmin - margin, y0 = ymin - margin;
  const double wx = xmax + margin - x0, wy = ymax + margin - y0;
  while ((wx / cell) * (wy / cell) > 4.0e6 || (wx / cell) * (wy / cell) * n > 4.0e8) cell *= 2.0;
  nx = (int)std::ceil(wx / cell), ny = (int)std::ceil(wy / cell);
  cells.assign((size_t)nx * ny * 8, 0);
  const double hh = 0.5 * cell * (1.0 + 1e-9) + 1e-6;
  std::vector<double> dmin(n);
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
      const double cx = x0 + (ix + 0.5) * cell, cy = y0 + (iy + 0.5) * cell;
      double bound = INFINITY;
      for (int j = 0; j < n; ++j) {
        const double ax = std::fabs(cx - x[j]), ay = std::fabs(cy - y[j]);
        const double fx = ax + hh, fy = ay + hh, nxm = std::max(ax - hh, 0.0), nym = std::max(ay - hh, 0.0);
        bound = std::min(bound, std::sqrt(fx * fx + fy * fy));
        dmin[j] = std::sqrt(nxm * nxm + nym * nym);
      }
      bound = bound * (1.0 + 1e-9) + 1e-9;
      uint16_t *c = &cells[((size_t)iy * nx + ix) * 8];
      int cnt = 0;
      for (int j = 0; j < n && cnt <= 7; ++j)
        if (dmin[j] <= bound) {
          if (cnt < 7) c[1 + cnt] = (uint16_t)j;
          ++cnt;
        }
      c[0] = cnt <= 7 ? (uint16_t)cnt : (uint16_t)0xFFFF;
    }
}

int ensure_elite_capacity(mpopis_t *h, int m) {
  const long long ldm = ((long long)m + 31) / 32 * 32;
  if (h->d_X && ldm <= h->ldm) return 0;
  if (h->d_X) cudaFree(h->d_X), cudaFree(h->d_mask);
  h->ldm = ldm;
  if (int rc = dalloc(&h->d_X, (size_t)h->cs * ldm)) return rc;
  return dalloc(&h->d_mask, (size_t)ldm);
}

int upload_inputs(mpopis_t *h, const double *state, int64_t env_t, const double *U) {
  if (h->ss) memcpy(h->h_in, state, sizeof(double) * h->ss);
  memcpy(h->h_in + h->ss, U, sizeof(double) * h->cs);
  long long t = env_t;
  memcpy(h->h_in + h->ss + h->cs, &t, sizeof t);
  if (h->ss) CU(cudaMemcpyAsync(h->d_state, h->h_in, sizeof(double) * h->ss, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->d_U_orig, h->h_in + h->ss, sizeof(double) * h->cs, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->d_env_t, h->h_in + h->ss + h->cs, sizeof(long long), cudaMemcpyHostToDevice, h->st));
  return 0;
}

int download_outputs(mpopis_t *h, double *U_out, double *control_out, int32_t *its_out, double *state_out) {
  CU(cudaMemcpyAsync(h->h_out, h->d_control, sizeof(double) * h->as, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(h->h_out + h->as, h->d_U_next, sizeof(double) * h->cs, cudaMemcpyDeviceToHost, h->st));
  if (state_out)
    CU(cudaMemcpyAsync(h->h_out + h->as + h->cs, h->d_state, sizeof(double) * h->ss, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(h->h_flags, h->d_flags, sizeof(int) * 3, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  if (int rc = finish_timing(h)) return rc;
  if (int rc = check_info(h, h->h_flags[2])) return rc;
  if (control_out) memcpy(control_out, h->h_out, sizeof(double) * h->as);
  if (U_out) memcpy(U_out, h->h_out + h->as, sizeof(double) * h->cs);
  if (state_out) memcpy(state_out, h->h_out + h->as + h->cs, sizeof(double) * h->ss);
  if (its_out) *its_out = h->h_flags[1];
  return 0;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

int mpopis_b200_abi_version(void) { return MPOPIS_B200_ABI_VERSION; }
const char *mpopis_b200_last_error(void) { return g_err; }

int mpopis_b200_create(const mpopis_cfg_t *cfg, mpopis_t **out) {
  if (!cfg || !out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MPOPIS_B200_ABI_VERSION) return fail(MPOPIS_ERR_BAD_ARG, "abi version mismatch");
  if (cfg->policy < 0 || cfg->policy > MPOPIS_POLICY_PMCMPPI)
    return fail(MPOPIS_ERR_BAD_ARG, "No policy_type of %d", cfg->policy);
  if (cfg->num_samples < 1 || cfg->horizon < 1 || cfg->opt_its < 1 || cfg->num_samples > (1LL << 30))
    return fail(MPOPIS_ERR_BAD_ARG, "num_samples, horizon and opt_its must be positive");
  if (!(cfg->lambda > 0.0)) return fail(MPOPIS_ERR_BAD_ARG, "λ must be positive");
  if ((cfg->policy == MPOPIS_POLICY_MUAISMPPI || cfg->policy == MPOPIS_POLICY_MUSIGMAAISMPPI ||
       cfg->policy == MPOPIS_POLICY_PMCMPPI) && cfg->opt_its > 1 && !(cfg->lambda_ais > 0.0))
    return fail(MPOPIS_ERR_BAD_ARG, "λ_ais must be positive (it divides the costs of the AIS weights, POL:660,730,803)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(MPOPIS_ERR_NO_DEVICE, "no CUDA device visible: the engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MPOPIS_ERR_BAD_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(MPOPIS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                prop.major, prop.minor);
  const int world = cfg->world_size < 1 ? 1 : cfg->world_size;
  if (cfg->rank < 0 || cfg->rank >= world) return fail(MPOPIS_ERR_BAD_ARG, "rank out of range");
  if (world > 64) return fail(MPOPIS_ERR_BAD_ARG, "world_size must be <= 64");
  if (cfg->num_samples % world) return fail(MPOPIS_ERR_BAD_ARG, "num_samples must be divisible by world_size");

  mpopis_t *h = new mpopis_handle();
  h->cfg = *cfg;
  h->dev = cfg->device, h->world = world, h->rank = cfg->rank;
  h->comm.world = world, h->comm.rank = cfg->rank;
  h->K = (int)cfg->num_samples, h->T = (int)cfg->horizon;
  h->N = (cfg->policy == MPOPIS_POLICY_MPPI || cfg->policy == MPOPIS_POLICY_GMPPI) ? 1 : (int)cfg->opt_its;
  if (cfg->env == MPOPIS_ENV_CAR_RACING) {
    if (cfg->n_cars < 1 || cfg->n_cars > MPOPIS_MAX_CARS) {
      delete h;
      return fail(MPOPIS_ERR_BAD_ARG, "n_cars must be in 1..%d", MPOPIS_MAX_CARS);
    }
    h->as = 2 * cfg->n_cars, h->ss = 8 * cfg->n_cars;
  } else if (cfg->env == MPOPIS_ENV_MOUNTAIN_CAR) {
    h->as = 1, h->ss = 2;
  } else if (cfg->env == MPOPIS_ENV_EXTERNAL) {
    if (cfg->ext_action_size < 1 || cfg->ext_action_size > 4096 || world > 1 || cfg->log_trajectories) {
      delete h;
      return fail(MPOPIS_ERR_BAD_ARG, "external env: ext_action_size must be in 1..4096, world_size 1 and "
                                      "log_trajectories 0 (the simulator's states stay with the caller)");
    }
    h->as = cfg->ext_action_size, h->ss = 0;
  } else {
    delete h;
    return fail(MPOPIS_ERR_BAD_ARG, "unknown env %d", cfg->env);
  }
  h->cs = h->as * h->T;  // POL:59
  h->Kloc = h->K / world, h->k0 = (long long)h->rank * h->Kloc;
  h->ldk = ((long long)h->Kloc + 31) / 32 * 32;
  h->gamma = cfg->lambda * (1 - cfg->alpha);  // POL:266
  h->sigma_bs = 1;
  if (cfg->policy == MPOPIS_POLICY_CEMPPI) {
    h->m_elite = (int)llrint((double)h->K * (1 - cfg->ce_elite_threshold));  // POL:437, ties-to-even
    if (h->N > 1 && (h->m_elite < 2 || h->m_elite > h->K)) {
      delete h;
      return fail(MPOPIS_ERR_BAD_ARG, "m_elite = %d out of range", h->m_elite);
    }
  }
  if (const char *e = getenv("MPOPIS_TRACE")) h->trace = atoi(e) != 0;
  if (const char *e = getenv("MPOPIS_ROLLOUT_VARIANT")) h->rollout_variant = atoi(e);
  if (const char *e = getenv("MPOPIS_ROLLOUT_BLOCK")) h->rollout_block = atoi(e);
  if (const char *e = getenv("MPOPIS_ROLLOUT_STAGE")) h->rollout_stage = atoi(e) != 0;
  if (h->rollout_block < 32 || h->rollout_block > 128 || h->rollout_block % 32) h->rollout_block = 64;

  auto bail = [&](int rc) {
    mpopis_b200_destroy(h);
    return rc;
  };
#define TRY(x)                    \
  do {                            \
    if (int rc_ = (x)) return bail(rc_); \
  } while (0)
  TRY(set_device(h));
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_z_free, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_z_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_q_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_q_join, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(MPOPIS_ERR_CUDA, "cudaStreamCreate failed"));
  const size_t cs = h->cs, K = h->K, Kloc = h->Kloc, ld = h->ldk;
  TRY(dalloc(&h->d_state, h->ss));
  TRY(dalloc(&h->d_env_t, 1));
  TRY(dalloc(&h->d_step, 1));
  if (const char *e = getenv("MPOPIS_GRAPH")) h->graph_enabled = atoi(e) != 0;
  TRY(dalloc(&h->d_U_orig, cs));
  TRY(dalloc(&h->d_U_cur, cs));
  TRY(dalloc(&h->d_U_next, cs));
  TRY(dalloc(&h->d_control, h->as));
  TRY(dalloc(&h->d_Sigma0, cs * cs));
  TRY(dalloc(&h->d_Sigma, cs * cs));
  TRY(dalloc(&h->d_Lt, cs * cs));
  TRY(dalloc(&h->d_Lt0, cs * cs));
  TRY(dalloc(&h->d_cholW, cs * (cs + 1)));
  TRY(dalloc(&h->d_bvec, cs));
  TRY(dalloc(&h->d_Z, cs * ld));
  TRY(dalloc(&h->d_E, cs * ld));
  TRY(dalloc(&h->d_stage, cs * Kloc));
  TRY(dalloc(&h->d_costs, K));
  TRY(dalloc(&h->d_w, K));
  TRY(dalloc(&h->d_sorted, K));
  TRY(dalloc(&h->d_keys_a, K));
  TRY(dalloc(&h->d_keys_b, K));
  TRY(dalloc(&h->d_order, K));
  TRY(dalloc(&h->d_vals_b, K));
  TRY(dalloc(&h->d_hist, 1));
  TRY(dalloc(&h->d_flags, 4));
  TRY(dalloc(&h->d_sums, 2 * cs + 1 + 64));
  TRY(dalloc(&h->d_bvec2, cs));  // [Σx (cs) | count | per-rank early-stop statistics (<= 64 ranks)]
  TRY(dalloc(&h->d_mu, cs));
  TRY(dalloc(&h->d_Sraw, cs * cs + 1));
  TRY(dalloc(&h->d_P, (size_t)98 * cs * cs));  // <= 96 SYRK chunks (stats.cu: syrk_chunk)
  TRY(dalloc(&h->d_q, 1));
  TRY(dalloc(&h->d_lambda, 1));
  TRY(dalloc(&h->d_ones, 512));  // scratch of the multi-CTA weights kernels
  TRY(dalloc(&h->d_reward, 2));  // [reward of the last env step | running sum of the resident loop]
  TRY(dalloc(&h->d_done, 1));
  const size_t nmax = K;  // moments may run over Kloc samples or up to K elites
  h->part_doubles = (size_t)(rowsum_nchunks((int)nmax) + 1) * (2 * cs + 1);
  TRY(dalloc(&h->d_part, h->part_doubles));
  TRY(dalloc(&h->d_qpart, (size_t)shrink_q_nblocks((int)nmax) + 1));
  if (cfg->policy == MPOPIS_POLICY_CEMPPI) TRY(ensure_elite_capacity(h, h->m_elite));
  TRY(dalloc(&h->d_mloc, 1));
  // :cemppi selects its elites without a sort whenever the policy is sharded or K exceeds the single-CTA sort
  // (K <= 2048 on one GPU keeps small_sort_kernel + the single-CTA moment chain: one launch each)
  h->use_select = cfg->policy == MPOPIS_POLICY_CEMPPI && h->N > 1 && (world > 1 || h->K > 2048);
  if (const char *e = getenv("MPOPIS_CE_SELECT")) h->use_select = h->use_select && atoi(e) != 0;  // 0: round-1 sort path (1 GPU)
  if (world > 1 && cfg->policy == MPOPIS_POLICY_CEMPPI && h->N > 1) h->use_select = true;
  if (cfg->policy == MPOPIS_POLICY_CEMPPI && h->N > 1) {
    const size_t mmax = (size_t)std::min(h->m_elite, h->Kloc);
    h->nb_cap = select_bucket_capacity(h->m_elite);
    if (cudaMalloc(&h->d_sel_ws, SELECT_WS_BYTES) != cudaSuccess) return bail(fail(MPOPIS_ERR_CUDA, "cudaMalloc failed"));
    TRY(dalloc(&h->d_bmin, (size_t)h->nb_cap));
    TRY(dalloc(&h->d_bmax, (size_t)h->nb_cap));
    TRY(dalloc(&h->d_eidx, mmax));
    TRY(dalloc(&h->d_tau, 4));
    launch_select_init(h->d_sel_ws, h->d_bmin, h->d_bmax, h->nb_cap, h->st);
    const size_t need = (size_t)elite_gather_nchunks((int)mmax) * 2 * cs + 2 * cs + 2;
    if (need > h->part_doubles) {
      cudaFree(h->d_part);
      h->d_part = nullptr, h->part_doubles = need;
      TRY(dalloc(&h->d_part, need));
    }
  }
  if (cfg->policy == MPOPIS_POLICY_PMCMPPI || cfg->policy == MPOPIS_POLICY_CMAMPPI) {
    TRY(dalloc(&h->d_u, K));
    TRY(dalloc(&h->d_cdf, K));
    TRY(dalloc(&h->d_counts, K));
    TRY(dalloc(&h->d_wcnt, Kloc));
  }
  if (cfg->policy == MPOPIS_POLICY_CMAMPPI) {
    TRY(dalloc(&h->d_ws, K));
    TRY(dalloc(&h->d_sigma, 1));
    TRY(dalloc(&h->d_psig, cs));
    TRY(dalloc(&h->d_pSig, cs));
    TRY(dalloc(&h->d_C, cs * cs));
    TRY(dalloc(&h->d_ns, 5 * cs * cs + ((cs + 31) / 32) * ((cs + 31) / 32) + 8));
  }
  if (cfg->log_trajectories) TRY(dalloc(&h->d_traj, Kloc * (size_t)h->T * h->ss));
  if (cfg->env == MPOPIS_ENV_EXTERNAL) {
    TRY(dalloc(&h->d_ext_cc, Kloc));
    TRY(dalloc(&h->d_ext_in, Kloc));
    TRY(dalloc(&h->d_ext_bounds, 2 * (size_t)h->as));
    if (cudaMallocHost((void **)&h->h_ext_controls, sizeof(double) * cs * Kloc) != cudaSuccess ||
        cudaMallocHost((void **)&h->h_ext_costs, sizeof(double) * Kloc) != cudaSuccess ||
        cudaMallocHost((void **)&h->h_ext_stop, sizeof(int)) != cudaSuccess)
      return bail(fail(MPOPIS_ERR_CUDA, "cudaMallocHost failed"));
  }
  if (cudaMallocHost((void **)&h->h_in, sizeof(double) * (h->ss + cs + 2)) != cudaSuccess ||
      cudaMallocHost((void **)&h->h_out, sizeof(double) * (h->as + cs + h->ss + 2)) != cudaSuccess ||
      cudaMallocHost((void **)&h->h_flags, sizeof(int) * 4) != cudaSuccess)
    return bail(fail(MPOPIS_ERR_CUDA, "cudaMallocHost failed"));
  h->ev.resize(2 + 2 * h->N);
  for (auto &e : h->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(MPOPIS_ERR_CUDA, "cudaEventCreate failed"));
  h->num_sms = prop.multiProcessorCount;
  h->coop_max = inv_sqrt_max_ctas(prop.multiProcessorCount);
  h->sort_max = sort_max_ctas(prop.multiProcessorCount);
  h->sel_max = select_max_ctas(prop.multiProcessorCount);
  {  // Σ defaults to the identity until set_sigma()
    std::vector<double> I(cs * cs, 0.0);
    for (size_t i = 0; i < cs; ++i) I[i * cs + i] = 1.0;
    if (cudaMemcpy(h->d_Sigma0, I.data(), sizeof(double) * cs * cs, cudaMemcpyHostToDevice) != cudaSuccess)
      return bail(fail(MPOPIS_ERR_CUDA, "cudaMemcpy failed"));
  }
#undef TRY
  *out = h;
  return 0;
}

int mpopis_b200_destroy(mpopis_t *h) {
  if (!h) return 0;
  cudaSetDevice(h->dev);
  if (h->st) cudaStreamSynchronize(h->st);
  drop_graph(h);
  comm_destroy(h->comm);
  if (h->d_step) cudaFree(h->d_step);
  void *ptrs[] = {h->d_trk,    h->d_state, h->d_U_orig, h->d_U_cur,  h->d_U_next, h->d_control, h->d_Sigma0,
                  h->d_Sigma,  h->d_Lt,    h->d_Lt0,    h->d_cholW,  h->d_bvec,   h->d_Z,       h->d_E,
                  h->d_stage,  h->d_costs, h->d_w,      h->d_sorted, h->d_X,      h->d_mask,    h->d_part,
                  h->d_sums,   h->d_mu,    h->d_P,      h->d_Sraw,   h->d_qpart,  h->d_q,       h->d_lambda,
                  h->d_traj,   h->d_u,     h->d_cdf,    h->d_wcnt,   h->d_ws,     h->d_sigma,   h->d_psig,
                  h->d_pSig,   h->d_C,     h->d_ns,     h->d_reward, h->d_env_t,  h->d_keys_a,  h->d_keys_b,
                  h->d_order,  h->d_vals_b, h->d_hist,  h->d_counts, h->d_flags,  h->d_done,   h->d_lut,
                  h->d_ones,   h->d_mloc,   h->d_bvec2,  h->d_sel_ws, h->d_bmin,  h->d_bmax,    h->d_eidx,
                  h->d_tau};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (h->d_warp_cycles) cudaFree(h->d_warp_cycles);
  if (h->h_ext_controls) cudaFreeHost(h->h_ext_controls);
  if (h->h_ext_costs) cudaFreeHost(h->h_ext_costs);
  if (h->h_ext_stop) cudaFreeHost(h->h_ext_stop);
  if (h->d_ext_cc) cudaFree(h->d_ext_cc);
  if (h->d_ext_in) cudaFree(h->d_ext_in);
  if (h->d_ext_bounds) cudaFree(h->d_ext_bounds);
  if (h->h_in) cudaFreeHost(h->h_in);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  for (auto &e : h->ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : h->mark_pool) cudaEventDestroy(e);
  for (int b = 0; b < 2; ++b) {
    if (h->h_pin[b]) cudaFreeHost(h->h_pin[b]);
    if (h->ev_pin[b]) cudaEventDestroy(h->ev_pin[b]);
  }
  if (h->ev_z_free) cudaEventDestroy(h->ev_z_free);
  if (h->ev_z_ready) cudaEventDestroy(h->ev_z_ready);
  if (h->ev_q_fork) cudaEventDestroy(h->ev_q_fork);
  if (h->ev_q_join) cudaEventDestroy(h->ev_q_join);
  if (h->st2) cudaStreamSynchronize(h->st2), cudaStreamDestroy(h->st2);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

int mpopis_b200_comm_id(void *out128) {
  if (!out128) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (comm_unique_id(out128)) return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

int mpopis_b200_comm_init(mpopis_t *h, const void *id128) {
  if (!h || !id128) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->world == 1) return 0;
  if (h->comm.nccl || h->comm.loop) return fail(MPOPIS_ERR_BAD_ARG, "communicator already initialised");
  if (int rc = set_device(h)) return rc;
  if (comm_init_nccl(h->comm, id128)) return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

// Loop-back group: `world` virtual ranks on ONE device in ONE process, each handle driven from its own host thread
// (comm.cu). Verification of the sharded path on a single-GPU box; the production communicator is NCCL.
int mpopis_b200_loopback_create(int32_t world, void **group_out) {
  if (!group_out || world < 1 || world > COMM_MAX_WORLD) return fail(MPOPIS_ERR_BAD_ARG, "world must be 1..%d", COMM_MAX_WORLD);
  *group_out = loop_group_create(world);
  return *group_out ? 0 : fail(MPOPIS_ERR_BAD_ARG, "cannot create the loop-back group");
}

int mpopis_b200_loopback_destroy(void *group) {
  loop_group_destroy((LoopGroup *)group);
  return 0;
}

int mpopis_b200_comm_init_loopback(mpopis_t *h, void *group) {
  if (!h || !group) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->world == 1) return 0;
  if (h->comm.nccl || h->comm.loop) return fail(MPOPIS_ERR_BAD_ARG, "communicator already initialised");
  if (int rc = set_device(h)) return rc;
  const size_t cap = std::max((size_t)h->cs * h->cs + 64, (size_t)h->K) + 2 * (size_t)h->cs + 128;
  if (comm_init_loopback(h->comm, (LoopGroup *)group, h->dev, cap)) return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

// Peer-memory collectives (comm.cu): export -> the host exchanges the 128-byte blobs of all ranks -> attach.
int mpopis_b200_comm_peer_export(mpopis_t *h, void *out128) {
  if (!h || !out128) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->world == 1) return fail(MPOPIS_ERR_BAD_ARG, "not a sharded handle");
  if (int rc = set_device(h)) return rc;
  if (comm_peer_alloc(h->comm, (size_t)h->cs * h->cs + 2 * (size_t)h->cs + 64) ||
      comm_peer_export(h->comm, h->d_costs, out128))
    return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

int mpopis_b200_comm_peer_attach(mpopis_t *h, const void *handles, int64_t n_bytes) {
  if (!h || !handles) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (n_bytes != (int64_t)h->world * COMM_PEER_HANDLE)
    return fail(MPOPIS_ERR_BAD_ARG, "expected world_size x %d bytes of peer handles", COMM_PEER_HANDLE);
  if (int rc = set_device(h)) return rc;
  drop_graph(h);
  h->comm.peer_err = h->info();
  if (comm_peer_attach(h->comm, h->d_costs, handles)) return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

int mpopis_b200_comm_peer_loopback(mpopis_t *h) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->world == 1) return 0;
  // :cmamppi keeps the host-barrier transport here: its K-sized all-reduce is beyond the peer slots and would mix a
  // host-synchronous collective into a sequence of device-side spins. (Grid-cooperative kernels — the selection — do
  // run beside a spinning kernel of another stream: tools/coop_concurrency.cu.)
  if (h->cfg.policy == MPOPIS_POLICY_CMAMPPI)
    return fail(MPOPIS_ERR_BAD_ARG, "peer-memory collectives between loop-back ranks do not support :cmamppi");
  if (int rc = set_device(h)) return rc;
  drop_graph(h);
  h->comm.peer_err = h->info();
  // allocations may synchronise the device: make the ones the injected-noise path needs now, not under a spinning peer
  if (int rc = ensure_pin(h, sizeof(double) * std::max((size_t)h->cs * h->Kloc, (size_t)h->K))) return rc;
  if (comm_peer_alloc(h->comm, (size_t)h->cs * h->cs + 2 * (size_t)h->cs + 64) ||
      comm_peer_attach_loopback(h->comm, h->d_costs))
    return fail(MPOPIS_ERR_NCCL, "%s", comm_error());
  return 0;
}

int mpopis_b200_set_car_env(mpopis_t *h, int32_t n_cars, const double *params, double dt, double ddt,
                            const double *trk_x, const double *trk_y, const double *trk_w, int64_t n_trk) {
  if (!h || !params || !trk_x || !trk_y || !trk_w) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_CAR_RACING || n_cars != h->cfg.n_cars)
    return fail(MPOPIS_ERR_BAD_ARG, "handle was created for a different environment");
  if (n_trk < 2 || n_trk > 9000) return fail(MPOPIS_ERR_BAD_ARG, "track must have 2..9000 sampled points");
  if (!(dt > 0) || !(ddt > 0)) return fail(MPOPIS_ERR_BAD_ARG, "dt and δt must be positive");
  if (int rc = set_device(h)) return rc;
  static_assert(sizeof(CarParams) == sizeof(double) * MPOPIS_CAR_NPARAMS, "CarParams layout");
  for (int c = 0; c < n_cars; ++c) {
    memcpy(&h->car.car[c], params + (size_t)c * MPOPIS_CAR_NPARAMS, sizeof(CarParams));
    h->car.cos_blimit[c] = h->car.car[c].b_limit >= M_PI ? -2.0 : std::cos(h->car.car[c].b_limit);
  }
  h->car.dt = dt, h->car.ddt = ddt, h->car.nsub = (int)lrint(dt / ddt);  // CAR:299
  for (int c = 0; c < n_cars; ++c) h->car.der[c] = derive_car(h->car.car[c], ddt);
  drop_graph(h);  // the env arguments are kernel parameters of the captured step
  h->car.n_cars = n_cars, h->car.n_trk = (int)n_trk;
  if (h->d_trk) cudaFree(h->d_trk), h->d_trk = nullptr;
  if (int rc = dalloc(&h->d_trk, 3 * (size_t)n_trk)) return rc;
  CU(cudaMemcpy(h->d_trk, trk_x, sizeof(double) * n_trk, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_trk + n_trk, trk_y, sizeof(double) * n_trk, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_trk + 2 * n_trk, trk_w, sizeof(double) * n_trk, cudaMemcpyHostToDevice));
  h->car.trk = h->d_trk;
  {  // exact nearest-point pruning table for the fast rollout variant
    std::vector<uint16_t> cells;
    double x0, y0, cell;
    int nx, ny;
    build_track_lut(trk_x, trk_y, trk_w, (int)n_trk, cells, x0, y0, cell, nx, ny);
    if (h->d_lut) cudaFree(h->d_lut), h->d_lut = nullptr;
    CU(cudaMalloc((void **)&h->d_lut, cells.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(h->d_lut, cells.data(), cells.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    h->car.lut = h->d_lut, h->car.lut_x0 = x0, h->car.lut_y0 = y0, h->car.lut_inv_c = 1.0 / cell;
    h->car.lut_nx = nx, h->car.lut_ny = ny;
  }
  h->env_set = true;
  return 0;
}

int mpopis_b200_set_mountaincar_env(mpopis_t *h, const double *p, int64_t max_steps) {
  if (!h || !p) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_MOUNTAIN_CAR)
    return fail(MPOPIS_ERR_BAD_ARG, "handle was created for a different environment");
  drop_graph(h);
  h->mc = McEnvArgs{p[0], p[1], p[2], p[3], p[4], p[5], p[6], (long long)max_steps};
  h->env_set = true;
  return 0;
}

// cov_mat handling of MPPI_Policy_Params (POL:66-81) + block_diagm (UTL:9-21)
int mpopis_b200_set_sigma(mpopis_t *h, const double *Sigma, int64_t n) {
  if (!h || !Sigma) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  const int cs = h->cs, as = h->as;
  const int check = is_g_family(h->cfg.policy) ? cs : as;  // POL:66-74
  std::vector<double> S((size_t)cs * cs, 0.0);
  if (n == as) {
    for (int b = 0; b < cs; b += as)
      for (int i = 0; i < as; ++i)
        for (int j = 0; j < as; ++j) S[(size_t)(b + j) * cs + b + i] = Sigma[(size_t)j * as + i];
    h->sigma_bs = as;
  } else if (n == cs && check == cs) {
    memcpy(S.data(), Sigma, sizeof(double) * cs * cs);
    h->sigma_bs = cs;
  } else {
    return fail(MPOPIS_ERR_BAD_ARG, "Covariance matrix size problem");
  }
  for (int i = 0; i < cs; ++i)
    for (int j = 0; j < i; ++j)
      if (S[(size_t)j * cs + i] != S[(size_t)i * cs + j])
        return fail(MPOPIS_ERR_NOT_PD, "PosDefException: covariance matrix is not symmetric");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemcpy(h->d_Sigma0, S.data(), sizeof(double) * cs * cs, cudaMemcpyHostToDevice));
  h->L0_valid = false;
  drop_graph(h);  // sigma_bs selects the E = L·Z kernel of the first iteration
  return 0;
}

int mpopis_b200_set_cma(mpopis_t *h, const mpopis_cma_t *cma, const double *ws, int64_t n_ws) {
  if (!h || !cma || !ws) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.policy != MPOPIS_POLICY_CMAMPPI) return fail(MPOPIS_ERR_BAD_ARG, "not a :cmamppi handle");
  if (n_ws != h->K) return fail(MPOPIS_ERR_BAD_ARG, "ws must have num_samples entries");
  if (h->N > 1 && (cma->m_elite < 2 || cma->m_elite > h->K)) return fail(MPOPIS_ERR_BAD_ARG, "m_elite out of range");
  if (int rc = set_device(h)) return rc;
  drop_graph(h);
  h->cma = *cma;
  h->m_elite = (int)cma->m_elite;
  if (int rc = ensure_elite_capacity(h, h->m_elite)) return rc;
  CU(cudaMemcpy(h->d_ws, ws, sizeof(double) * n_ws, cudaMemcpyHostToDevice));
  h->cma_set = true;
  return 0;
}

int mpopis_b200_seed(mpopis_t *h, uint64_t seed) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  h->seed = seed;  // a kernel argument of the captured step: re-capture
  h->step = 0;
  drop_graph(h);
  CU(cudaMemsetAsync(h->d_step, 0, sizeof(unsigned), h->st));
  return 0;
}

int mpopis_b200_set_option(mpopis_t *h, const char *key, double value) {
  if (!h || !key) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  drop_graph(h);  // every option changes the captured launch sequence
  if (!strcmp(key, "graph")) {  // 0: launch every kernel of a control step individually (A/B, debugging)
    h->graph_enabled = value != 0.0;
    return 0;
  }
  if (!strcmp(key, "rollout_variant")) {
    if (value != 0.0 && value != 1.0 && value != 3.0 && value != 4.0 && value != 5.0 && value != 6.0)
      return fail(MPOPIS_ERR_BAD_ARG, "rollout_variant must be 0 (v3), 1 (literal), 3 (v4), 4 / 5 (v5, warp-specialised) or 6 (auto)");
    h->rollout_variant = (int)value;
  }
  else if (!strcmp(key, "rollout_profile")) {  // per-warp clock64() of the rollout kernel, read with warp_cycles()
    if (value != 0.0 && !h->d_warp_cycles) {
      if (int rc = dalloc(&h->d_warp_cycles, 8 * ((size_t)h->Kloc / 32 + 2))) return rc;  // split kernel: 3 warps x 2 per 64 + phases
    } else if (value == 0.0 && h->d_warp_cycles) {
      cudaFree(h->d_warp_cycles), h->d_warp_cycles = nullptr;
    }
  }
  else if (!strcmp(key, "rollout_stage")) {
    if (value != 0.0 && value != 1.0) return fail(MPOPIS_ERR_BAD_ARG, "rollout_stage must be 0 or 1");
    h->rollout_stage = (int)value;
  }
  else if (!strcmp(key, "rollout_block")) {
    const int b = (int)value;
    if (b < 32 || b > 128 || b % 32) return fail(MPOPIS_ERR_BAD_ARG, "rollout_block must be 32, 64, 96 or 128");
    h->rollout_block = b;
  } else if (!strcmp(key, "moments_small")) {
    h->moments_small = value != 0.0;
  } else if (!strcmp(key, "rollout_spin")) {
    h->rollout_spin = value < 0 ? -1 : (value != 0.0);
  } else if (!strcmp(key, "ce_small_fused")) {
    h->small_fused = value != 0.0;
  } else if (!strcmp(key, "fuse_cov")) {
    h->fuse_cov = value != 0.0;
  } else
    return fail(MPOPIS_ERR_BAD_ARG, "unknown option %s", key);
  return 0;
}

int mpopis_b200_get_option(mpopis_t *h, const char *key, double *value_out) {
  if (!h || !key || !value_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!strcmp(key, "rollout_variant")) *value_out = h->rollout_variant;
  else if (!strcmp(key, "rollout_variant_used")) *value_out = h->rollout_variant_used;  // of the latest launch (6 resolved)
  else if (!strcmp(key, "comm_peer")) *value_out = h->comm.peer;
  else if (!strcmp(key, "ce_small_fused")) *value_out = h->small_fused;
  else if (!strcmp(key, "rollout_block")) *value_out = h->rollout_block;
  else if (!strcmp(key, "rollout_stage")) *value_out = h->rollout_stage;
  else if (!strcmp(key, "moments_small")) *value_out = h->moments_small;
  else if (!strcmp(key, "graph")) *value_out = h->graph_enabled;
  else if (!strcmp(key, "graph_active")) *value_out = h->gexec != nullptr;
  else if (!strcmp(key, "ce_select")) *value_out = h->use_select;
  else return fail(MPOPIS_ERR_BAD_ARG, "unknown option %s", key);
  return 0;
}

int mpopis_b200_plan_with_noise(mpopis_t *h, const double *state, int64_t env_t, double *U_inout,
                                const double *Z, const double *resample_u, double *control_out,
                                int32_t *its_run_out) {
  if (!h || !state || !U_inout || !control_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemsetAsync(h->info(), 0, sizeof(int), h->st));
  if (int rc = upload_inputs(h, state, env_t, U_inout)) return rc;
  if (Z || resample_u) factor_sigma0(h);
  if (int rc = (Z || resample_u) ? plan_core(h, Z, resample_u) : plan_step(h)) {
    cudaStreamSynchronize(h->st);
    return rc;
  }
  return download_outputs(h, U_inout, control_out, its_run_out, nullptr);
}

int mpopis_b200_plan(mpopis_t *h, const double *state, int64_t env_t, double *U_inout, double *control_out,
                     int32_t *its_run_out) {
  return mpopis_b200_plan_with_noise(h, state, env_t, U_inout, nullptr, nullptr, control_out, its_run_out);
}

int mpopis_b200_set_external_env(mpopis_t *h, const double *action_lo, const double *action_hi) {
  if (!h || !action_lo || !action_hi) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_EXTERNAL) return fail(MPOPIS_ERR_BAD_ARG, "handle was created for a different environment");
  for (int a = 0; a < h->as; ++a)
    if (!(action_lo[a] <= action_hi[a])) return fail(MPOPIS_ERR_BAD_ARG, "action bounds: lo must be <= hi");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemcpy(h->d_ext_bounds, action_lo, sizeof(double) * h->as, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_ext_bounds + h->as, action_hi, sizeof(double) * h->as, cudaMemcpyHostToDevice));
  h->env_set = true;
  return 0;
}

int mpopis_b200_plan_external(mpopis_t *h, double *U_inout, mpopis_rollout_fn rollout, void *user, const double *Z,
                              const double *resample_u, double *control_out, int32_t *its_run_out) {
  if (!h || !U_inout || !rollout || !control_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (h->cfg.env != MPOPIS_ENV_EXTERNAL) return fail(MPOPIS_ERR_BAD_ARG, "handle was not created with MPOPIS_ENV_EXTERNAL");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemsetAsync(h->info(), 0, sizeof(int), h->st));
  if (int rc = upload_inputs(h, nullptr, 0, U_inout)) return rc;
  h->ext_fn = rollout, h->ext_user = user;
  factor_sigma0(h);
  const int rc = plan_core(h, Z, resample_u);
  h->ext_fn = nullptr, h->ext_user = nullptr;
  if (rc) {
    cudaStreamSynchronize(h->st);
    return rc;
  }
  return download_outputs(h, U_inout, control_out, its_run_out, nullptr);
}

int mpopis_b200_fetch(mpopis_t *h, double *costs, double *weights, double *E, double *traj) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  if (costs) CU(cudaMemcpyAsync(costs, h->d_costs, sizeof(double) * h->K, cudaMemcpyDeviceToHost, h->st));
  if (weights) CU(cudaMemcpyAsync(weights, h->d_w, sizeof(double) * h->K, cudaMemcpyDeviceToHost, h->st));
  if (E) {  // E .+ (pol.U − U_orig), POL:370,468,602,668,739,814, back in Julia's cs x K layout
    launch_transpose_out(h->d_E, h->d_stage, h->cs, h->Kloc, h->ldk, h->d_U_cur, h->d_U_orig, h->st);
    h->launches += 1;
    CU(cudaMemcpyAsync(E, h->d_stage, sizeof(double) * h->cs * h->Kloc, cudaMemcpyDeviceToHost, h->st));
  }
  if (traj) {
    if (!h->d_traj) return fail(MPOPIS_ERR_BAD_ARG, "log_trajectories was not enabled");
    // device layout [k][ss][T] is exactly K column-major T x ss matrices
    CU(cudaMemcpyAsync(traj, h->d_traj, sizeof(double) * (size_t)h->Kloc * h->T * h->ss, cudaMemcpyDeviceToHost, h->st));
  }
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

int mpopis_b200_fetch_proposal(mpopis_t *h, double *Sigma_last, double *U_last) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  const size_t cs = h->cs;
  if (Sigma_last) {  // Σ′ the last executed iteration sampled from = L Lᵀ is not stored; return the matrix factored
    const bool adapted = adapts_sigma(h->cfg.policy) && h->N > 1;
    CU(cudaMemcpyAsync(Sigma_last, adapted ? h->d_Sigma : h->d_Sigma0, sizeof(double) * cs * cs,
                       cudaMemcpyDeviceToHost, h->st));
  }
  if (U_last) CU(cudaMemcpyAsync(U_last, h->d_U_cur, sizeof(double) * cs, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

int mpopis_b200_rollout_costs(mpopis_t *h, const double *state, int64_t env_t, const double *U,
                              const double *U_orig, const double *E, const double *Sigma_inv,
                              double *costs_out) {
  if (!h || !state || !U || !U_orig || !E || !costs_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!h->env_set) return fail(MPOPIS_ERR_BAD_ARG, "environment not set");
  if (h->gamma != 0.0 && !Sigma_inv) return fail(MPOPIS_ERR_BAD_ARG, "Sigma_inv required when γ = λ(1-α) != 0");
  if (int rc = set_device(h)) return rc;
  const int cs = h->cs;
  cudaStream_t st = h->st;
  if (int rc = upload_inputs(h, state, env_t, U_orig)) return rc;
  CU(cudaMemcpyAsync(h->d_U_cur, U, sizeof(double) * cs, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 3, st));
  CU(cudaMemcpyAsync(h->d_stage, E + (size_t)h->k0 * cs, sizeof(double) * cs * h->Kloc, cudaMemcpyHostToDevice, st));
  launch_transpose_in(h->d_stage, h->d_E, cs, h->Kloc, h->ldk, st);
  const double *bvec = nullptr;
  if (h->gamma != 0.0) {
    CU(cudaMemcpyAsync(h->d_cholW, Sigma_inv, sizeof(double) * cs * cs, cudaMemcpyHostToDevice, st));
    launch_ctrl_vec(h->d_cholW, cs, h->d_U_orig, h->gamma, h->d_bvec, st);
    bvec = h->d_bvec;
    h->launches += 1;
  }
  h->launches += 1;
  if (int rc = launch_rollouts(h, h->d_U_cur, h->d_U_orig, bvec)) return rc;
  if (int rc = allgather_costs(h)) return rc;
  CU(cudaMemcpyAsync(costs_out, h->d_costs, sizeof(double) * h->K, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  return 0;
}

int mpopis_b200_weights(mpopis_t *h, const double *costs, int64_t K, double lambda, double *w_out) {
  if (!h || !costs || !w_out || K < 1) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dc, dw;  // freed on every return path
  if (int rc = dc.alloc((size_t)K)) return rc;
  if (int rc = dw.alloc((size_t)K)) return rc;
  CU(cudaMemcpyAsync(dc, costs, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  h->launches += launch_weights(dc, (int)K, lambda, dw, h->d_ones, nullptr, h->st);
  CU(cudaMemcpyAsync(w_out, dw, sizeof(double) * K, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

int mpopis_b200_sortperm(mpopis_t *h, const double *costs, int64_t K, int64_t *perm_out) {
  if (!h || !costs || !perm_out || K < 1 || K > (1LL << 30)) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dc;
  DevBuf<unsigned long long> ka, kb;
  DevBuf<int> ord, vb;
  if (int rc = dc.alloc((size_t)K)) return rc;
  if (int rc = ka.alloc((size_t)K)) return rc;
  if (int rc = kb.alloc((size_t)K)) return rc;
  if (int rc = ord.alloc((size_t)K)) return rc;
  if (int rc = vb.alloc((size_t)K)) return rc;
  CU(cudaMemcpyAsync(dc, costs, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  {
    const cudaError_t e = (cudaError_t)launch_sortperm(dc, (int)K, ka, kb, ord, vb, 0, 0, nullptr, nullptr, h->sort_max,
                                                       h->st);
    if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  }
  h->launches += sort_launches((int)K);
  std::vector<int> tmp((size_t)K);
  CU(cudaMemcpyAsync(tmp.data(), ord, sizeof(int) * K, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  for (int64_t i = 0; i < K; ++i) perm_out[i] = tmp[(size_t)i];
  return 0;
}

// Parity surface of select.cu: the elite set order[1:m] of sortperm(costs) (returned as ascending sample ids) and the
// early-stop decision of POL:458-461, computed WITHOUT a sort. [k0, k0 + kloc) emulates the shard ownership window.
int mpopis_b200_elite_select(mpopis_t *h, const double *costs, int64_t K, int64_t m, int64_t k0, int64_t kloc,
                             int32_t early_stop, int64_t *elite_ids_out, int64_t *n_out, int32_t *stop_out,
                             double *tau_out4) {
  if (!h || !costs || !elite_ids_out || !n_out || !stop_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (K < 1 || K > (1LL << 30) || m < 1 || m > K || k0 < 0 || kloc < 1 || k0 + kloc > K)
    return fail(MPOPIS_ERR_BAD_ARG, "bad sizes");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dc, dtau;
  DevBuf<unsigned long long> bmin, bmax;
  DevBuf<int> eidx, misc;
  DevBuf<unsigned char> ws;
  const long long cap = select_bucket_capacity((int)m);
  const size_t mmax = (size_t)std::min(m, kloc);
  if (int rc = dc.alloc((size_t)K)) return rc;
  if (int rc = dtau.alloc(4)) return rc;
  if (int rc = bmin.alloc((size_t)cap)) return rc;
  if (int rc = bmax.alloc((size_t)cap)) return rc;
  if (int rc = eidx.alloc(mmax)) return rc;
  if (int rc = misc.alloc(2)) return rc;  // [m_loc, stop]
  if (int rc = ws.alloc(SELECT_WS_BYTES)) return rc;
  CU(cudaMemcpyAsync(dc, costs, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  launch_select_init(ws.p, bmin, bmax, cap, h->st);
  for (int rep = 0; rep < 2; ++rep) {  // twice: the second run starts from the workspace the first one left behind
    CU(cudaMemsetAsync(misc, 0, sizeof(int) * 2, h->st));
    const cudaError_t e = (cudaError_t)launch_ce_select(dc, (int)K, (int)m, k0, (int)kloc, early_stop, ws.p, bmin, bmax, cap,
                                                        eidx, misc.p, dtau, misc.p + 1, nullptr, h->sel_max,
                                                        h->st);
    if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  }
  h->launches += 3;
  int hm[2] = {0, 0};
  CU(cudaMemcpyAsync(hm, misc, sizeof hm, cudaMemcpyDeviceToHost, h->st));
  if (tau_out4) CU(cudaMemcpyAsync(tau_out4, dtau, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  if (hm[0] < 0 || (size_t)hm[0] > mmax) return fail(MPOPIS_ERR_CUDA, "selection returned %d local elites", hm[0]);
  std::vector<int> tmp((size_t)hm[0] + 1);
  CU(cudaMemcpy(tmp.data(), eidx, sizeof(int) * (size_t)hm[0], cudaMemcpyDeviceToHost));
  for (int i = 0; i < hm[0]; ++i) elite_ids_out[i] = tmp[(size_t)i] + k0;
  *n_out = hm[0], *stop_out = hm[1];
  return 0;
}

int mpopis_b200_track_query(mpopis_t *h, const double *pos, int64_t n, int32_t *idx_out, int32_t *idx2_out,
                            double *dist_out, uint8_t *within_out) {
  if (!h || !pos || n < 1) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (!h->env_set || h->cfg.env != MPOPIS_ENV_CAR_RACING) return fail(MPOPIS_ERR_BAD_ARG, "car env not set");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dp, dd;
  DevBuf<int> di, dj;
  DevBuf<unsigned char> dw;
  if (int rc = dp.alloc(2 * (size_t)n)) return rc;
  if (int rc = dd.alloc((size_t)n)) return rc;
  if (int rc = di.alloc((size_t)n)) return rc;
  if (int rc = dj.alloc((size_t)n)) return rc;
  if (int rc = dw.alloc((size_t)n)) return rc;
  CU(cudaMemcpyAsync(dp, pos, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, h->st));
  launch_track_query(h->car, dp, (int)n, di, dj, dd, dw, h->rollout_variant != 1, h->st);
  h->launches += 1;
  if (idx_out) CU(cudaMemcpyAsync(idx_out, di, sizeof(int) * n, cudaMemcpyDeviceToHost, h->st));
  if (idx2_out) CU(cudaMemcpyAsync(idx2_out, dj, sizeof(int) * n, cudaMemcpyDeviceToHost, h->st));
  if (dist_out) CU(cudaMemcpyAsync(dist_out, dd, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st));
  if (within_out) CU(cudaMemcpyAsync(within_out, dw, n, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  return 0;
}

static int env_step_device(mpopis_t *h, const double *d_action) {
  if (h->cfg.env == MPOPIS_ENV_CAR_RACING)
    launch_env_step_car(h->car, h->d_state, d_action, h->d_env_t, h->d_reward, h->rollout_variant >= 4 ? 3 : h->rollout_variant, h->st);
  else
    launch_env_step_mc(h->mc, h->d_state, d_action, h->d_env_t, h->d_reward, h->d_done, h->st);
  h->launches += 1;
  return 0;
}

int mpopis_b200_env_step(mpopis_t *h, double *state_inout, const double *action, int64_t *env_t_inout,
                         double *reward_out, uint8_t *done_out) {
  if (!h || !state_inout || !action || !env_t_inout) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!h->env_set) return fail(MPOPIS_ERR_BAD_ARG, "environment not set");
  if (h->cfg.env == MPOPIS_ENV_EXTERNAL) return fail(MPOPIS_ERR_BAD_ARG, "external env: the simulator lives with the caller");
  if (int rc = set_device(h)) return rc;
  long long t = *env_t_inout;
  CU(cudaMemcpyAsync(h->d_state, state_inout, sizeof(double) * h->ss, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->d_env_t, &t, sizeof t, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->d_control, action, sizeof(double) * h->as, cudaMemcpyHostToDevice, h->st));
  env_step_device(h, h->d_control);
  double rew = 0;
  unsigned char done = 0;
  CU(cudaMemcpyAsync(state_inout, h->d_state, sizeof(double) * h->ss, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(&t, h->d_env_t, sizeof t, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(&rew, h->d_reward, sizeof rew, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(&done, h->d_done, 1, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  *env_t_inout = t;
  if (reward_out) *reward_out = rew;
  if (done_out) *done_out = h->cfg.env == MPOPIS_ENV_MOUNTAIN_CAR ? done : 0;
  return 0;
}

int mpopis_b200_env_reward(mpopis_t *h, const double *state, uint8_t done, double *reward_out) {
  if (!h || !state || !reward_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!h->env_set) return fail(MPOPIS_ERR_BAD_ARG, "environment not set");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemcpyAsync(h->d_state, state, sizeof(double) * h->ss, cudaMemcpyHostToDevice, h->st));
  if (h->cfg.env == MPOPIS_ENV_CAR_RACING)
    launch_env_reward_car(h->car, h->d_state, h->d_reward, h->rollout_variant >= 4 ? 3 : h->rollout_variant, h->st);
  else
    launch_env_reward_mc(h->mc, h->d_state, done, h->d_reward, h->st);
  h->launches += 1;
  CU(cudaMemcpyAsync(reward_out, h->d_reward, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  return 0;
}

int mpopis_b200_sample_normals(mpopis_t *h, int64_t step, int64_t iteration, double *Z_out) {
  if (!h || !Z_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  launch_philox_normals(h->d_Z, h->ldk, h->cs, h->Kloc, h->k0, h->seed, (uint32_t)step, nullptr, (uint32_t)iteration, nullptr,
                        h->st);
  launch_transpose_out(h->d_Z, h->d_stage, h->cs, h->Kloc, h->ldk, nullptr, nullptr, h->st);
  h->launches += 2;
  CU(cudaMemcpyAsync(Z_out, h->d_stage, sizeof(double) * h->cs * h->Kloc, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  return 0;
}

int mpopis_b200_cov_estimate(mpopis_t *h, int32_t sigma_est, const double *X, int64_t p, int64_t n,
                             const double *w, int32_t corrected, double *mean_out, double *cov_out) {
  if (!h || !X || n < 2) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (p != h->cs) return fail(MPOPIS_ERR_BAD_ARG, "p must equal the handle's control size cs = %d", h->cs);
  if (n > h->K) return fail(MPOPIS_ERR_BAD_ARG, "n exceeds the handle's capacity (num_samples)");
  if (h->world != 1) return fail(MPOPIS_ERR_BAD_ARG, "cov_estimate is a single-shard parity surface");
  if (int rc = set_device(h)) return rc;
  const long long ld = ((long long)n + 31) / 32 * 32;
  DevBuf<double> dcm, dX, dw, dS;
  if (int rc = dcm.alloc((size_t)p * n)) return rc;
  if (int rc = dX.alloc((size_t)p * ld)) return rc;
  if (int rc = dS.alloc((size_t)p * p)) return rc;
  CU(cudaMemcpyAsync(dcm, X, sizeof(double) * p * n, cudaMemcpyHostToDevice, h->st));
  launch_transpose_in(dcm, dX, (int)p, (int)n, ld, h->st);
  if (w) {
    if (int rc = dw.alloc((size_t)n)) return rc;
    CU(cudaMemcpyAsync(dw, w, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
  }
  CU(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 3, h->st));
  const int method = (w || corrected) ? MPOPIS_SIGMA_MLE : sigma_est;
  if (int rc = moments(h, dX, ld, (int)n, dw, cov_out != nullptr, corrected, method, 0.0, false, nullptr, dS)) return rc;
  if (mean_out) CU(cudaMemcpyAsync(mean_out, h->d_mu, sizeof(double) * p, cudaMemcpyDeviceToHost, h->st));
  if (cov_out) CU(cudaMemcpyAsync(cov_out, dS, sizeof(double) * p * p, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  return 0;
}

int mpopis_b200_warp_cycles(mpopis_t *h, int64_t *cycles_out, int64_t n) {
  if (!h || !cycles_out || n < 1) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (!h->d_warp_cycles) return fail(MPOPIS_ERR_BAD_ARG, "set_option(\"rollout_profile\", 1) first");
  if (int rc = set_device(h)) return rc;
  const int64_t nw = 8 * ((int64_t)h->Kloc / 32 + 2);
  CU(cudaStreamSynchronize(h->st));
  CU(cudaMemcpy(cycles_out, h->d_warp_cycles, sizeof(long long) * (size_t)(n < nw ? n : nw), cudaMemcpyDeviceToHost));
  return 0;
}

int mpopis_b200_last_shrinkage(mpopis_t *h, double *lambda_out) {
  if (!h || !lambda_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemcpy(lambda_out, h->d_lambda, sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int mpopis_b200_cholesky(mpopis_t *h, const double *A, int64_t n, double *L_out) {
  if (!h || !A || !L_out || n < 1 || n > 4096) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dA, dLt, dW;
  DevBuf<int> dinfo;
  int info = 0;
  const size_t nn = (size_t)n * n;
  if (int rc = dA.alloc(nn)) return rc;
  if (int rc = dLt.alloc(nn)) return rc;
  if (int rc = dW.alloc(nn + (size_t)n)) return rc;
  if (int rc = dinfo.alloc(1)) return rc;
  CU(cudaMemcpyAsync(dA, A, sizeof(double) * nn, cudaMemcpyHostToDevice, h->st));
  launch_chol(dA, (int)n, nullptr, dLt, dW, dinfo, 1, nullptr, h->st);
  launch_transpose_sq(dLt, dA, (int)n, h->st);  // row-major L -> column-major L
  h->launches += 2;
  CU(cudaMemcpyAsync(L_out, dA, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  if (info) return fail(MPOPIS_ERR_NOT_PD, "PosDefException: matrix is not positive definite; Cholesky factorization failed.");
  return 0;
}

int mpopis_b200_inv_sqrt(mpopis_t *h, const double *A, int64_t n, double *C_out) {
  if (!h || !A || !C_out || n < 1 || n > 4096) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (int rc = set_device(h)) return rc;
  DevBuf<double> dA, dC, dws;
  DevBuf<int> dinfo;
  int info = 0;
  const size_t nn = (size_t)n * n;
  if (int rc = dA.alloc(nn)) return rc;
  if (int rc = dC.alloc(nn)) return rc;
  if (int rc = dws.alloc(5 * nn + (size_t)((n + 31) / 32) * ((n + 31) / 32) + 8)) return rc;
  if (int rc = dinfo.alloc(1)) return rc;
  CU(cudaMemcpyAsync(dA, A, sizeof(double) * nn, cudaMemcpyHostToDevice, h->st));
  cudaError_t e = (cudaError_t)launch_inv_sqrt(dA, (int)n, dC, dws, dinfo, 1, nullptr, h->coop_max, h->st);
  if (e != cudaSuccess) return fail(MPOPIS_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  h->launches += 1;
  CU(cudaMemcpyAsync(C_out, dC, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  if (info) return fail(MPOPIS_ERR_NOT_PD, "Σ^-0.5: matrix is not positive definite");
  return 0;
}

int mpopis_b200_resident_reset(mpopis_t *h, const double *state, int64_t env_t, const double *U) {
  if (!h || !state || !U) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemsetAsync(h->info(), 0, sizeof(int) * 2, h->st));  // info + accumulated iteration count
  CU(cudaMemsetAsync(h->d_reward, 0, sizeof(double) * 2, h->st));
  if (int rc = upload_inputs(h, state, env_t, U)) return rc;
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

__global__ void accumulate_reward_kernel(const double *reward, double *sum) { *sum += *reward; }

int mpopis_b200_resident_plan(mpopis_t *h, int32_t advance_env) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  if (int rc = plan_step(h)) return rc;
  if (advance_env) {
    env_step_device(h, h->d_control);  // env(act), car_example.jl:205-207
    accumulate_reward_kernel<<<1, 1, 0, h->st>>>(h->d_reward, h->d_reward + 1);  // rew += reward(env), car_example.jl:209
    h->launches += 1;
  }
  CU(cudaMemcpyAsync(h->d_U_orig, h->d_U_next, sizeof(double) * h->cs, cudaMemcpyDeviceToDevice, h->st));
  return 0;
}

int mpopis_b200_resident_read(mpopis_t *h, double *state_out, double *U_out, double *control_out,
                              int32_t *its_run_out) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  return download_outputs(h, U_out, control_out, its_run_out, state_out ? state_out : nullptr);
}

int mpopis_b200_resident_reward_sum(mpopis_t *h, double *sum_out) {
  if (!h || !sum_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  CU(cudaMemcpyAsync(sum_out, h->d_reward + 1, sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return 0;
}

int mpopis_b200_measure_fp64_peak(mpopis_t *h, double *dfma_per_s_out) {
  if (!h || !dfma_per_s_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->dev));
  const int grid = prop.multiProcessorCount * 8;  // 2048 threads / SM
  DevBuf<double> out;
  if (int rc = out.alloc((size_t)grid * 256)) return rc;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  launch_dfma_peak(out, grid, 2000, h->st);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, h->st));
    const long long n = launch_dfma_peak(out, grid, 20000, h->st);
    CU(cudaEventRecord(e1, h->st));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double rate = (double)n / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  h->launches += 6;
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  *dfma_per_s_out = best;
  return 0;
}

// Times the weighted row-sum kernel (G8: Σ_k w_k E[r,k], POL:226-229) alone over this handle's E operand.
// It is the one HBM-bound kernel of the path; with num_samples = 2^20 the operand (839 MB) exceeds L2.
int mpopis_b200_bench_rowsum(mpopis_t *h, int32_t reps, double *ms_per_launch_out, double *bytes_per_launch_out) {
  if (!h || !ms_per_launch_out || reps < 1) return fail(MPOPIS_ERR_BAD_ARG, "bad argument");
  if (int rc = set_device(h)) return rc;
  launch_philox_normals(h->d_E, h->ldk, h->cs, h->Kloc, h->k0, 1234u, 0u, nullptr, 0u, nullptr, h->st);
  CU(cudaMemsetAsync(h->d_costs, 0, sizeof(double) * h->K, h->st));
  h->launches += 1 + launch_weights(h->d_costs, h->K, 1.0, h->d_w, h->d_ones, nullptr, h->st);
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) launch_rowsum_partial(h->d_E, h->ldk, h->cs, h->Kloc, h->d_w + h->k0, h->d_part, nullptr, h->st);
  CU(cudaEventRecord(e0, h->st));
  for (int i = 0; i < reps; ++i)
    launch_rowsum_partial(h->d_E, h->ldk, h->cs, h->Kloc, h->d_w + h->k0, h->d_part, nullptr, h->st);
  CU(cudaEventRecord(e1, h->st));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  h->launches += reps + 2;
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  *ms_per_launch_out = ms / reps;
  if (bytes_per_launch_out) *bytes_per_launch_out = 8.0 * h->cs * h->Kloc + 8.0 * h->Kloc;  // E rows + weights
  return 0;
}

int mpopis_b200_resident_total_its(mpopis_t *h, int64_t *total_its_out) {
  if (!h || !total_its_out) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (int rc = set_device(h)) return rc;
  int v = 0;
  CU(cudaMemcpyAsync(&v, h->d_flags + 3, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  *total_its_out = v;
  return 0;
}

int64_t mpopis_b200_launch_count(mpopis_t *h) { return h ? h->launches : 0; }

int mpopis_b200_last_timing(mpopis_t *h, double *rollout_ms, double *total_ms, int32_t *rollout_launches) {
  if (!h) return fail(MPOPIS_ERR_BAD_ARG, "null argument");
  if (!h->timing_valid) return fail(MPOPIS_ERR_BAD_ARG, "no completed plan to report");
  if (rollout_ms) *rollout_ms = h->last_rollout_ms;
  if (total_ms) *total_ms = h->last_total_ms;
  if (rollout_launches) *rollout_launches = h->last_its_launched;
  return 0;
}

void *mpopis_b200_stream(mpopis_t *h) { return h ? (void *)h->st : nullptr; }

}  // extern "C"
