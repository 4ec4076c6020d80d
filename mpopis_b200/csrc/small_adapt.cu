// small_adapt.cu — the whole cross-entropy adaptation of ONE AIS iteration (POL:455-465 + the next iteration's
// MvNormal(Σ′), POL:447) as a single CTA, for the reference's own problem sizes (car_example.jl:57: K = 150, m = 30,
// cs = 100).
//
// At these sizes the adaptation is four dependent single-CTA latency kernels — stable sort + early-stop test, elite
// moments, shrinkage, Cholesky — 103 µs per iteration next to a 90 µs rollout launch (phase trace, K = 150), most of it
// launch gaps and global-memory round trips of a 100 x 100 matrix between kernels that each use one SM. Here the matrix
// never leaves the SM: the centred elite columns are staged once in shared memory, the scatter matrix is accumulated
// DIRECTLY in the 16 x 16 thread grid's register tile that the factorisation starts from (chol_reg_kernel's layout,
// linalg.cu), the shrinkage sums are block reductions over that tile, and L is written out for E = L·Z.
//   1. bitonic sort of the composite (cost key, sample id) keys — Julia's stable sortperm (sort.cu)
//   2. maximum(abs.(diff(elite costs))) < 10e-3 -> raise the stop flag and leave (POL:458-461)
//   3. X = E[:, order[1:m]], μ′ = mean, pol.U += μ′, X −= μ′                      (POL:462-465, mean_and_cov)
//   4. S = X Xᵀ / m in registers; λ̂ of :lw / :ss / :rblw / :oas; Σ′ = shrink(S) + 1e-8 I
//   5. right-looking register-tiled Cholesky of Σ′ (identical to chol_cov_kernel) -> Lt
// Same formulas and summation structure as moments_small_kernel + chol_cov_kernel; tests/test_gpu_parity.py pins the
// control step against the oracle at these sizes, with the fused kernel on (default) and off ("ce_small_fused" = 0).
#include "chol_tile.cuh"
#include "engine.cuh"

namespace mpopis {

namespace {

constexpr int SA_MAXK = 512, SA_MAXM = 128;
constexpr unsigned long long SA_KEY_PAD = ~0ULL;

__device__ __forceinline__ bool sa_lt(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}
// maximum(abs.(diff(...))) propagates NaN in Julia (then `NaN < 10e-3` is false): fmax alone would drop it
__device__ __forceinline__ double sa_gap_max(double mx, double d) { return (d != d || mx != mx) ? d + mx : fmax(mx, d); }
__device__ __forceinline__ double sa_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double sa_block_sum(double v, double *red) {  // 256 threads, fixed order
  v = sa_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

template <int R>
__global__ void __launch_bounds__(256) ce_small_adapt_kernel(
    const double *__restrict__ costs, int K, int tn, int m, int early_stop, const double *__restrict__ E, long long ldk,
    int n, int method, double ridge, int pitch, unsigned long long *__restrict__ keys_out, int *__restrict__ order_out,
    double *__restrict__ mu_out, double *__restrict__ U_cur, double *__restrict__ sums_out, double *__restrict__ Sigma,
    double *__restrict__ Lt, double *lambda_out, int *info, int tag, int *stop_flag) {
  if (*stop_flag) return;
  extern __shared__ double Xs[];  // [n][pitch]: the centred elite columns
  __shared__ unsigned long long sk[SA_MAXK];
  __shared__ int sv[SA_MAXK];
  __shared__ CholTileSmem<R> sm;
  __shared__ double dinv[16 * R], red[8];
  __shared__ int s_stop;
  double *dg = sm.dg;  // diag(S) until the factorisation takes the array over
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  // ---- 1. order = sortperm(costs) ----
  for (int e = threadIdx.x; e < tn; e += 256) {
    sk[e] = e < K ? cost_key(costs[e]) : SA_KEY_PAD;
    sv[e] = e < K ? e : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= tn; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      if (t < (tn >> 1)) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long ka = sk[i], kb = sk[q];
        const int va = sv[i], vb = sv[q];
        if (sa_lt(kb, vb, ka, va) == up) sk[i] = kb, sv[i] = vb, sk[q] = ka, sv[q] = va;
      }
      __syncthreads();
    }
  for (int e = threadIdx.x; e < K; e += 256) keys_out[e] = sk[e], order_out[e] = sv[e];

  // ---- 2. early stop on the elite costs (POL:458-461) ----
  {
    double mx = -1.0;
    for (int j = threadIdx.x; j + 1 < m; j += 256) mx = sa_gap_max(mx, fabs(key_cost(sk[j + 1]) - key_cost(sk[j])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = sa_gap_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[wid] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) mx = sa_gap_max(mx, red[w]);
      s_stop = early_stop && mx < 10e-3;
      if (s_stop) *stop_flag = 1;
    }
    __syncthreads();
    if (s_stop) return;
  }

  // ---- 3. elite columns (all n·m gathered loads in flight at once), then per row: mean, pol.U += μ′, centring ----
  const double cnt = (double)m, inv = 1.0 / cnt;
  for (int e = threadIdx.x; e < n * m; e += 256) {
    const int r = e / m, s = e - r * m;
    Xs[r * pitch + s] = E[(size_t)r * ldk + sv[s]];
  }
  __syncthreads();
  for (int r = wid; r < n; r += 8) {  // one warp per row, lanes over the elites
    double a = 0.0;
    for (int s = lane; s < m; s += 32) a = fma(1.0, Xs[r * pitch + s], a);
    a = sa_warp_sum(a);
    const double mean = a / cnt;
    for (int s = lane; s < m; s += 32) Xs[r * pitch + s] -= mean;
    if (lane == 0) {
      if (mu_out) mu_out[r] = mean;
      sums_out[r] = a;
      U_cur[r] = U_cur[r] + mean;  // pol.U = pol.U + vec(μ′), POL:465
    }
  }
  if (threadIdx.x == 0) sums_out[n] = cnt;
  __syncthreads();

  // ---- 4. S = X Xᵀ / m straight into the factorisation's register tile: thread (ty, tx) owns S[ty + 16a][tx + 16b] ----
  double w[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) w[a][b] = 0.0;
  for (int s = 0; s < m; ++s) {
    double xi[R], xk[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
      xi[a] = (ty + 16 * a < n) ? Xs[(ty + 16 * a) * pitch + s] : 0.0;
      xk[a] = (tx + 16 * a < n) ? Xs[(tx + 16 * a) * pitch + s] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) w[a][b] = fma(xi[a], xk[b], w[a][b]);
  }
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      w[a][b] = (i < n && k <= i) ? w[a][b] * inv : 0.0;  // lower triangle only, like chol_reg_kernel
      if (i < n && k == i) dg[i] = w[a][b];
    }
  __syncthreads();
  double lam = 0.0, F = 0.0;
  if (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS) {
    const bool ss = method == MPOPIS_SIGMA_SS;
    for (int i = threadIdx.x; i < n; i += 256) dinv[i] = ss ? 1.0 / sqrt(dg[i]) : 1.0;
    __syncthreads();
    // Σ_{i≠j} Σ_s (z_si z_sj)² = Σ_s [(Σ_i z_si²)² − Σ_i z_si⁴], z = (x − μ)/σ: one warp per elite
    double q = 0.0;
    for (int s = wid; s < m; s += 8) {
      double a = 0.0, b = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double z = Xs[i * pitch + s] * dinv[i];
        const double z2 = z * z;
        a += z2;
        b = fma(z2, z2, b);
      }
      a = sa_warp_sum(a), b = sa_warp_sum(b);
      if (lane == 0) q += a * a - b;
    }
    q = sa_block_sum(q, red);
    double r2 = 0.0;  // Σ_{i≠j} (s_ij d_i d_j)²: twice the strictly-lower sum
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        const int i = ty + 16 * a, k = tx + 16 * b;
        if (i < n && k < i) {
          const double v = w[a][b] * dinv[i] * dinv[k];
          r2 = fma(v, v, r2);
        }
      }
    r2 = 2.0 * sa_block_sum(r2, red);
    const double num = (q - cnt * r2) * cnt / ((cnt - 1.0) * cnt * cnt);
    lam = fmin(fmax(num / r2, 0.0), 1.0);
  } else if (method == MPOPIS_SIGMA_RBLW || method == MPOPIS_SIGMA_OAS) {
    double tr = 0.0, tr2 = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        const int i = ty + 16 * a, k = tx + 16 * b;
        if (i < n && k <= i) {
          const double v = w[a][b];
          tr2 = fma(k == i ? 1.0 : 2.0, v * v, tr2);
          if (k == i) tr += v;
        }
      }
    tr = sa_block_sum(tr, red);
    tr2 = sa_block_sum(tr2, red);
    const double pd = (double)n, trsq = tr * tr;
    if (method == MPOPIS_SIGMA_RBLW) lam = ((cnt - 2) / cnt * tr2 + trsq) / ((cnt + 2) * (tr2 - trsq / pd));
    else lam = ((1.0 - 2.0 / pd) * tr2 + trsq) / ((cnt + 1.0 - 2.0 / pd) * (tr2 - trsq / pd));
    lam = fmin(fmax(lam, 0.0), 1.0);
    F = tr / pd;
  }
  const bool common = method == MPOPIS_SIGMA_RBLW || method == MPOPIS_SIGMA_OAS;
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      if (i < n && k <= i) {
        const double v = w[a][b];
        const double sgm = k == i ? (common ? (1.0 - lam) * v + lam * F : v) + ridge : (1.0 - lam) * v;
        w[a][b] = sgm;
        Sigma[(size_t)i * n + k] = sgm;  // kept for fetch_proposal
        Sigma[(size_t)k * n + i] = sgm;
      }
    }
  if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
  __syncthreads();

  // ---- 5. Cholesky of Σ′ in the register tile (chol_tile.cuh) ----
  if (!chol_tile_factor<R>(w, n, sm, tx, ty)) {
    if (threadIdx.x == 0) atomicCAS(info, 0, tag);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Lt[e] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  chol_tile_finish<R>(sm, n);
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      if (i < n && k < n) Lt[(size_t)i * n + k] = k <= i ? chol_tile_entry<R>(w, sm, a, b, i, k) : 0.0;
    }
}

template <int R>
int launch_sa(const double *costs, int K, int tn, int m, int early_stop, const double *E, long long ldk, int n, int method,
              double ridge, unsigned long long *keys_out, int *order_out, double *mu_out, double *U_cur, double *sums_out,
              double *Sigma, double *Lt, double *lambda_out, int *info, int tag, int *stop_flag, cudaStream_t s) {
  const int pitch = m | 1;  // odd: the 16 rows a half-warp reads for one elite sit in distinct banks
  const size_t smem = sizeof(double) * (size_t)n * pitch;
  cudaFuncSetAttribute(ce_small_adapt_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);  // per device
  ce_small_adapt_kernel<R><<<1, 256, smem, s>>>(costs, K, tn, m, early_stop, E, ldk, n, method, ridge, pitch, keys_out,
                                                order_out, mu_out, U_cur, sums_out, Sigma, Lt, lambda_out, info, tag,
                                                stop_flag);
  return 1;
}

}  // namespace

// returns 0 when the sizes are beyond the fused kernel (the caller then runs sort, moments and Cholesky separately)
int launch_ce_small_adapt(const double *costs, int K, int m, int early_stop, const double *E, long long ldk, int n,
                          int method, double ridge, unsigned long long *keys_out, int *order_out, double *mu_out,
                          double *U_cur, double *sums_out, double *Sigma, double *Lt, double *lambda_out, int *info, int tag,
                          int *stop_flag, cudaStream_t s) {
  if (K > SA_MAXK || m > SA_MAXM || m < 2 || n > 112 || (size_t)n * (m | 1) * sizeof(double) > 118 * 1024) return 0;
  int tn = 32;
  while (tn < K) tn <<= 1;
#define MPOPIS_SA(R) launch_sa<R>(costs, K, tn, m, early_stop, E, ldk, n, method, ridge, keys_out, order_out, mu_out, U_cur, sums_out, Sigma, Lt, lambda_out, info, tag, stop_flag, s)
  if (n <= 16) return MPOPIS_SA(1);
  if (n <= 64) return MPOPIS_SA(4);
  return MPOPIS_SA(7);
#undef MPOPIS_SA
}

}  // namespace mpopis
