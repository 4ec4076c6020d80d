// linalg.cu — G6: the small dense factorizations of the AIS loop.
//
//   chol      lower Cholesky factor of Σ′ (what MvNormal(Σ′) builds: POL:154,192,307,352,447,551,650,
//             723,796). Failure mirrors Julia's PosDefException -> MPOPIS_ERR_NOT_PD.
//   ctrl_vec  b = γ Σ′⁻¹ U_orig by two triangular solves (replaces invcov(P), POL:156,...,798, which
//             the reference only ever uses inside γ U_origᵀ Σ⁻¹ (V − U_orig), POL:272).
// cs is 15..300 for the reference's configs, so these are single-CTA latency kernels; every rank of
// a sharded policy runs them redundantly on bit-identical inputs.
#include "chol_tile.cuh"
#include "engine.cuh"

namespace mpopis {

// Right-looking Cholesky on W (row-major n x n, lower part used). W lives in shared memory when it
// fits (n <= CHOL_SMEM_N) and in global scratch otherwise. Output Lt = L stored row-major with an
// explicit zero upper triangle. `sigma_dev` (nullable): factor σ²·A instead (CMA, POL:551).
// info: set to `tag` (first failure wins) if a pivot is not > 0.
__global__ void __launch_bounds__(1024) chol_kernel(const double *__restrict__ A, int n, const double *sigma_dev,
                                                     double *__restrict__ Lt, double *__restrict__ Wglobal,
                                                     int use_smem, int *info, int tag, const int *stop) {
  if (stop && *stop) return;
  extern __shared__ double Wsm[];
  double *W = use_smem ? Wsm : Wglobal;
  const int ldw = n | 1;  // odd row pitch: the column reads W[k][j] (k across lanes) are bank-conflict free
  const double sc = sigma_dev ? (*sigma_dev) * (*sigma_dev) : 1.0;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x)
    W[(size_t)(e / n) * ldw + e % n] = sigma_dev ? sc * A[e] : A[e];
  // Square-root-free right-looking elimination with ONE barrier per column: column j is left unscaled
  // (W[i][j] = l_ij * l_jj) and the trailing update divides by the pivot d_j = l_jj², so there is no
  // separate scaling phase; L is formed at the end as W[i][j] / sqrt(d_j). Threads form a 32 x 32 grid
  // over (row, column) of the trailing block — no integer divisions in the loop.
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  bool failed = false;
  for (int j = 0; j < n; ++j) {
    const double d = W[(size_t)j * ldw + j];  // final after the barrier that ended step j-1
    if (!(d > 0.0)) {                         // uniform: every thread reads the same value
      failed = true;
      break;
    }
    const double inv_d = 1.0 / d;
    for (int i = j + 1 + ty; i < n; i += 32) {
      const double wij = W[(size_t)i * ldw + j] * inv_d;
      for (int k = j + 1 + tx; k <= i; k += 32) W[(size_t)i * ldw + k] -= wij * W[(size_t)k * ldw + j];
    }
    __syncthreads();
  }
  if (failed) {
    if (threadIdx.x == 0) atomicCAS(info, 0, tag);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Lt[e] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, k = e % n;
    double v = 0.0;
    if (k <= i) {
      const double r = sqrt(W[(size_t)k * ldw + k]);
      v = k == i ? r : W[(size_t)i * ldw + k] / r;
    }
    Lt[e] = v;
  }
}

// Register-tiled variant for n <= 16 R (the reference's control sizes 15 and 100): the matrix lives in the
// registers of a 16 x 16 thread grid (thread (ty,tx) owns W[ty+16a][tx+16b]); per column only the pivot
// column travels through shared memory (double-buffered: ONE barrier per column) and the trailing update
// is R² predicated register FMAs with compile-time indices. The shared-memory kernel above spent its time
// issuing loop/index overhead from 32 warps on one SM (ncu: 84 µs for n = 100); this one is ~5x shorter.
template <int R>
__global__ void __launch_bounds__(256) chol_reg_kernel(const double *__restrict__ A, int n, const double *sigma_dev,
                                                        double *__restrict__ Lt, int *info, int tag,
                                                        const int *stop) {
  if (stop && *stop) return;
  __shared__ CholTileSmem<R> sm;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const double sc = sigma_dev ? (*sigma_dev) * (*sigma_dev) : 1.0;
  double w[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      w[a][b] = (i < n && k <= i) ? sc * A[(size_t)i * n + k] : 0.0;  // A symmetric: row-major == column-major
    }
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

// cov_finalize (stats.cu) + chol_reg in ONE launch for the :cemppi update: Σ′ = shrink(method, Sraw/denom) + ridge·I is
// formed in the registers that the factorisation starts from — the p x p matrix is read once, the trace / correlation
// sums of the shrinkage intensity are block reductions over the register tile, and Σ′ is written out only for
// fetch_proposal. Replaces two single-CTA latency kernels (≈16 + 31 µs at p = 100) and the launch gap between them.
// Same formulas as cov_finalize_kernel; the sums run in register-tile order instead of row order.
__device__ __forceinline__ double block_sum_256(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

template <int R>
__global__ void __launch_bounds__(256) chol_cov_kernel(const double *__restrict__ Sraw, int n, const double *cnt_dev,
                                                        int corrected, int method, const double *__restrict__ q_dev,
                                                        double ridge, double *__restrict__ Sigma,
                                                        double *__restrict__ Lt, double *lambda_out, int *info, int tag,
                                                        const int *stop) {
  if (stop && *stop) return;
  __shared__ CholTileSmem<R> sm;
  __shared__ double dinv[16 * R];
  __shared__ double red[8];
  double *dg = sm.dg;  // diag(S) until the factorisation takes the array over
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const double cnt = *cnt_dev, inv = 1.0 / (cnt - (corrected ? 1.0 : 0.0));
  double w[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      w[a][b] = (i < n && k <= i) ? Sraw[(size_t)i * n + k] * inv : 0.0;  // S = Sraw / denom, lower triangle
      if (i < n && k == i) dg[i] = w[a][b];
    }
  __syncthreads();
  double lam = 0.0, F = 0.0;
  if (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS) {
    const bool ss = method == MPOPIS_SIGMA_SS;
    for (int i = threadIdx.x; i < n; i += 256) dinv[i] = ss ? 1.0 / sqrt(dg[i]) : 1.0;  // 1/σ_i once, not per element
    __syncthreads();
    double r2 = 0.0;  // Σ_{i≠j} (s_ij d_i d_j)², d = 1/σ (:ss) or 1 (:lw): twice the strictly-lower sum
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
    r2 = 2.0 * block_sum_256(r2, red);
    const double num = (*q_dev - cnt * r2) * cnt / ((cnt - 1.0) * cnt * cnt);
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
    tr = block_sum_256(tr, red);
    tr2 = block_sum_256(tr2, red);
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
        // diag(S) target: the diagonal is kept; tr(S)/p·I target: it is shrunk as well (cov_finalize_kernel)
        const double s = k == i ? (common ? (1.0 - lam) * v + lam * F : v) + ridge : (1.0 - lam) * v;
        w[a][b] = s;
        Sigma[(size_t)i * n + k] = s;
        Sigma[(size_t)k * n + i] = s;
      }
    }
  if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
  __syncthreads();
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

// returns 0 when n is beyond the register-tiled kernels (the caller then runs cov_finalize + chol separately)
int launch_chol_cov(const double *Sraw, int n, const double *cnt_dev, int corrected, int method, const double *q_dev,
                    double ridge, double *Sigma, double *Lt, double *lambda_out, int *info, int tag, const int *stop,
                    cudaStream_t s) {
#define MPOPIS_CC(R) chol_cov_kernel<R><<<1, 256, 0, s>>>(Sraw, n, cnt_dev, corrected, method, q_dev, ridge, Sigma, Lt, lambda_out, info, tag, stop)
  if (n <= 16) MPOPIS_CC(1);
  else if (n <= 64) MPOPIS_CC(4);
  else if (n <= 112) MPOPIS_CC(7);
  else if (n <= 160) MPOPIS_CC(10);
  else return 0;
#undef MPOPIS_CC
  return 1;
}

// Blocked Cholesky for 160 < n <= 430 (MultiCarRacing: cs = 200 / 300), still ONE CTA — the matrix (720 KB at n = 300)
// fits neither the register tile nor shared memory, and the unblocked kernel above walked it in L2 at 5.8 µs per column
// (1.75 ms per factorisation at n = 300: 42 % of a 3-car :cmamppi control step, profiles/r2 notes). Left to right in
// panels of 64 columns, working in place in Lt:
//   1. the 64 x 64 diagonal block is factored in the register tile (chol_tile.cuh, R = 4) and kept in shared memory;
//   2. the rows below solve X·L_bbᵀ = A_panel, ONE THREAD PER ROW, 16 entries at a time in registers (L_bb entries are
//      warp-uniform shared-memory broadcasts, a thread re-reads only its own row of the panel);
//   3. the trailing matrix gets A −= X·Xᵀ from the panel in shared memory: 64 x 64 tiles, a 4 x 4 register tile per
//      thread with rows/columns strided by 16 (conflict-free with the odd pitch).
// ≈ n³/3 FMAs on one SM = 70 µs of DFMA issue at n = 300, plus five 64-column tile factorisations.
constexpr int CB_NB = 64, CB_P = 65, CB_MAX_N = 430;

__global__ void __launch_bounds__(256) chol_blocked_kernel(const double *__restrict__ A, int n, const double *sigma_dev,
                                                            double *__restrict__ Lt, int *info, int tag, const int *stop) {
  if (stop && *stop) return;
  extern __shared__ double cb[];  // Lbb[64][65] | X[n − 64][65]
  __shared__ CholTileSmem<4> sm;
  double *Lbb = cb, *X = cb + CB_NB * CB_P;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const double sc = sigma_dev ? (*sigma_dev) * (*sigma_dev) : 1.0;
  for (int e = threadIdx.x; e < n * n; e += 256) {
    const int i = e / n, k = e - i * n;
    Lt[e] = k <= i ? sc * A[e] : 0.0;
  }
  __syncthreads();
  bool ok = true;
  for (int c0 = 0; c0 < n; c0 += CB_NB) {
    const int nb = min(CB_NB, n - c0), m = n - c0 - nb;
    // ---- 1. diagonal block ----
    double w[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 16 * a, k = tx + 16 * b;
        w[a][b] = (i < nb && k <= i) ? Lt[(size_t)(c0 + i) * n + c0 + k] : 0.0;
      }
    if (!chol_tile_factor<4>(w, nb, sm, tx, ty)) {  // block-uniform
      ok = false;
      break;
    }
    chol_tile_finish<4>(sm, nb);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 16 * a, k = tx + 16 * b;
        const double v = (i < nb && k <= i) ? chol_tile_entry<4>(w, sm, a, b, i, k) : 0.0;
        Lbb[i * CB_P + k] = v;
        if (i < nb && k <= i) Lt[(size_t)(c0 + i) * n + c0 + k] = v;
      }
    __syncthreads();
    if (m == 0) break;  // (only the last panel can be narrower than 64, and it has no rows below)
    // ---- 2. rows below: X L_bbᵀ = A_panel, one thread per row ----
    for (int r = threadIdx.x; r < m; r += 256) {
      double *row = Lt + (size_t)(c0 + nb + r) * n + c0;
      double *xr = X + (size_t)r * CB_P;
      for (int jb = 0; jb < CB_NB; jb += 16) {  // 16 entries at a time in registers
        double acc[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = row[jb + t];
        for (int k = 0; k < jb; ++k) {  // the entries already solved (read back from this thread's own panel row)
          const double xk = xr[k];
#pragma unroll
          for (int t = 0; t < 16; ++t) acc[t] = fma(-xk, Lbb[(jb + t) * CB_P + k], acc[t]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double xj = acc[j] * sm.col[1][jb + j];  // 1 / l_jj
          acc[j] = xj;
#pragma unroll
          for (int t = j + 1; t < 16; ++t) acc[t] = fma(-xj, Lbb[(jb + t) * CB_P + jb + j], acc[t]);
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) row[jb + t] = acc[t], xr[jb + t] = acc[t];
      }
    }
    __syncthreads();
    // ---- 3. trailing update A −= X Xᵀ (lower tiles) ----
    const int nt = (m + 63) / 64;
    for (int ti = 0; ti < nt; ++ti)
      for (int tj = 0; tj <= ti; ++tj) {
        double c[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
        const double *Xi = X + (size_t)(ti * 64 + ty) * CB_P, *Xj = X + (size_t)(tj * 64 + tx) * CB_P;
        bool vi[4], vj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) vi[a] = ti * 64 + ty + 16 * a < m, vj[a] = tj * 64 + tx + 16 * a < m;
#pragma unroll 4
        for (int k = 0; k < CB_NB; ++k) {
          double xi[4], xj[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            xi[a] = vi[a] ? Xi[(size_t)16 * a * CB_P + k] : 0.0;
            xj[a] = vj[a] ? Xj[(size_t)16 * a * CB_P + k] : 0.0;
          }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) c[a][b] = fma(xi[a], xj[b], c[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int i = ti * 64 + ty + 16 * a, j = tj * 64 + tx + 16 * b;
            if (i < m && j <= i) Lt[(size_t)(c0 + nb + i) * n + c0 + nb + j] -= c[a][b];
          }
      }
    __syncthreads();
  }
  if (!ok) {
    if (threadIdx.x == 0) atomicCAS(info, 0, tag);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Lt[e] = __longlong_as_double(0x7ff8000000000000LL);
  }
}

constexpr int CHOL_SMEM_N = 160;  // 160 · 161 · 8 B = 201 KB

// Wglobal: scratch of n (n|1) doubles, used when n > CHOL_SMEM_N
void launch_chol(const double *A, int n, const double *sigma_dev, double *Lt, double *Wglobal, int *info, int tag,
                 const int *stop, cudaStream_t s) {
  if (n <= 16) return (void)chol_reg_kernel<1><<<1, 256, 0, s>>>(A, n, sigma_dev, Lt, info, tag, stop);
  if (n <= 64) return (void)chol_reg_kernel<4><<<1, 256, 0, s>>>(A, n, sigma_dev, Lt, info, tag, stop);
  if (n <= 112) return (void)chol_reg_kernel<7><<<1, 256, 0, s>>>(A, n, sigma_dev, Lt, info, tag, stop);
  if (n <= 160) return (void)chol_reg_kernel<10><<<1, 256, 0, s>>>(A, n, sigma_dev, Lt, info, tag, stop);
  if (n <= CB_MAX_N) {
    const size_t smem = sizeof(double) * (size_t)n * CB_P;
    cudaFuncSetAttribute(chol_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(double) * CB_MAX_N * CB_P));  // per device: set on every launch
    return (void)chol_blocked_kernel<<<1, 256, smem, s>>>(A, n, sigma_dev, Lt, info, tag, stop);
  }
  const int use_smem = n <= CHOL_SMEM_N;
  const size_t smem = use_smem ? sizeof(double) * n * (n | 1) : 0;
  cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(sizeof(double) * CHOL_SMEM_N * (CHOL_SMEM_N | 1)));  // per device: set on every launch
  chol_kernel<<<1, 1024, smem, s>>>(A, n, sigma_dev, Lt, Wglobal, use_smem, info, tag, stop);
}

// b = γ (L Lᵀ)⁻¹ u : forward then backward substitution, column-oriented (one barrier per column).
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double *__restrict__ Lt, int n,
                                                           const double *__restrict__ u, double gamma,
                                                           double *__restrict__ b, const int *stop) {
  if (stop && *stop) return;
  extern __shared__ double y[];  // n
  for (int i = threadIdx.x; i < n; i += blockDim.x) y[i] = u[i];
  __syncthreads();
  for (int j = 0; j < n; ++j) {  // L y = u
    if (threadIdx.x == 0) y[j] = y[j] / Lt[(size_t)j * n + j];
    __syncthreads();
    const double yj = y[j];
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) y[i] -= Lt[(size_t)i * n + j] * yj;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {  // Lᵀ x = y
    if (threadIdx.x == 0) y[j] = y[j] / Lt[(size_t)j * n + j];
    __syncthreads();
    const double xj = y[j];
    for (int i = threadIdx.x; i < j; i += blockDim.x) y[i] -= Lt[(size_t)j * n + i] * xj;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = gamma * y[i];
}

void launch_chol_solve(const double *Lt, int n, const double *u, double gamma, double *b, const int *stop,
                       cudaStream_t s) {
  chol_solve_kernel<<<1, 1024, sizeof(double) * n, s>>>(Lt, n, u, gamma, b, stop);
}

// plain transposes of small square matrices (Lt <-> column-major L at the ABI boundary)
__global__ void transpose_sq_kernel(const double *__restrict__ in, double *__restrict__ out, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n) return;
  out[(size_t)(e % n) * n + e / n] = in[e];
}
void launch_transpose_sq(const double *in, double *out, int n, cudaStream_t s) {
  transpose_sq_kernel<<<(n * n + 255) / 256, 256, 0, s>>>(in, out, n);
}

}  // namespace mpopis
