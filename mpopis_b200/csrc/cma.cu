// cma.cu — G9: CMA-ES path / step-size / covariance update of :cmamppi (POL:571-599) and the
// matrix function Σ^-0.5 (POL:580).
//
// Σ^-0.5 is computed by the coupled Newton–Schulz iteration on A/‖A‖_F
//     T = (3I − Z Y)/2,  Y ← Y T,  Z ← T Z,   Y → (A/c)^{1/2},  Z → (A/c)^{-1/2}
// which is three n x n x n contractions per step — GEMM-shaped work that spreads over the chip,
// unlike a Jacobi eigensolver. The whole iteration runs inside ONE cooperative kernel (grid-wide
// barriers between the products, convergence decided identically by every CTA), so a CMA
// iteration costs one launch instead of hundreds. It converges for every SPD matrix; for an
// indefinite Σ (where the reference's Σ^-0.5 turns complex and the following MvNormal throws) it
// does not, and the kernel raises the handle's info flag -> MPOPIS_ERR_NOT_PD.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace mpopis {

namespace {
__device__ __forceinline__ double block_sum(double v, double *red /* >= 33 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// one 32x32 tile of C = alpha * A * B + beta * I  (row-major n x n), 256 threads, 2x2 per thread
__device__ __forceinline__ double gemm_tile(const double *__restrict__ A, const double *__restrict__ B, int n,
                                          int ti, int tj, double alpha, double beta, double *__restrict__ C,
                                          double (*As)[33], double (*Bs)[33]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = ti * 32, j0 = tj * 32;
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
  for (int k0 = 0; k0 < n; k0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      As[r][c] = (i0 + r < n && k0 + c < n) ? A[(size_t)(i0 + r) * n + k0 + c] : 0.0;
      Bs[r][c] = (k0 + r < n && j0 + c < n) ? B[(size_t)(k0 + r) * n + j0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double a0 = As[2 * ty][kk], a1 = As[2 * ty + 1][kk];
      const double b0 = Bs[kk][2 * tx], b1 = Bs[kk][2 * tx + 1];
      c00 = fma(a0, b0, c00), c01 = fma(a0, b1, c01), c10 = fma(a1, b0, c10), c11 = fma(a1, b1, c11);
    }
  }
  const int i = i0 + 2 * ty, j = j0 + 2 * tx;
  double res = 0.0;  // this thread's share of ‖C − I‖_F²
  auto put = [&](int ii, int jj, double acc) {
    if (ii < n && jj < n) {
      const double v = alpha * acc + (ii == jj ? beta : 0.0);
      C[(size_t)ii * n + jj] = v;
      const double d = v - (ii == jj ? 1.0 : 0.0);
      res = fma(d, d, res);
    }
  };
  put(i, j, c00), put(i, j + 1, c01), put(i + 1, j, c10), put(i + 1, j + 1, c11);
  return res;
}
}  // namespace

constexpr int NS_MAX_IT = 100;

// ws: 5 n² doubles (Y, Z, T, Y2, Z2) + one per 32 x 32 tile. Launched cooperatively with <= (#tiles) CTAs of 256 threads.
__global__ void __launch_bounds__(256) inv_sqrt_ns_kernel(const double *__restrict__ A, int n,
                                                           double *__restrict__ Cout, double *__restrict__ ws,
                                                           int *info, int tag, const int *stop) {
  if (stop && *stop) return;  // grid-uniform
  cg::grid_group grid = cg::this_grid();
  __shared__ double As[32][33], Bs[32][33], red[33];
  const size_t nn = (size_t)n * n;
  double *Y = ws, *Z = ws + nn, *T = ws + 2 * nn, *Y2 = ws + 3 * nn, *Z2 = ws + 4 * nn, *tile_res = ws + 5 * nn;
  const int nt = (n + 31) / 32, ntiles = nt * nt;
  // c = ‖A‖_F, computed redundantly (and identically) by every CTA
  double s = 0.0;
  for (size_t e = threadIdx.x; e < nn; e += blockDim.x) s = fma(A[e], A[e], s);
  const double c = sqrt(block_sum(s, red));
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    Y[e] = A[e] / c;
    Z[e] = (e / n == e % n) ? 1.0 : 0.0;
  }
  grid.sync();
  bool ok = false;
  for (int it = 0; it < NS_MAX_IT; ++it) {
    // ‖T − I‖_F² per tile while the tile is in registers (every CTA used to re-read all n² entries of T for it: ~10 µs of
    // the ~75 µs an iteration took at n = 300), summed in tile order after the barrier: identical in every CTA
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const double part = block_sum(gemm_tile(Z, Y, n, t / nt, t % nt, -0.5, 1.5, T, As, Bs), red);
      if (threadIdx.x == 0) tile_res[t] = part;
    }
    grid.sync();
    double r = 0.0;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) r += __ldcg(tile_res + t);
    r = block_sum(r, red);
    if (!(r == r)) break;                           // NaN: diverged
    if (r < 1e-28 * (double)n) { ok = true; break; }  // ‖T − I‖_F < 1e-14 sqrt(n)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      gemm_tile(Y, T, n, t / nt, t % nt, 1.0, 0.0, Y2, As, Bs);
      gemm_tile(T, Z, n, t / nt, t % nt, 1.0, 0.0, Z2, As, Bs);
    }
    grid.sync();
    double *tmp = Y; Y = Y2; Y2 = tmp;
    tmp = Z; Z = Z2; Z2 = tmp;
  }
  const double isc = 1.0 / sqrt(c);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
    Cout[e] = ok ? Z[e] * isc : __longlong_as_double(0x7ff8000000000000LL);
  if (!ok && blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(info, 0, tag);
}

int inv_sqrt_max_ctas(int num_sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inv_sqrt_ns_kernel, 256, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return per_sm * num_sms;
}

int launch_inv_sqrt(const double *A, int n, double *Cout, double *ws, int *info, int tag, const int *stop,
                    int max_ctas, cudaStream_t s) {
  const int nt = (n + 31) / 32;
  int grid = nt * nt;
  if (grid > max_ctas) grid = max_ctas;
  void *args[] = {(void *)&A, (void *)&n, (void *)&Cout, (void *)&ws, (void *)&info, (void *)&tag, (void *)&stop};
  return (int)cudaLaunchCooperativeKernel((const void *)inv_sqrt_ns_kernel, dim3(grid), dim3(256), args, 0, s);
}

// ---- FP64 pipe roofline denominator ------------------------------------------------------------
// The rollout kernel is bound by FP64 instruction issue, and MEASURED_PEAKS.json has no FP64 figure,
// so bench.py measures it in the same run: 8 independent DFMA chains per thread, all SMs busy.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
    x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// returns DFMA thread-instructions launched; out must hold grid*256 doubles
long long launch_dfma_peak(double *out, int grid, int iters, cudaStream_t s) {
  dfma_peak_kernel<<<grid, 256, 0, s>>>(out, iters, 0.999999, 1e-9);
  return (long long)grid * 256 * iters * 8;
}

// d_ii = elite_E[order[ii]] for ii < K: `order[ii]` is used as a LINEAR index into the cs x m elite
// matrix (SURVEY App. B-1 — the reference's "rank-μ" term is a scalar). Columns of X not owned by
// this shard hold zeros and are summed in by the all-reduce.
__global__ void cma_lin_gather_kernel(const double *__restrict__ X, long long ldx, int cs,
                                      const int *__restrict__ order, int K, double *__restrict__ dvec,
                                      const int *stop) {
  if (stop && *stop) return;
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= K) return;
  const int l = order[ii];
  dvec[ii] = X[(size_t)(l % cs) * ldx + l / cs];
}

void launch_cma_lin_gather(const double *X, long long ldx, int cs, const int *order, int K, double *dvec,
                           const int *stop, cudaStream_t s) {
  cma_lin_gather_kernel<<<(K + 255) / 256, 256, 0, s>>>(X, ldx, cs, order, K, dvec, stop);
}

// All vector-sized CMA updates in one CTA. dw[r] = Σ_j ws[j] elite_E[r,j] (POL:573-576).
__global__ void __launch_bounds__(1024) cma_vec_kernel(const double *__restrict__ dw, const double *__restrict__ C,
                                                        const double *__restrict__ dvec,
                                                        const double *__restrict__ ws, int K, int cs, int n_iter,
                                                        const mpopis_cma_t c, double *__restrict__ psig,
                                                        double *__restrict__ pSig, double *__restrict__ sigma_dev,
                                                        double *__restrict__ U, double *__restrict__ Sigma,
                                                        const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  const double sigma_ds = *sigma_dev;  // δs = elite_E / σ (POL:572) uses σ before POL:582 updates it
  const double cf = sqrt(c.c_sigma * (2 - c.c_sigma) * c.mu_eff);
  double nps2 = 0.0;
  for (int i = threadIdx.x; i < cs; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < cs; ++j) s = fma(C[(size_t)i * cs + j], dw[j], s);  // C δw (C symmetric)
    const double p = (1 - c.c_sigma) * psig[i] + cf * s;                    // POL:581
    psig[i] = p;
    nps2 = fma(p, p, nps2);
    U[i] += sigma_ds * dw[i];  // POL:577
  }
  nps2 = block_sum(nps2, red);
  double nc2 = 0.0;
  for (int e = threadIdx.x; e < cs * cs; e += blockDim.x) nc2 = fma(C[e], C[e], nc2);
  nc2 = block_sum(nc2, red);
  const double nps = sqrt(nps2), normC = sqrt(nc2);
  const int hs = nps / sqrt(1 - pow(1 - c.c_sigma, 2.0 * (double)n_iter)) <
                 (1.4 + 2.0 / ((double)cs + 1)) * c.E_norm;  // POL:585
  const double cgf = hs * sqrt(c.c_Sigma * (2 - c.c_Sigma) * c.mu_eff);
  for (int i = threadIdx.x; i < cs; i += blockDim.x) pSig[i] = (1 - c.c_Sigma) * pSig[i] + cgf * dw[i];  // POL:586
  double ts = 0.0;  // POL:588-596
  for (int ii = threadIdx.x; ii < K; ii += blockDim.x) {
    const double d = dvec[ii] / sigma_ds;
    double w0 = ws[ii];
    if (!(w0 >= 0)) {
      const double nrm = fabs(d) * normC;  // norm(C * scalar) = |scalar| ‖C‖_F
      w0 = (double)n_iter * w0 / (nrm * nrm);  // `n` is the AIS iteration counter here (POL:593)
    }
    ts += w0 * d * d;
  }
  const double temp_sum = block_sum(ts, red);  // (barriers inside also publish pSig)
  // POL:598 on the upper triangle, POL:599 mirrors it
  for (int e = threadIdx.x; e < cs * cs; e += blockDim.x) {
    const int i = e / cs, j = e % cs;
    if (i > j) continue;
    const double sij = Sigma[(size_t)j * cs + i];
    const double v = (1 - c.c1 - c.c_mu) * sij +
                     c.c1 * (pSig[i] * pSig[j] + (1 - hs) * c.c_Sigma * (2 - c.c_Sigma) * sij) + c.c_mu * temp_sum;
    Sigma[(size_t)j * cs + i] = v;
    Sigma[(size_t)i * cs + j] = v;
  }
  if (threadIdx.x == 0) *sigma_dev = sigma_ds * exp(c.c_sigma / c.d_sigma * (nps / c.E_norm - 1));  // POL:582
}

void launch_cma_vec(const double *dw, const double *C, const double *dvec, const double *ws, int K, int cs,
                    int n_iter, const mpopis_cma_t &c, double *psig, double *pSig, double *sigma_dev, double *U,
                    double *Sigma, const int *stop, cudaStream_t s) {
  cma_vec_kernel<<<1, 1024, 0, s>>>(dw, C, dvec, ws, K, cs, n_iter, c, psig, pSig, sigma_dev, U, Sigma, stop);
}

}  // namespace mpopis
