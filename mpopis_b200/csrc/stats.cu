// stats.cu — G4/G5/G7/G8/G10: everything between two rollout batches of the AIS loop.
//
//   G7  compute_weights(Information_Theoretic, costs)                      UTL:79-86
//   G8  weighted noise  Σ_k w_k E[r,k]  + shift/clamp/roll                 POL:226-231, UTL:88-101
//   G4  elite gather + early-stop test                                     POL:455-461, 563-569
//   G5  mean / covariance of elites or weighted samples                    POL:464-465, 364, 662, 732, 807
//       + CovarianceEstimation shrinkage (:lw :ss :rblw :oas)              POL:414-426
//   G10 multinomial resampling (PMC) as per-sample counts                  POL:804-806
//
// All reductions are two-phase (per-chunk partials in a fixed grid, then an ordered sum), so
// results are deterministic run to run; in the sharded configuration the ordered sums are what
// gets all-reduced. Every kernel of the AIS loop takes the device-side `stop` flag (CE/CMA
// early stop) and returns immediately once it is set, so the host never has to synchronise
// inside a control step.
#include <math_constants.h>

#include <cstdlib>

#include "engine.cuh"

namespace mpopis {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions (result valid in every thread); red = shared scratch of >= 33 doubles
template <int OP>  // 0 sum, 1 min, 2 max
__device__ __forceinline__ double block_reduce(double v, double *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = OP == 0 ? warp_sum(v) : (OP == 1 ? warp_min(v) : warp_max(v));
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : (OP == 0 ? 0.0 : (OP == 1 ? CUDART_INF : -CUDART_INF));
    t = OP == 0 ? warp_sum(t) : (OP == 1 ? warp_min(t) : warp_max(t));
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---- G7: importance weights -------------------------------------------------------------------
// w_k = exp(-(c_k − ρ)/λ) / η, ρ = min c, η = Σ exp(...)  (UTL:79-86). Single CTA: K values are
// at most a few MB and this runs once per AIS iteration.
__global__ void __launch_bounds__(1024) weights_kernel(const double *__restrict__ costs, int K, double lambda,
                                                        double *__restrict__ w, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  double mn = CUDART_INF;
  for (int k = threadIdx.x; k < K; k += blockDim.x) mn = fmin(mn, costs[k]);
  const double rho = block_reduce<1>(mn, red);
  const double ninv = -1 / lambda;
  double s = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double e = exp(ninv * (costs[k] - rho));
    w[k] = e;
    s += e;
  }
  const double eta = block_reduce<0>(s, red);
  for (int k = threadIdx.x; k < K; k += blockDim.x) w[k] = w[k] / eta;
}

// Multi-CTA version for large K: (1) per-CTA minima, (2) exp + per-CTA sums, (3) normalise. The partials
// are combined in a fixed order by every CTA, so the result is deterministic and identical on all ranks.
constexpr int WG_MAX = 256;
__global__ void __launch_bounds__(256) weights_min_kernel(const double *__restrict__ costs, int K,
                                                           double *__restrict__ scratch, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  double mn = CUDART_INF;
  for (int k = blockIdx.x * 256 + threadIdx.x; k < K; k += gridDim.x * 256) mn = fmin(mn, costs[k]);
  mn = block_reduce<1>(mn, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = mn;
}
__global__ void __launch_bounds__(256) weights_exp_kernel(const double *__restrict__ costs, int K, double lambda,
                                                           double *__restrict__ w, double *__restrict__ scratch,
                                                           const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  double mn = threadIdx.x < gridDim.x ? scratch[threadIdx.x] : CUDART_INF;
  const double rho = block_reduce<1>(mn, red);
  const double ninv = -1 / lambda;
  double s = 0.0;
  for (int k = blockIdx.x * 256 + threadIdx.x; k < K; k += gridDim.x * 256) {
    const double e = exp(ninv * (costs[k] - rho));
    w[k] = e;
    s += e;
  }
  s = block_reduce<0>(s, red);
  if (threadIdx.x == 0) scratch[WG_MAX + blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) weights_norm_kernel(int K, double *__restrict__ w,
                                                            const double *__restrict__ scratch, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  const double part = threadIdx.x < gridDim.x ? scratch[WG_MAX + threadIdx.x] : 0.0;
  const double eta = block_reduce<0>(part, red);
  for (int k = blockIdx.x * 256 + threadIdx.x; k < K; k += gridDim.x * 256) w[k] = w[k] / eta;
}

// scratch: 2 * WG_MAX doubles
int launch_weights(const double *costs, int K, double lambda, double *w, double *scratch, const int *stop,
                   cudaStream_t s) {
  if (K <= 8192 || !scratch) {
    weights_kernel<<<1, 1024, 0, s>>>(costs, K, lambda, w, stop);
    return 1;
  }
  const int g = min(WG_MAX, (K + 1023) / 1024);
  weights_min_kernel<<<g, 256, 0, s>>>(costs, K, scratch, stop);
  weights_exp_kernel<<<g, 256, 0, s>>>(costs, K, lambda, w, scratch, stop);
  weights_norm_kernel<<<g, 256, 0, s>>>(K, w, scratch, stop);
  return 3;
}

// ---- G8 / G5: weighted row sums ---------------------------------------------------------------
// partial[c][r] = Σ_{k in chunk c} w_k X[r][k] for r < rows, and partial[c][rows] = Σ_{k in chunk c} w_k
// (w == nullptr: w_k = 1). X is [rows][ld] with the sample index contiguous, so every warp streams
// 256-byte row segments: this is the one genuinely HBM-bound kernel of the path (8·cs·K bytes).
constexpr int RS_THREADS = 256, RS_ROWS = 4, RS_CHUNK = 4096;
template <bool SQ>  // SQ: also the second raw moment (sharded shrinkage); the plain instance is the G8 kernel
__global__ void __launch_bounds__(RS_THREADS) rowsum_partial_kernel(const double *__restrict__ X, long long ld,
                                                                     int rows, int n,
                                                                     const double *__restrict__ w,
                                                                     double *__restrict__ partial,
                                                                     const int *stop, const int *n_dev) {
  constexpr bool sq = SQ;
  if (stop && *stop) return;
  if (n_dev) n = min(n, *n_dev);  // device-side column count (sharded elite sets); chunks beyond it write zeros
  __shared__ double red[33];
  const int c = blockIdx.x, r0 = blockIdx.y * RS_ROWS;
  const int kbeg = c * RS_CHUNK, kend = min(n, kbeg + RS_CHUNK);
  const int width = sq ? 2 * rows + 1 : rows + 1;  // sq: also Σ_k w_k X[r][k]² at column rows + 1 + r
  double acc[RS_ROWS], acc2[RS_ROWS];
#pragma unroll
  for (int q = 0; q < RS_ROWS; ++q) acc[q] = acc2[q] = 0.0;
  for (int k = kbeg + threadIdx.x; k < kend; k += RS_THREADS) {
    const double wk = w ? w[k] : 1.0;
#pragma unroll
    for (int q = 0; q < RS_ROWS; ++q) {
      const int r = r0 + q;
      if (r < rows) {
        const double x = X[(size_t)r * ld + k];
        acc[q] = fma(wk, x, acc[q]);
        if (sq) acc2[q] = fma(wk * x, x, acc2[q]);
      } else if (r == rows) acc[q] += wk;
    }
  }
#pragma unroll
  for (int q = 0; q < RS_ROWS; ++q) {
    const double t = block_reduce<0>(acc[q], red);
    if (threadIdx.x == 0 && r0 + q <= rows) partial[(size_t)c * width + r0 + q] = t;
    if (sq) {
      const double t2 = block_reduce<0>(acc2[q], red);
      if (threadIdx.x == 0 && r0 + q < rows) partial[(size_t)c * width + rows + 1 + r0 + q] = t2;
    }
  }
}

int rowsum_nchunks(int n) { return (n + RS_CHUNK - 1) / RS_CHUNK; }

void launch_rowsum_partial(const double *X, long long ld, int rows, int n, const double *w, double *partial,
                           const int *stop, cudaStream_t s, const int *n_dev, int sq) {
  dim3 grid(rowsum_nchunks(n), (rows + 1 + RS_ROWS - 1) / RS_ROWS);
  if (sq) rowsum_partial_kernel<true><<<grid, RS_THREADS, 0, s>>>(X, ld, rows, n, w, partial, stop, n_dev);
  else rowsum_partial_kernel<false><<<grid, RS_THREADS, 0, s>>>(X, ld, rows, n, w, partial, stop, n_dev);
}

// out[i] = Σ_c partial[c][i] in chunk order (deterministic)
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nchunks, int n, int stride,
                                       double *__restrict__ out, const int *stop) {
  if (stop && *stop) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += partial[(size_t)c * stride + i];
  out[i] = s;
}

// Few columns, many chunks (the shrinkage statistic: ONE column, K_loc/64 partials — 1024 at 65 536 samples per rank):
// the kernel above would walk them with one thread, a serial chain of L2 loads that outlasted the scatter matrix it
// runs beside (profiles/r2_multi_gpu.md). One CTA per column: 256 strided partial sums in chunk order, then a fixed
// shared-memory tree (deterministic; the same order on every rank).
__global__ void __launch_bounds__(256) reduce_partials_tall_kernel(const double *__restrict__ partial, int nchunks, int stride,
                                                                    double *__restrict__ out, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[256];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int c = threadIdx.x; c < nchunks; c += 256) s += partial[(size_t)c * stride + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[i] = red[0];
}

// stride = row pitch of `partial` (0: n)
void launch_reduce_partials(const double *partial, int nchunks, int n, double *out, const int *stop,
                            cudaStream_t s, int stride) {
  if (n <= 4 && nchunks > 64)
    reduce_partials_tall_kernel<<<n, 256, 0, s>>>(partial, nchunks, stride ? stride : n, out, stop);
  else
    reduce_partials_kernel<<<(n + 255) / 256, 256, 0, s>>>(partial, nchunks, n, stride ? stride : n, out, stop);
}

// μ = sums[0:rows] / sums[rows]; optionally U += scale * μ  (pol.U = pol.U + vec(μ′), POL:365,465,...)
__global__ void finalize_mean_kernel(const double *__restrict__ sums, int rows, double *__restrict__ mu,
                                     double *__restrict__ U, const double *scale_dev, const int *stop) {
  if (stop && *stop) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const double m = sums[r] / sums[rows];
  if (mu) mu[r] = m;
  if (U) U[r] = U[r] + (scale_dev ? *scale_dev : 1.0) * m;
}

void launch_finalize_mean(const double *sums, int rows, double *mu, double *U, const double *scale_dev,
                          const int *stop, cudaStream_t s) {
  finalize_mean_kernel<<<(rows + 127) / 128, 128, 0, s>>>(sums, rows, mu, U, scale_dev, stop);
}

// ---- G5: centred (weighted) scatter matrix ------------------------------------------------------
// P[c][i][j] = Σ_{k in chunk c} w_k (X[i][k] − μ_i)(X[j][k] − μ_j) for the lower 32x32 tiles.
// 16x16 threads, 2x2 outputs per thread, 32-sample slabs staged (centred) in shared memory.
// 64 x 64 output tile per CTA (lower tiles only), 16 x 16 threads with a 4 x 4 register tile each,
// 16-sample slabs staged sample-major in shared memory: per sample a thread issues four 16-byte
// shared loads (two of them warp-broadcast) for 16 DFMAs.
constexpr int SY_T = 64, SY_K = 16;
__global__ void __launch_bounds__(256) syrk_partial_kernel(const double *__restrict__ X, long long ld, int p,
                                                            int n, const double *__restrict__ w,
                                                            const double *__restrict__ mu, int chunk,
                                                            double *__restrict__ P, const int *stop,
                                                            const int *n_dev) {
  if (stop && *stop) return;
  if (n_dev) {  // device-side column count: re-balance the chunks over the grid (all CTAs stay busy)
    n = min(n, *n_dev);
    chunk = max(SY_K, (((n + (int)gridDim.y - 1) / (int)gridDim.y + SY_K - 1) / SY_K) * SY_K);
  }
  // [sample][row], row pitch padded by 2 doubles: the transposing stores are 2-way instead of 16-way
  // bank-conflicted while rows stay 16-byte aligned for the vector reads
  __shared__ __align__(16) double As[SY_K][SY_T + 2], Bs[SY_K][SY_T + 2];
  int t = blockIdx.x, bi = 0;  // decode the lower-triangular tile index
  while (t > bi) t -= bi + 1, ++bi;
  const int bj = t;
  const int c = blockIdx.y;
  const int kbeg = c * chunk, kend = min(n, kbeg + chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  // each thread stages 4 + 4 elements per slab: rows rr0 + 16 q, sample kk0 (16 consecutive samples of a
  // row = one 128-byte segment); the NEXT slab is fetched into registers while the current one is consumed
  const int kk0 = threadIdx.x & 15, rr0 = threadIdx.x >> 4;
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
    const int k = k0 + kk0;
    const bool ok = k < kend;
    const double wk = ok ? (w ? w[k] : 1.0) : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ia = bi * SY_T + rr0 + 16 * q, ib = bj * SY_T + rr0 + 16 * q;
      ra[q] = (ok && ia < p) ? wk * (X[(size_t)ia * ld + k] - mu[ia]) : 0.0;
      rb[q] = (ok && ib < p) ? (X[(size_t)ib * ld + k] - mu[ib]) : 0.0;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += SY_K) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) As[kk0][rr0 + 16 * q] = ra[q], Bs[kk0][rr0 + 16 * q] = rb[q];
    __syncthreads();
    if (k0 + SY_K < kend) fetch(k0 + SY_K);
#pragma unroll
    for (int kk = 0; kk < SY_K; ++kk) {
      const double2 x01 = *reinterpret_cast<const double2 *>(&As[kk][ty * 4]);
      const double2 x23 = *reinterpret_cast<const double2 *>(&As[kk][ty * 4 + 2]);
      const double2 y01 = *reinterpret_cast<const double2 *>(&Bs[kk][tx * 2]);       // columns 2tx, 2tx+1
      const double2 y23 = *reinterpret_cast<const double2 *>(&Bs[kk][32 + tx * 2]);  // columns 32+2tx, +1
      const double x[4] = {x01.x, x01.y, x23.x, x23.y}, y[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(x[a], y[b], acc[a][b]);
    }
  }
  double *Pc = P + (size_t)c * p * p;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = bi * SY_T + ty * 4 + a;
    if (i >= p) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = bj * SY_T + (b >> 1) * 32 + tx * 2 + (b & 1);
      if (j < p) Pc[(size_t)i * p + j] = acc[a][b];
    }
  }
}

int syrk_chunk(int n) {
  // aim for <= 96 chunks (≈ two CTAs per SM for cs = 100: 3 tiles x 96; the kernel is latency-bound with
  // one), multiples of 16 samples
  int chunk = ((n + 95) / 96 + SY_K - 1) / SY_K * SY_K;
  return chunk < SY_K ? SY_K : chunk;
}
int syrk_nchunks(int n) {
  const int ch = syrk_chunk(n);
  return (n + ch - 1) / ch;
}

// The same scatter tile on the FP64 tensor cores (mma.sync.m8n8k4.f64 -> SASS DMMA): the Σ outer-product
// update is the contraction BASELINE.json's north star reserves the tensor cores for. 4 warps per CTA, each
// owning a 32 x 32 quadrant of the 64 x 64 tile as 4 x 4 accumulator fragments; slabs of 16 samples are
// staged row-major with a pitch of 20 doubles, so the A fragment As[g][t] and the B fragment Bs[g][t]
// (B is "column-major k x j", i.e. the row-major centred data itself) load conflict-free.
constexpr int SD_P = 20;
__global__ void __launch_bounds__(128) syrk_dmma_kernel(const double *__restrict__ X, long long ld, int p, int n,
                                                         const double *__restrict__ w,
                                                         const double *__restrict__ mu, int chunk,
                                                         double *__restrict__ P, const int *stop,
                                                         const int *n_dev) {
  if (stop && *stop) return;
  if (n_dev) {  // device-side column count: re-balance the chunks over the grid (all CTAs stay busy)
    n = min(n, *n_dev);
    chunk = max(SY_K, (((n + (int)gridDim.y - 1) / (int)gridDim.y + SY_K - 1) / SY_K) * SY_K);
  }
  __shared__ double As[SY_T][SD_P], Bs[SY_T][SD_P];  // [row][sample]
  int tt = blockIdx.x, bi = 0;  // decode the lower-triangular tile index
  while (tt > bi) tt -= bi + 1, ++bi;
  const int bj = tt;
  const int c = blockIdx.y;
  const int kbeg = c * chunk, kend = min(n, kbeg + chunk);
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wi = (wq >> 1) * 32, wj = (wq & 1) * 32;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // staging: thread -> sample kk0 = tid % 16 and rows rr0 + 8 q (q < 8); next slab prefetched into registers
  const int kk0 = threadIdx.x & 15, rr0 = threadIdx.x >> 4;
  double ra[8], rb[8];
  auto fetch = [&](int k0) {
    const int k = k0 + kk0;
    const bool ok = k < kend;
    const double wk = ok ? (w ? w[k] : 1.0) : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ia = bi * SY_T + rr0 + 8 * q, ib = bj * SY_T + rr0 + 8 * q;
      ra[q] = (ok && ia < p) ? wk * (X[(size_t)ia * ld + k] - mu[ia]) : 0.0;
      rb[q] = (ok && ib < p) ? (X[(size_t)ib * ld + k] - mu[ib]) : 0.0;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += SY_K) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) As[rr0 + 8 * q][kk0] = ra[q], Bs[rr0 + 8 * q][kk0] = rb[q];
    __syncthreads();
    if (k0 + SY_K < kend) fetch(k0 + SY_K);
#pragma unroll
    for (int k4 = 0; k4 < SY_K; k4 += 4) {
      double bf[4];
#pragma unroll
      for (int cf = 0; cf < 4; ++cf) bf[cf] = Bs[wj + cf * 8 + g][k4 + t];
#pragma unroll
      for (int rf = 0; rf < 4; ++rf) {
        const double af = As[wi + rf * 8 + g][k4 + t];
#pragma unroll
        for (int cf = 0; cf < 4; ++cf)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[rf][cf][0]), "+d"(acc[rf][cf][1])
                       : "d"(af), "d"(bf[cf]));
      }
    }
  }
  double *Pc = P + (size_t)c * p * p;
#pragma unroll
  for (int rf = 0; rf < 4; ++rf) {
    const int i = bi * SY_T + wi + rf * 8 + g;
    if (i >= p) continue;
#pragma unroll
    for (int cf = 0; cf < 4; ++cf) {
      const int j = bj * SY_T + wj + cf * 8 + 2 * t;
      if (j < p) Pc[(size_t)i * p + j] = acc[rf][cf][0];
      if (j + 1 < p) Pc[(size_t)i * p + j + 1] = acc[rf][cf][1];
    }
  }
}

static int syrk_path() {  // MPOPIS_SYRK=fma selects the DFMA kernel (A/B evidence, profiles/)
  static int path = -1;
  if (path < 0) {
    const char *e = getenv("MPOPIS_SYRK");
    path = (e && e[0] == 'f') ? 0 : 1;
  }
  return path;
}

void launch_syrk_partial(const double *X, long long ld, int p, int n, const double *w, const double *mu,
                         double *P, const int *stop, cudaStream_t s, const int *n_dev) {
  const int nt = (p + SY_T - 1) / SY_T;
  dim3 grid(nt * (nt + 1) / 2, syrk_nchunks(n));
  if (syrk_path() == 1) syrk_dmma_kernel<<<grid, 128, 0, s>>>(X, ld, p, n, w, mu, syrk_chunk(n), P, stop, n_dev);
  else syrk_partial_kernel<<<grid, 256, 0, s>>>(X, ld, p, n, w, mu, syrk_chunk(n), P, stop, n_dev);
}

// S[i][j] (lower, i >= j) = Σ_c P[c][i][j] in chunk order; raw scatter sums, mirrored to full storage.
// 32 elements x 8 chunk-groups per CTA: group g sums the chunks c ≡ g (mod 8) in order, then the 8 group
// sums are combined in order — a fixed summation tree (deterministic), 8x the loads in flight of a plain
// per-element loop (that version was latency-bound: 41 µs for 96 chunks of a 100 x 100 matrix).
__global__ void __launch_bounds__(256) scatter_reduce_kernel(const double *__restrict__ P, int nchunks, int p,
                                                              double *__restrict__ S, const int *stop) {
  if (stop && *stop) return;
  __shared__ double part[8][33];
  const int le = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + le;
  const bool ok = e < p * p && (e % p) <= (e / p);
  double s = 0.0;
  if (ok)
    for (int c = g; c < nchunks; c += 8) s += P[(size_t)c * p * p + e];
  part[g][le] = s;
  __syncthreads();
  if (g == 0 && ok) {
    double t = part[0][le];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += part[q][le];
    const int i = e / p, j = e % p;
    S[(size_t)i * p + j] = t;
    S[(size_t)j * p + i] = t;
  }
}

void launch_scatter_reduce(const double *P, int nchunks, int p, double *S, const int *stop, cudaStream_t s) {
  scatter_reduce_kernel<<<(p * p + 31) / 32, 256, 0, s>>>(P, nchunks, p, S, stop);
}

// Σ_{i≠j} Σ_k (z_ki z_kj)² = Σ_k [(Σ_i z_ki²)² − Σ_i z_ki⁴], z = (x − μ)·d  (d_i = 1/sqrt(S_ii/n) for :ss,
// 1 for :lw). O(n p) instead of a second p x p x n contraction. Sraw = raw scatter sums (S = Sraw/n).
__global__ void __launch_bounds__(256) shrink_q_partial_kernel(const double *__restrict__ X, long long ld,
                                                                int p, int n, const double *__restrict__ w,
                                                                const double *__restrict__ mu,
                                                                const double *__restrict__ Sraw,
                                                                const double *cnt_dev, int standardise,
                                                                double *__restrict__ partial, const int *stop,
                                                                const int *n_dev,
                                                                const double *__restrict__ dinv_ext) {
  if (stop && *stop) return;
  if (n_dev) n = min(n, *n_dev);
  __shared__ double red[33];
  extern __shared__ double dsc[];  // p per-variable scales 1/σ_i (1 for :lw), p means
  {
    const double cnt = *cnt_dev;  // global number of observations (Σ of ownership weights)
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      dsc[i] = dinv_ext ? dinv_ext[i] : (standardise ? 1.0 / sqrt(Sraw[(size_t)i * p + i] / cnt) : 1.0);
      dsc[p + i] = mu[i];
    }
  }
  __syncthreads();
  // 64 samples x 4 row groups per CTA: the rows of a sample are split over 4 threads and recombined
  __shared__ double sa[4][64], sb[4][64];
  const int ks = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + ks;
  double a = 0.0, b = 0.0;
  if (k < n && (!w || w[k] != 0.0)) {
#pragma unroll 4
    for (int i = g; i < p; i += 4) {
      const double z = (X[(size_t)i * ld + k] - dsc[p + i]) * dsc[i];
      const double z2 = z * z;
      a += z2;
      b = fma(z2, z2, b);
    }
  }
  sa[g][ks] = a, sb[g][ks] = b;
  __syncthreads();
  double q = 0.0;
  if (g == 0) {
    const double at = (sa[0][ks] + sa[1][ks]) + (sa[2][ks] + sa[3][ks]);
    const double bt = (sb[0][ks] + sb[1][ks]) + (sb[2][ks] + sb[3][ks]);
    q = at * at - bt;
  }
  const double t = block_reduce<0>(q, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

int shrink_q_nblocks(int n) { return (n + 63) / 64; }

void launch_shrink_q_partial(const double *X, long long ld, int p, int n, const double *w, const double *mu,
                             const double *Sraw, const double *cnt_dev, int standardise, double *partial,
                             const int *stop, cudaStream_t s, const int *n_dev, const double *dinv_ext) {
  shrink_q_partial_kernel<<<shrink_q_nblocks(n), 256, sizeof(double) * 2 * p, s>>>(
      X, ld, p, n, w, mu, Sraw, cnt_dev, standardise, partial, stop, n_dev, dinv_ext);
}

// Final covariance: Σ′ = shrink(method, Sraw / denom) + ridge·I, written to Sigma (symmetric, so
// row/column-major coincide). denom = count − corrected where count = *cnt_dev (Σw or n).
//   :mle          SimpleCovariance()                                           S
//   :lw / :ss     (1−λ)S + λ diag(S), λ = Σ_{i≠j}Var^(s_ij)/Σ_{i≠j}s_ij² (on correlations for :ss)
//   :rblw / :oas  (1−λ)S + λ tr(S)/p I, Chen et al. (2010) eq. 17/19 and 23
// (restated from the published formulas — unpinned, like the oracle.)
__global__ void __launch_bounds__(1024) cov_finalize_kernel(const double *__restrict__ Sraw, int p,
                                                             const double *cnt_dev, int corrected, int method,
                                                             const double *__restrict__ qpart, int nq,
                                                             double ridge, double *__restrict__ Sigma,
                                                             double *lambda_out, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[33];
  const double cnt = *cnt_dev, denom = cnt - (corrected ? 1.0 : 0.0), inv = 1.0 / denom;
  double lam = 0.0;
  if (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS) {
    const bool ss = method == MPOPIS_SIGMA_SS;
    double q = 0.0, r2 = 0.0;
    for (int c = threadIdx.x; c < nq; c += blockDim.x) q += qpart[c];
    q = block_reduce<0>(q, red);
    extern __shared__ double dinv[];  // p entries: 1/σ_i (:ss) or 1 (:lw)
    for (int i = threadIdx.x; i < p; i += blockDim.x) dinv[i] = ss ? 1.0 / sqrt(Sraw[(size_t)i * p + i] * inv) : 1.0;
    __syncthreads();
    for (int i = threadIdx.x >> 5; i < p; i += blockDim.x >> 5) {  // one warp per row: no integer divisions
      const double di = dinv[i] * inv;
      for (int j = threadIdx.x & 31; j < p; j += 32) {
        if (i == j) continue;
        const double v = Sraw[(size_t)i * p + j] * di * dinv[j];
        r2 = fma(v, v, r2);
      }
    }
    r2 = block_reduce<0>(r2, red);
    const double n = cnt;
    const double num = (q - n * r2) * n / ((n - 1.0) * n * n);
    lam = fmin(fmax(num / r2, 0.0), 1.0);
  } else if (method == MPOPIS_SIGMA_RBLW || method == MPOPIS_SIGMA_OAS) {
    double tr = 0.0, tr2 = 0.0;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
      const double v = Sraw[e] * inv;
      tr2 = fma(v, v, tr2);
      if (e / p == e % p) tr += v;
    }
    tr = block_reduce<0>(tr, red);
    tr2 = block_reduce<0>(tr2, red);
    const double n = cnt, pd = (double)p, trsq = tr * tr;
    if (method == MPOPIS_SIGMA_RBLW) lam = ((n - 2) / n * tr2 + trsq) / ((n + 2) * (tr2 - trsq / pd));
    else lam = ((1.0 - 2.0 / pd) * tr2 + trsq) / ((n + 1.0 - 2.0 / pd) * (tr2 - trsq / pd));
    lam = fmin(fmax(lam, 0.0), 1.0);
    const double F = tr / pd;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
      const bool dg = e / p == e % p;
      Sigma[e] = (1.0 - lam) * (Sraw[e] * inv) + (dg ? lam * F : 0.0) + (dg ? ridge : 0.0);
    }
    if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
    return;
  }
  for (int i = threadIdx.x >> 5; i < p; i += blockDim.x >> 5)
    for (int j = threadIdx.x & 31; j < p; j += 32) {
      const double v = Sraw[(size_t)i * p + j] * inv;
      Sigma[(size_t)i * p + j] = i == j ? v + ridge : (1.0 - lam) * v;
    }
  if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
}

void launch_cov_finalize(const double *Sraw, int p, const double *cnt_dev, int corrected, int method,
                         const double *qpart, int nq, double ridge, double *Sigma, double *lambda_out,
                         const int *stop, cudaStream_t s) {
  cov_finalize_kernel<<<1, 1024, sizeof(double) * p, s>>>(Sraw, p, cnt_dev, corrected, method, qpart, nq, ridge,
                                                          Sigma, lambda_out, stop);
}

// ---- G5, small-n path: the whole moment chain in ONE CTA -----------------------------------------------------
// For n <= 512 columns (the reference's own sizes: K = 150 -> 30 elites, K = 20 for MountainCar) the chain above is
// eight launches of a few µs of latency each — more than the work. This kernel does count, (weighted) mean, U += μ,
// centred scatter matrix, the shrinkage statistic and the final Σ′ = shrink(S) + ridge·I with the same formulas
// (rowsum_partial → finalize_mean → syrk_partial → shrink_q_partial → cov_finalize), single GPU only. `cols` (nullable)
// selects the columns — the elite set order[0:m] of :cemppi, which saves the gather kernel as well. The centred
// columns are staged in shared memory in chunks of kc (Xs[p][pitch], odd pitch: conflict-free for lanes over rows);
// a first cut that read the (gathered) columns straight from global memory inside the p²/2 dot products was a serial
// chain of dependent L2 loads: 119 µs for p = 100, n = 30 (profiles/README.md). Sraw is a p x p global scratch
// (L1/L2-resident) accumulated chunk by chunk and re-read by the same CTA after a barrier.
__global__ void __launch_bounds__(1024) moments_small_kernel(const double *__restrict__ X, long long ld, int p, int n,
                                                              const double *__restrict__ w, const int *__restrict__ cols,
                                                              int want_cov, int corrected, int method, double ridge,
                                                              int kc, int pitch, double *__restrict__ mu_out,
                                                              double *__restrict__ U, const double *scale_dev,
                                                              double *__restrict__ sums_out, double *__restrict__ Sraw,
                                                              double *__restrict__ Sigma, double *lambda_out,
                                                              const int *stop) {
  if (stop && *stop) return;
  extern __shared__ double msm[];  // mu[p] | dinv[p] | red[40] | sw[kc] | Xs[p][pitch]
  double *mu = msm, *dinv = msm + p, *red = msm + 2 * p, *sw = msm + 2 * p + 40, *Xs = sw + kc;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double c = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) c += w ? w[k] : 1.0;
  const double cnt = block_reduce<0>(c, red);
  for (int r = wid; r < p; r += nw) {  // one warp per row, lanes over the columns
    double a = 0.0;
    for (int k = lane; k < n; k += 32) a = fma(w ? w[k] : 1.0, X[(size_t)r * ld + (cols ? cols[k] : k)], a);
    a = warp_sum(a);
    if (lane == 0) {
      const double m = a / cnt;
      mu[r] = m;
      if (mu_out) mu_out[r] = m;
      if (sums_out) sums_out[r] = a;
      if (U) U[r] = U[r] + (scale_dev ? *scale_dev : 1.0) * m;  // pol.U = pol.U + vec(μ′), POL:365,465,...
    }
  }
  if (threadIdx.x == 0 && sums_out) sums_out[p] = cnt;
  __syncthreads();
  if (!want_cov) return;
  auto stage = [&](int c0, int nc) {  // Xs[i][k] = X[i][col(c0 + k)] − μ_i, sw[k] = w_k (all loads independent)
    for (int e = threadIdx.x; e < p * nc; e += blockDim.x) {
      const int i = e / nc, k = e - i * nc;
      const int ck = cols ? cols[c0 + k] : c0 + k;
      Xs[i * pitch + k] = X[(size_t)i * ld + ck] - mu[i];
    }
    for (int k = threadIdx.x; k < nc; k += blockDim.x) sw[k] = w ? w[c0 + k] : 1.0;
  };
  for (int c0 = 0; c0 < n; c0 += kc) {  // centred scatter matrix, lower triangle: one warp per row i, lanes over j <= i
    const int nc = min(kc, n - c0);
    __syncthreads();
    stage(c0, nc);
    __syncthreads();
    for (int i = wid; i < p; i += nw) {
      const double *xi = Xs + i * pitch;
      for (int j = lane; j <= i; j += 32) {
        const double *xj = Xs + j * pitch;
        double a = 0.0;
        for (int k = 0; k < nc; ++k) a = fma(sw[k] * xi[k], xj[k], a);
        const double v = (c0 ? Sraw[(size_t)i * p + j] : 0.0) + a;
        Sraw[(size_t)i * p + j] = v;
        Sraw[(size_t)j * p + i] = v;
      }
    }
  }
  __syncthreads();
  const double denom = cnt - (corrected ? 1.0 : 0.0), inv = 1.0 / denom;
  double lam = 0.0;
  if (method == MPOPIS_SIGMA_LW || method == MPOPIS_SIGMA_SS) {
    const bool ss = method == MPOPIS_SIGMA_SS;
    // Σ_{i≠j} Σ_k (z_ki z_kj)² = Σ_k [(Σ_i z_ki²)² − Σ_i z_ki⁴]  (shrink_q_partial_kernel); one warp per column
    for (int i = threadIdx.x; i < p; i += blockDim.x) dinv[i] = ss ? 1.0 / sqrt(Sraw[(size_t)i * p + i] / cnt) : 1.0;
    __syncthreads();
    double q = 0.0;
    for (int c0 = 0; c0 < n; c0 += kc) {
      const int nc = min(kc, n - c0);
      if (n > kc) {  // a single chunk is still resident
        __syncthreads();
        stage(c0, nc);
        __syncthreads();
      }
      for (int k = wid; k < nc; k += nw) {
        if (sw[k] == 0.0) continue;  // warp-uniform
        double a = 0.0, b = 0.0;
        for (int i = lane; i < p; i += 32) {
          const double z = Xs[i * pitch + k] * dinv[i];
          const double z2 = z * z;
          a += z2;
          b = fma(z2, z2, b);
        }
        a = warp_sum(a), b = warp_sum(b);
        if (lane == 0) q += a * a - b;
      }
    }
    q = block_reduce<0>(q, red);
    // cov_finalize_kernel, :lw / :ss
    for (int i = threadIdx.x; i < p; i += blockDim.x) dinv[i] = ss ? 1.0 / sqrt(Sraw[(size_t)i * p + i] * inv) : 1.0;
    __syncthreads();
    double r2 = 0.0;
    for (int i = wid; i < p; i += nw) {
      const double di = dinv[i] * inv;
      for (int j = lane; j < p; j += 32) {
        if (i == j) continue;
        const double v = Sraw[(size_t)i * p + j] * di * dinv[j];
        r2 = fma(v, v, r2);
      }
    }
    r2 = block_reduce<0>(r2, red);
    const double nn = cnt;
    const double num = (q - nn * r2) * nn / ((nn - 1.0) * nn * nn);
    lam = fmin(fmax(num / r2, 0.0), 1.0);
  } else if (method == MPOPIS_SIGMA_RBLW || method == MPOPIS_SIGMA_OAS) {
    double tr = 0.0, tr2 = 0.0;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
      const double v = Sraw[e] * inv;
      tr2 = fma(v, v, tr2);
      if (e / p == e % p) tr += v;
    }
    tr = block_reduce<0>(tr, red);
    tr2 = block_reduce<0>(tr2, red);
    const double nn = cnt, pd = (double)p, trsq = tr * tr;
    if (method == MPOPIS_SIGMA_RBLW) lam = ((nn - 2) / nn * tr2 + trsq) / ((nn + 2) * (tr2 - trsq / pd));
    else lam = ((1.0 - 2.0 / pd) * tr2 + trsq) / ((nn + 1.0 - 2.0 / pd) * (tr2 - trsq / pd));
    lam = fmin(fmax(lam, 0.0), 1.0);
    const double F = tr / pd;
    for (int e = threadIdx.x; e < p * p; e += blockDim.x) {
      const bool dg = e / p == e % p;
      Sigma[e] = (1.0 - lam) * (Sraw[e] * inv) + (dg ? lam * F : 0.0) + (dg ? ridge : 0.0);
    }
    if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
    return;
  }
  for (int i = wid; i < p; i += nw)
    for (int j = lane; j < p; j += 32) {
      const double v = Sraw[(size_t)i * p + j] * inv;
      Sigma[(size_t)i * p + j] = i == j ? v + ridge : (1.0 - lam) * v;
    }
  if (threadIdx.x == 0 && lambda_out) *lambda_out = lam;
}

void launch_moments_small(const double *X, long long ld, int p, int n, const double *w, const int *cols, int want_cov,
                          int corrected, int method, double ridge, double *mu_out, double *U, const double *scale_dev,
                          double *sums_out, double *Sraw, double *Sigma, double *lambda_out, const int *stop,
                          cudaStream_t s) {
  // chunk of columns staged in shared memory: 2p + 40 + kc + p·pitch doubles <= 96 KB with pitch <= kc + 1
  int kc = (int)((96LL * 1024 / 8 - 3LL * p - 40) / (p + 1));
  kc = kc < 1 ? 1 : (kc > n ? n : kc);
  const int pitch = kc | 1;  // odd: lanes over rows hit distinct banks
  const size_t smem = sizeof(double) * (2 * (size_t)p + 40 + kc + (size_t)p * pitch);
  // function attributes are per device: set it on every launch (a host-side table lookup, like rollout_car does)
  cudaFuncSetAttribute(moments_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  moments_small_kernel<<<1, 1024, smem, s>>>(X, ld, p, n, w, cols, want_cov, corrected, method, ridge, kc, pitch, mu_out,
                                             U, scale_dev, sums_out, Sraw, Sigma, lambda_out, stop);
}

// ---- G4: elite gather + early-stop test ---------------------------------------------------------
// X[r][j] = E[r][order[j] − k0] for the elites that live in this shard (k0 <= order[j] < k0 + Kloc);
// elites owned by other shards contribute zeros (they are summed in by the all-reduce).
__global__ void gather_cols_kernel(const double *__restrict__ E, long long ldk, int cs, const int *__restrict__ order,
                                   int m, long long k0, int Kloc, double *__restrict__ X, long long ldx,
                                   double *__restrict__ mask, const int *stop, const int *m_dev) {
  if (stop && *stop) return;
  if (m_dev) m = min(m, *m_dev);
  const int j = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (j >= m) return;
  const long long k = (long long)order[j] - k0;
  const bool own = k >= 0 && k < Kloc;
  X[(size_t)r * ldx + j] = own ? E[(size_t)r * ldk + k] : 0.0;
  if (mask && r == 0) mask[j] = own ? 1.0 : 0.0;  // ownership weights for the sharded moments
}

void launch_gather_cols(const double *E, long long ldk, int cs, const int *order, int m, long long k0, int Kloc,
                        double *X, long long ldx, double *mask, const int *stop, cudaStream_t s,
                        const int *m_dev) {
  dim3 grid((m + 255) / 256, cs);
  gather_cols_kernel<<<grid, 256, 0, s>>>(E, ldk, cs, order, m, k0, Kloc, X, ldx, mask, stop, m_dev);
}

// maximum(abs.(diff(elite_traj_cost))) < 10e-3 -> break (POL:458-461, 566-569). Sets *stop.
__global__ void __launch_bounds__(1024) elite_stop_kernel(const double *__restrict__ sorted_costs, int m,
                                                           int enabled, int *stop) {
  if (*stop) return;
  __shared__ double red[33];
  double mx = -CUDART_INF;
  for (int j = threadIdx.x; j + 1 < m; j += blockDim.x) mx = fmax(mx, fabs(sorted_costs[j + 1] - sorted_costs[j]));
  mx = block_reduce<2>(mx, red);
  if (threadIdx.x == 0 && enabled && mx < 10e-3) *stop = 1;
}

void launch_elite_stop(const double *sorted_costs, int m, int enabled, int *stop, cudaStream_t s) {
  elite_stop_kernel<<<1, 1024, 0, s>>>(sorted_costs, m, enabled, stop);
}

__global__ void iter_begin_kernel(const int *stop, int *its, int *total_its) {
  if (!*stop) *its += 1, *total_its += 1;
}
void launch_iter_begin(const int *stop, int *its, int *total_its, cudaStream_t s) {
  iter_begin_kernel<<<1, 1, 0, s>>>(stop, its, total_its);
}

// ---- G10: PMC multinomial resampling -------------------------------------------------------------
// Categorical(ws) draws by inverse CDF (POL:804-805); E′ = E[:, idxs] (POL:806) enters the moments
// only through how often each column was drawn, so the gather is replaced by integer counts.
__global__ void __launch_bounds__(1024) inclusive_scan_kernel(const double *__restrict__ w, int K,
                                                               double *__restrict__ cdf, const int *stop) {
  if (stop && *stop) return;
  // sequential-in-chunks scan: each thread owns a contiguous segment (deterministic order)
  __shared__ double seg[1024];
  const int per = (K + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(K, beg + per);
  double s = 0.0;
  for (int k = beg; k < end; ++k) s += w[k];
  seg[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (int t = 0; t < (int)blockDim.x; ++t) {
      const double v = seg[t];
      seg[t] = run;
      run += v;
    }
  }
  __syncthreads();
  double run = seg[threadIdx.x];
  for (int k = beg; k < end; ++k) {
    run += w[k];
    cdf[k] = run;
  }
}

__global__ void resample_count_kernel(const double *__restrict__ cdf, int K, const double *__restrict__ u,
                                      int ndraws, int *__restrict__ counts, const int *stop) {
  if (stop && *stop) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndraws) return;
  const double ui = u[i];
  int lo = 0, hi = K - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ui < cdf[mid]) hi = mid;
    else lo = mid + 1;
  }
  atomicAdd(counts + lo, 1);
}

__global__ void counts_to_weights_kernel(const int *__restrict__ counts, long long k0, int Kloc,
                                         double *__restrict__ w, const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < Kloc) w[k] = (double)counts[k0 + k];
}

void launch_pmc_counts(const double *wglobal, int K, const double *u, double *cdf, int *counts, long long k0,
                       int Kloc, double *wloc, const int *stop, cudaStream_t s) {
  cudaMemsetAsync(counts, 0, sizeof(int) * K, s);
  inclusive_scan_kernel<<<1, 1024, 0, s>>>(wglobal, K, cdf, stop);
  resample_count_kernel<<<(K + 255) / 256, 256, 0, s>>>(cdf, K, u, K, counts, stop);
  counts_to_weights_kernel<<<(Kloc + 255) / 256, 256, 0, s>>>(counts, k0, Kloc, wloc, stop);
}

// ---- control-cost vector ---------------------------------------------------------------------------
// b = (γ U_orig') Σ_inv for a caller-supplied Σ_inv (depth-(i) ABI): b[j] = Σ_i γ U_orig[i] Σ_inv[i][j]
__global__ void ctrl_vec_kernel(const double *__restrict__ Sinv, int cs, const double *__restrict__ U_orig,
                                double gamma, double *__restrict__ b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cs) return;
  double s = 0.0;
  for (int i = 0; i < cs; ++i) s += (gamma * U_orig[i]) * Sinv[(size_t)j * cs + i];  // col-major, column j
  b[j] = s;
}
void launch_ctrl_vec(const double *Sinv, int cs, const double *U_orig, double gamma, double *b, cudaStream_t s) {
  ctrl_vec_kernel<<<(cs + 127) / 128, 128, 0, s>>>(Sinv, cs, U_orig, gamma, b);
}

// ---- final control: weighted_controls, clamp, roll (POL:226-231, UTL:88-101) -------------------------
// wsum[r] = Σ_k w_k E[r][k] (un-shifted E), wsum[cs] = Σ_k w_k. The reference shifts E by
// (pol.U − U_orig) before weighting (POL:468): Σ_k w_k (E[r,k] + Δ_r) = wsum[r] + Δ_r Σ_k w_k.
// bounds (nullable): [lo(as) | hi(as)] of action_space(pol.env) — the external (EnvpoolEnv-style) env's own limits;
// the built-in envs are ±1 per component (CAR:156-159, MCR:75-84, continuous MountainCar).
__global__ void finalize_control_kernel(const double *__restrict__ wsum, const double *__restrict__ U_orig,
                                        const double *__restrict__ U_cur, int cs, int as, int T,
                                        double *__restrict__ U_next, double *__restrict__ control,
                                        const double *__restrict__ bounds, unsigned *step_dev) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0 && step_dev) *step_dev += 1;  // every Philox draw of this control step has been made
  if (r >= cs) return;
  const double wc = U_orig[r] + (wsum[r] + (U_cur[r] - U_orig[r]) * wsum[cs]);
  if (r < as) control[r] = fmin(fmax(wc, bounds ? bounds[r] : -1.0), bounds ? bounds[as + r] : 1.0);  // UTL:91
  if (T > 1) {
    if (r >= as) U_next[r - as] = wc;                   // UTL:95
    if (r >= cs - as) U_next[r] = U_orig[r];            // UTL:96 is a no-op (App. B-2): tail keeps its values
  } else {
    U_next[r] = wc;  // UTL:98
  }
}
void launch_finalize_control(const double *wsum, const double *U_orig, const double *U_cur, int cs, int as, int T,
                             double *U_next, double *control, const double *bounds, unsigned *step_dev,
                             cudaStream_t s) {
  finalize_control_kernel<<<(cs + 127) / 128, 128, 0, s>>>(wsum, U_orig, U_cur, cs, as, T, U_next, control, bounds,
                                                          step_dev);
}

__global__ void set_scalar_kernel(double *dst, double value) { *dst = value; }
void launch_set_scalar(double *dst, double value, cudaStream_t s) { set_scalar_kernel<<<1, 1, 0, s>>>(dst, value); }

}  // namespace mpopis
