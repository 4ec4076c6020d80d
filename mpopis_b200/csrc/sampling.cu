// sampling.cu — G1: Gaussian control-perturbation draw.
//
// Replaces `E = rand(pol.rng, MvNormal(Σ′), K)` (POL:308,359,448,556,657,724,797; :mppi POL:193):
// Z ~ N(0, I) (cs x K), E = L Z with L the lower Cholesky factor of Σ′.
// Julia's MersenneTwister stream cannot be reproduced; the engine's generator is counter-based
// Philox4x32-10 keyed by the policy seed, counter = (global sample k, pair j, AIS iteration |
// purpose<<24, control-step counter), Box–Muller in FP64. Because the counter is the GLOBAL
// sample index, the draws are independent of how K is sharded across GPUs. oracle/ restates the
// same generator (tests compare them to ~1e-15).
#include "engine.cuh"

namespace mpopis {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

__device__ __forceinline__ void philox_u2(uint64_t seed, uint32_t k, uint32_t j, uint32_t it, uint32_t step,
                                          double *u1, double *u2) {
  uint32_t o[4];
  philox4x32_10(k, j, it, step, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
  *u1 = ((double)(a >> 12) + 0.5) * 0x1.0p-52;  // exact in double, strictly inside (0,1)
  *u2 = ((double)(b >> 12) + 0.5) * 0x1.0p-52;
}

// grid: (ceil(K/256), ceil(cs/2)); thread -> (sample k, pair j); writes rows 2j, 2j+1 coalesced in k
__global__ void __launch_bounds__(256) philox_normals_kernel(double *Z, long long ldk, int cs, int K,
                                                              long long k0, uint64_t seed, uint32_t step,
                                                              uint32_t iter, const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (k >= K) return;
  double u1, u2;
  philox_u2(seed, (uint32_t)(k0 + k), (uint32_t)j, iter, step, &u1, &u2);
  const double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586 * u2;
  double sn, cn;
  sincos(ang, &sn, &cn);
  Z[(size_t)(2 * j) * ldk + k] = rad * cn;
  if (2 * j + 1 < cs) Z[(size_t)(2 * j + 1) * ldk + k] = rad * sn;
}

void launch_philox_normals(double *Z, long long ldk, int cs, int K, long long k0, uint64_t seed,
                           uint32_t step, uint32_t iter, const int *stop, cudaStream_t s) {
  dim3 grid((K + 255) / 256, (cs + 1) / 2);
  philox_normals_kernel<<<grid, 256, 0, s>>>(Z, ldk, cs, K, k0, seed, step, iter, stop);
}

__global__ void philox_uniforms_kernel(double *u, int K, uint64_t seed, uint32_t step, uint32_t iter,
                                       const int *stop) {
  if (stop && *stop) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  double u1, u2;
  philox_u2(seed, (uint32_t)i, 0u, iter | (1u << 24), step, &u1, &u2);
  u[i] = u1;
}

void launch_philox_uniforms(double *u, int K, uint64_t seed, uint32_t step, uint32_t iter, const int *stop,
                            cudaStream_t s) {
  philox_uniforms_kernel<<<(K + 255) / 256, 256, 0, s>>>(u, K, seed, step, iter, stop);
}

// E[r][k] = Σ_{j in block(r), j<=r} L[r][j] Z[j][k] for block-diagonal L with bs x bs blocks
// (bs = 1: diagonal). Lt is L stored ROW-major (Lt[r*cs + j]), so a row's j-range is contiguous.
__global__ void __launch_bounds__(256) apply_L_block_kernel(const double *Lt, int cs, int bs, const double *Z,
                                                             double *E, long long ldk, int K,
                                                             const int *stop) {
  if (stop && *stop) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (k >= K) return;
  const int j0 = (r / bs) * bs;
  double acc = 0.0;
  for (int j = j0; j <= r; ++j) acc += __ldg(Lt + (size_t)r * cs + j) * Z[(size_t)j * ldk + k];
  E[(size_t)r * ldk + k] = acc;
}

// Dense lower-triangular E = L Z. CTA = 64 samples x 4 row-groups; it owns a 32-row block of E
// (8 rows per thread) and walks j in chunks of 32 rows of Z staged in shared memory. L is read
// through the row-major copy Lt with warp-uniform (broadcast) 16-byte loads.
constexpr int AL_KT = 64, AL_RB = 32, AL_JC = 32;
__global__ void __launch_bounds__(256) apply_L_dense_kernel(const double *__restrict__ Lt, int cs,
                                                             const double *__restrict__ Z,
                                                             double *__restrict__ E, long long ldk, int K,
                                                             const int *stop) {
  if (stop && *stop) return;
  __shared__ double Zs[AL_JC][AL_KT];
  const int tx = threadIdx.x & (AL_KT - 1), ty = threadIdx.x / AL_KT;  // ty in 0..3
  const int kbase = blockIdx.x * AL_KT, i0 = blockIdx.y * AL_RB;
  const int k = kbase + tx;
  double acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.0;
  const int jend = min(i0 + AL_RB, cs);  // rows of this block need j < jend
  for (int jc = 0; jc < jend; jc += AL_JC) {
    __syncthreads();
    for (int e = threadIdx.x; e < AL_JC * AL_KT; e += 256) {
      const int jj = e / AL_KT, kk = e % AL_KT;
      const int j = jc + jj, kg = kbase + kk;
      Zs[jj][kk] = (j < cs && kg < K) ? Z[(size_t)j * ldk + kg] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + ty * 8 + q;
      if (i >= cs) continue;
      const int jmax = min(AL_JC, i - jc + 1);  // j <= i
      const double *Lrow = Lt + (size_t)i * cs + jc;
      double a = acc[q];
      for (int jj = 0; jj < jmax; ++jj) a = fma(__ldg(Lrow + jj), Zs[jj][tx], a);
      acc[q] = a;
    }
  }
  if (k < K) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + ty * 8 + q;
      if (i < cs) E[(size_t)i * ldk + k] = acc[q];
    }
  }
}

void launch_apply_L(const double *Lt, int cs, int bs, const double *Z, double *E, long long ldk, int K,
                    const int *stop, cudaStream_t s) {
  if (bs < cs) {
    dim3 grid((K + 255) / 256, cs);
    apply_L_block_kernel<<<grid, 256, 0, s>>>(Lt, cs, bs, Z, E, ldk, K, stop);
  } else {
    dim3 grid((K + AL_KT - 1) / AL_KT, (cs + AL_RB - 1) / AL_RB);
    apply_L_dense_kernel<<<grid, 256, 0, s>>>(Lt, cs, Z, E, ldk, K, stop);
  }
}

// ---- ABI layout conversion: Julia cs x K column-major  <->  device [cs][ldk] --------------------
// 32x32 shared-memory tile transpose (padded against bank conflicts); `shift` (nullable, cs) is
// added on the way out: E .+ (pol.U − U_orig), POL:370,468,602,668,739,814.
__global__ void transpose_in_kernel(const double *__restrict__ cm, double *__restrict__ dev, int cs, int K,
                                    long long ldk) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int dk = threadIdx.y; dk < 32; dk += blockDim.y) {  // read: r fastest
    const int r = r0 + threadIdx.x, k = k0 + dk;
    tile[dk][threadIdx.x] = (r < cs && k < K) ? cm[(size_t)k * cs + r] : 0.0;
  }
  __syncthreads();
  for (int dr = threadIdx.y; dr < 32; dr += blockDim.y) {  // write: k fastest
    const int r = r0 + dr, k = k0 + threadIdx.x;
    if (r < cs && k < K) dev[(size_t)r * ldk + k] = tile[threadIdx.x][dr];
  }
}

__global__ void transpose_out_kernel(const double *__restrict__ dev, double *__restrict__ cm, int cs, int K,
                                     long long ldk, const double *__restrict__ shift_a,
                                     const double *__restrict__ shift_b) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int dr = threadIdx.y; dr < 32; dr += blockDim.y) {
    const int r = r0 + dr, k = k0 + threadIdx.x;
    double v = 0.0;
    if (r < cs && k < K) {
      v = dev[(size_t)r * ldk + k];
      if (shift_a) v = v + (shift_a[r] - shift_b[r]);
    }
    tile[dr][threadIdx.x] = v;
  }
  __syncthreads();
  for (int dk = threadIdx.y; dk < 32; dk += blockDim.y) {
    const int r = r0 + threadIdx.x, k = k0 + dk;
    if (r < cs && k < K) cm[(size_t)k * cs + r] = tile[threadIdx.x][dk];
  }
}

void launch_transpose_in(const double *colmajor, double *dev, int cs, int K, long long ldk, cudaStream_t s) {
  dim3 grid((K + 31) / 32, (cs + 31) / 32), block(32, 8);
  transpose_in_kernel<<<grid, block, 0, s>>>(colmajor, dev, cs, K, ldk);
}

void launch_transpose_out(const double *dev, double *colmajor, int cs, int K, long long ldk,
                          const double *shift_a, const double *shift_b, cudaStream_t s) {
  dim3 grid((K + 31) / 32, (cs + 31) / 32), block(32, 8);
  transpose_out_kernel<<<grid, block, 0, s>>>(dev, colmajor, cs, K, ldk, shift_a, shift_b);
}

}  // namespace mpopis
