// sampling.cu — G1: Gaussian control-perturbation draw.
//
// Replaces `E = rand(pol.rng, MvNormal(Σ′), K)` (POL:308,359,448,556,657,724,797; :mppi POL:193):
// Z ~ N(0, I) (cs x K), E = L Z with L the lower Cholesky factor of Σ′.
// Julia's MersenneTwister stream cannot be reproduced; the engine's generator is counter-based
// Philox4x32-10 keyed by the policy seed, counter = (global sample k, pair j, AIS iteration |
// purpose<<24, control-step counter), Box–Muller in FP64. Because the counter is the GLOBAL
// sample index, the draws are independent of how K is sharded across GPUs. oracle/ restates the
// same generator (tests compare them to ~1e-15).
#include <cstdlib>

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
// step_dev (nullable): the control-step counter lives in device memory so that a captured CUDA graph of the control
// step replays with the right counter; `step` is used when it is null (parity surface sample_normals).
__global__ void __launch_bounds__(256) philox_normals_kernel(double *Z, long long ldk, int cs, int K,
                                                              long long k0, uint64_t seed, uint32_t step,
                                                              const unsigned *step_dev, uint32_t iter,
                                                              const int *stop) {
  if (stop && *stop) return;
  if (step_dev) step = *step_dev;
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
                           uint32_t step, const unsigned *step_dev, uint32_t iter, const int *stop,
                           cudaStream_t s) {
  dim3 grid((K + 255) / 256, (cs + 1) / 2);
  philox_normals_kernel<<<grid, 256, 0, s>>>(Z, ldk, cs, K, k0, seed, step, step_dev, iter, stop);
}

__global__ void philox_uniforms_kernel(double *u, int K, uint64_t seed, uint32_t step, const unsigned *step_dev,
                                       uint32_t iter, const int *stop) {
  if (stop && *stop) return;
  if (step_dev) step = *step_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  double u1, u2;
  philox_u2(seed, (uint32_t)i, 0u, iter | (1u << 24), step, &u1, &u2);
  u[i] = u1;
}

void launch_philox_uniforms(double *u, int K, uint64_t seed, uint32_t step, const unsigned *step_dev, uint32_t iter,
                            const int *stop, cudaStream_t s) {
  philox_uniforms_kernel<<<(K + 255) / 256, 256, 0, s>>>(u, K, seed, step, step_dev, iter, stop);
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

// Dense lower-triangular E = L Z on the FP64 tensor path: mma.sync.m8n8k4.f64 (SASS: DMMA). FP64 has no tcgen05 kind,
// so this warp-level MMA is the tensor path that exists for doubles; one instruction performs 256 FMAs from 2+2
// register operands per lane, i.e. 8x less shared-memory traffic per FMA than a 4x4 DFMA register tile (round 1
// measured that tile bound by shared-memory wavefronts, and two other DMMA tilings slower than the one below:
// profiles/README.md items 5, 15; they were removed in round 2). Fragment layouts (PTX ISA, m8n8k4 .f64): A[g][t],
// B[t][g], C[g][2t..2t+1] with g = lane/4, t = lane%4.
//
// ONE CTA owns a 64-sample column tile and ALL rows of a 104-row block (13 row fragments of 8), so Z streams through shared memory exactly once per block, double-buffered with cp.async
// (16-byte vectors, zero-filled past cs / ldk); L (80 KB for cs = 100, shared by every CTA) is read through
// the read-only L1 path straight into A fragments. Seven warps: warp w owns row fragments {w, 12 − w}
// (w = 6: fragment 6 alone) — fragment f needs j < 8f + 8, so each pair costs 14 units and the triangle is
// balanced inside the CTA. Z pitch 68 doubles (≡ 4 mod 16) keeps the B-fragment
// loads conflict-free. ncu on the first cut (2 buffers, 2 barriers per chunk, L read at the point of use): 27 % of
// the stall samples on the barrier, 23 % on the L loads (long scoreboard), DMMA pipe 44 % busy — hence the 3-stage
// ring (one barrier per chunk) and A fragments fetched one chunk ahead. Measured DMMA cost on B200: ≈24.5 pipe
// cycles per m8n8k4 per SM sub-partition (≈42 FMA/clk/SM, i.e. ~2/3 of the DFMA pipe's 64).
constexpr int D2_NF = 13, D2_RB = 8 * D2_NF, D2_BN = 64, D2_BJ = 16, D2_ZP = 68, D2_NW = 7, D2_NT = 32 * D2_NW;

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(D2_NT, 2) apply_L_dmma2_kernel(const double *__restrict__ Lt, int cs,
                                                                  const double *__restrict__ Z,
                                                                  double *__restrict__ E, long long ldk, int K,
                                                                  const int *stop) {
  if (stop && *stop) return;
  __shared__ __align__(16) double Zs[3][D2_BJ][D2_ZP];  // 3-stage ring: one barrier per chunk
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int kbase = blockIdx.x * D2_BN, i0 = blockIdx.y * D2_RB;
  const int jend = min(i0 + D2_RB, cs);  // rows of this block only need j < jend (zeros above the diagonal)
  const int nchunks = (jend + D2_BJ - 1) / D2_BJ;
  // this warp's row fragments: fb, and fa (< fb) unless w == 6; first row and exclusive j bound of each
  const int fb = D2_NF - 1 - w, fa = w;
  const bool has_a = w < D2_NF / 2;
  const int rowb = i0 + 8 * fb + g, rowa = i0 + 8 * fa + g;
  const int jmax_b = (i0 + 8 * fb < cs) ? min(i0 + 8 * fb + 8, cs) : 0;
  const int jmax_a = (has_a && i0 + 8 * fa < cs) ? min(i0 + 8 * fa + 8, cs) : 0;
  const double *La = Lt + (size_t)min(rowa, cs - 1) * cs, *Lb = Lt + (size_t)min(rowb, cs - 1) * cs;
  const bool va = rowa < cs, vb = rowb < cs;
  double acca[8][2], accb[8][2];
#pragma unroll
  for (int c = 0; c < 8; ++c) acca[c][0] = acca[c][1] = accb[c][0] = accb[c][1] = 0.0;

  auto stage = [&](int chunk) {
    if (chunk < nchunks) {
      const int jc = chunk * D2_BJ, buf = chunk % 3;
      for (int e = threadIdx.x; e < D2_BJ * (D2_BN / 2); e += D2_NT) {
        const int jj = e / (D2_BN / 2), v = e % (D2_BN / 2);
        const int j = jc + jj;
        const long long kg = (long long)kbase + 2 * v;
        const bool pred = j < cs && kg + 1 < ldk;
        cp_async16_zfill(&Zs[buf][jj][2 * v], pred ? (const void *)(Z + (size_t)j * ldk + kg) : (const void *)Z, pred);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // always: keeps the group count uniform
  };
  // A fragments (rows of L) of one chunk: 4 k-steps x {a, b}; read through L1 one chunk ahead of their use
  auto load_A = [&](int chunk, double (&A)[4][2]) {
    const int jc = chunk * D2_BJ;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = jc + 4 * q + t;
      A[q][0] = (va && jc + 4 * q < jmax_a && j < cs) ? __ldg(La + j) : 0.0;
      A[q][1] = (vb && jc + 4 * q < jmax_b && j < cs) ? __ldg(Lb + j) : 0.0;
    }
  };

  double Ac[4][2], An[4][2];
  stage(0);
  stage(1);
  load_A(0, Ac);
  for (int c = 0; c < nchunks; ++c) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // groups 0..c have landed (c + 1 may be in flight)
    __syncthreads();  // ... for every thread; and everyone finished chunk c − 1, whose buffer stage(c + 2) refills
    stage(c + 2);
    if (c + 1 < nchunks) load_A(c + 1, An);
    const int buf = c % 3, jc = c * D2_BJ;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j4 = 4 * q;
      const bool do_b = jc + j4 < jmax_b, do_a = jc + j4 < jmax_a;  // warp-uniform (a partial block may hold a only)
      if (do_a | do_b) {
        double bf[8];
#pragma unroll
        for (int cf = 0; cf < 8; ++cf) bf[cf] = Zs[buf][j4 + t][cf * 8 + g];
        if (do_b) {
#pragma unroll
          for (int cf = 0; cf < 8; ++cf)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(accb[cf][0]), "+d"(accb[cf][1])
                         : "d"(Ac[q][1]), "d"(bf[cf]));
        }
        if (do_a) {
#pragma unroll
          for (int cf = 0; cf < 8; ++cf)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acca[cf][0]), "+d"(acca[cf][1])
                         : "d"(Ac[q][0]), "d"(bf[cf]));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) Ac[q][0] = An[q][0], Ac[q][1] = An[q][1];
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int cf = 0; cf < 8; ++cf) {
    const int k = kbase + cf * 8 + 2 * t;
    if (vb && jmax_b > 0) {
      double *dst = E + (size_t)rowb * ldk + k;
      if (k + 1 < K) *reinterpret_cast<double2 *>(dst) = make_double2(accb[cf][0], accb[cf][1]);
      else if (k < K) dst[0] = accb[cf][0];
    }
    if (has_a && va && jmax_a > 0) {
      double *dst = E + (size_t)rowa * ldk + k;
      if (k + 1 < K) *reinterpret_cast<double2 *>(dst) = make_double2(acca[cf][0], acca[cf][1]);
      else if (k < K) dst[0] = acca[cf][0];
    }
  }
}

void launch_apply_L(const double *Lt, int cs, int bs, const double *Z, double *E, long long ldk, int K,
                    const int *stop, cudaStream_t s) {
  if (bs < cs) {
    dim3 grid((K + 255) / 256, cs);
    apply_L_block_kernel<<<grid, 256, 0, s>>>(Lt, cs, bs, Z, E, ldk, K, stop);
  } else {
    dim3 grid((K + D2_BN - 1) / D2_BN, (cs + D2_RB - 1) / D2_RB);
    apply_L_dmma2_kernel<<<grid, D2_NT, 0, s>>>(Lt, cs, Z, E, ldk, K, stop);
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
