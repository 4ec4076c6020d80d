// chol_tile.cuh — right-looking Cholesky of an n x n SPD matrix (n <= 16 R) held in the registers of a 16 x 16 thread
// grid: thread (ty, tx) owns W[ty + 16a][tx + 16b], a, b < R; only the lower triangle is meaningful. Shared by
// chol_reg_kernel / chol_cov_kernel (linalg.cu) and ce_small_adapt_kernel (small_adapt.cu).
//
// Square-root free: column j is left unscaled (W[i][j] = l_ij·l_jj), the trailing update divides by the pivot
// d_j = l_jj²; L is formed at the end as W[i][j] / sqrt(d_j). One barrier per column, the pivot column travels through a
// double-buffered shared-memory vector. Round 2 (profiles/r2 notes: the kernel was 620 cycles per column, 105
// instructions per warp, on a chain barrier -> LDS d -> DSETP/BRA -> MUFU.RCP64H -> 4 DFMA -> DMUL -> DFMA -> FSEL -> STS):
//   * look-ahead reciprocal: during step j every thread updates its element of the pivot's diagonal block first and
//     starts 1/x on it in the same straight-line block (the compiler interleaves it with the bulk update); the thread that
//     actually owns W[j+1][j+1] publishes (d, 1/d). After the barrier of step j + 1 the reciprocal is a shared-memory
//     read — the MUFU + Newton chain and the positivity branch are off the critical path (except once per 16 columns,
//     where the next pivot sits in the next diagonal block and its owner inverts it after the bulk update: peeling that
//     step would double a code footprint that already stalls on instruction fetch);
//   * masks on operands instead of results: rows at or above the pivot get c_i = 0, columns at or left of it c_k = 0
//     (an FMA with a zero factor leaves its accumulator bit-exact), replacing a DFMA + 2 FSEL per masked element. The
//     strictly-upper entries of diagonal blocks then accumulate finite junk that nothing ever reads;
//   * a failed pivot (not > 0, NaN) raises a shared flag that is read once after the loop.
// Bit-identical to the round-1 loop on every entry of L.
#pragma once
#include <cuda_runtime.h>

namespace mpopis {

template <int R>
struct CholTileSmem {
  double col[2][16 * R];
  double dg[16 * R];  // pivots d_j
  double pinv[2];     // 1/d_j, parity-buffered
  int failed;
};

__device__ __forceinline__ double chol_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  return fma(fma(-d, r, 1.0), r, r);
}

// Factors the tile in place; returns false when a pivot was not positive. Must be called by all 256 threads; the
// caller's writes to `w` need no barrier, `sm` must not be in use. On return sm.dg[j] holds the pivots (a barrier has
// been passed since the last write).
template <int R>
__device__ __forceinline__ bool chol_tile_factor(double (&w)[R][R], int n, CholTileSmem<R> &sm, int tx, int ty) {
  if (threadIdx.x == 0) {
    const double d = w[0][0];
    sm.failed = !(d > 0.0);
    sm.dg[0] = d, sm.pinv[0] = chol_rcp(d);
  }
#pragma unroll
  for (int jb = 0; jb < R; ++jb) {
    for (int jt = 0; jt < 16; ++jt) {
      const int j = 16 * jb + jt, pb = jt & 1;
      if (j >= n) break;
      if (tx == jt) {  // publish column j (final: every update of the steps < j has been applied)
#pragma unroll
        for (int a = jb; a < R; ++a) sm.col[pb][ty + 16 * a] = w[a][jb];
      }
      __syncthreads();
      const double inv_d = sm.pinv[pb];
      double ci[R], ck[R];
#pragma unroll
      for (int a = jb; a < R; ++a) {
        ci[a] = sm.col[pb][ty + 16 * a] * inv_d;
        ck[a] = sm.col[pb][tx + 16 * a];
      }
      if (ty <= jt) ci[jb] = 0.0;  // rows at or above the pivot
      if (tx <= jt) ck[jb] = 0.0;  // columns at or left of the pivot
      // look-ahead: while the next pivot lives in this diagonal block (jt < 15) every thread updates its element of the
      // block first and starts 1/x on it; only the owner's value is the pivot, the others' are discarded.
      const bool own = (tx == ((jt + 1) & 15)) && (ty == ((jt + 1) & 15)) && j + 1 < n;
      w[jb][jb] = fma(-ci[jb], ck[jb], w[jb][jb]);
      {
        const double dn = w[jb][jb];
        const double rn = chol_rcp(dn);
        if (own && jt < 15) {
          sm.dg[j + 1] = dn, sm.pinv[pb ^ 1] = rn;
          if (!(dn > 0.0)) sm.failed = 1;
        }
      }
#pragma unroll
      for (int a = jb; a < R; ++a)
#pragma unroll
        for (int b = jb; b <= a; ++b)
          if (!(a == jb && b == jb)) w[a][b] = fma(-ci[a], ck[b], w[a][b]);
      // block boundary (once per 16 columns): the next pivot is element (0, 0) of the next diagonal block, final now
      if (jt == 15 && jb + 1 < R && own) {
        const double dn = w[jb + 1 < R ? jb + 1 : jb][jb + 1 < R ? jb + 1 : jb];
        sm.dg[j + 1] = dn, sm.pinv[pb ^ 1] = chol_rcp(dn);
        if (!(dn > 0.0)) sm.failed = 1;
      }
    }
  }
  __syncthreads();
  return sm.failed == 0;
}

// l_kk = sqrt(d_k) and 1/l_kk once per column (into the free column buffers), so that forming L costs one multiply per
// entry: l_ik = W[i][k]·(1/l_kk), the scaling LAPACK's dpotf2 applies (DSCAL by 1/a_jj). An IEEE square root and a
// division per ENTRY were 49 + 49 inline expansions per thread — half of the kernel's static code (the kernels stall
// mostly on instruction fetch: ncu `no_instruction`) and a quarter of its executed instructions.
template <int R>
__device__ __forceinline__ void chol_tile_finish(CholTileSmem<R> &sm, int n) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double r = sqrt(sm.dg[k]);
    sm.col[0][k] = r, sm.col[1][k] = 1.0 / r;
  }
  __syncthreads();
}
// L[i][k] of the factored tile for the entry (a, b) this thread owns (k <= i); after chol_tile_finish
template <int R>
__device__ __forceinline__ double chol_tile_entry(const double (&w)[R][R], const CholTileSmem<R> &sm, int a, int b, int i,
                                                  int k) {
  return k == i ? sm.col[0][k] : w[a][b] * sm.col[1][k];
}

}  // namespace mpopis
