// chain_bench.cu — where do the cycles of ONE velocity sub-step go? (round 2 diagnostic, profiles/r2_chain_bench.txt)
//
// One warp alone on an SM runs the loop-carried recurrence of rollout_split.cu's velocity warp for many sub-steps and
// reports clock64() cycles per sub-step, for the full sub-step and for variants with one ingredient removed each, plus
// dependent-issue latencies of the instruction pairs tools/fp64_mix_bench.cu did not cover (MUFU.RCP64H, DSETP -> select,
// integer work on the high word of a double feeding an FP64 instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_bench tools/chain_bench.cu && ./chain_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_seed(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}
__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }
__device__ __forceinline__ double with_opposite_sign(double mag, int shi) {
  return __hiloint2double(__double2hiint(mag) | (~shi & (int)0x80000000), __double2loint(mag));
}

struct Consts {
  double l_f, l_r, C_af, C_ar, cI1, cI2, cm, kx, cmCD0, ddt;
  double fxf, fxr, fymax_f, fymax_r, thr_f, thr_r, c2_f, c2_r, c3_f, c3_r, sdl, cdl;
};

// VAR: 0 full; 1 no validity/integer bookkeeping; 2 selects replaced by the cubic; 3 reciprocal seeds replaced by a
// (numerically wrong, latency-free) linear guess; 4 = 2 + 3; 5 = level-ordered source exactly as rollout_split.cu
// UNROLL: how many copies of the sub-step the loop body holds (code size ≈ UNROLL x 1.4 KB): the instruction-cache probe
template <int VAR, int UNROLL = 5>
__global__ void substep_kernel(const Consts c, int iters, double *out, long long *cycles) {
  double Vx = 10.0 + 0.01 * threadIdx.x, Vy = 0.1, psid = 0.05, sd = 0.02, cd = 0.9998;
  const double cI1fxf = c.cI1 * c.fxf, cmfxf = c.cm * c.fxf, cmfxr = c.cm * c.fxr, nddt = -c.ddt;
  int bad = 0;
  const int hvx0 = hi32(Vx), brake_mask = 0;
  const long long t0 = clock64();
#pragma unroll UNROLL
  for (int i = 0; i < iters; ++i) {
    const double ns = fma(sd, c.cdl, cd * c.sdl);
    const double nc = fma(cd, c.cdl, -(sd * c.sdl));
    sd = ns, cd = nc;
    const int hvx = hi32(Vx);
    const bool fwd = VAR == 1 ? true : hvx >= 0;
    const double rx = (VAR == 3 || VAR == 4) ? fma(Vx, -0.01, 0.2) : rcp_seed(Vx);
    const double yf = fma(c.l_f, psid, Vy), yr = fma(-c.l_r, psid, Vy);
    const double wx = Vx * nddt, wy = Vy * c.ddt;
    const double k3 = fma(Vx, c.kx, cmfxr + (VAR == 1 ? -c.cmCD0 : with_opposite_sign(c.cmCD0, hvx)));
    const double vxsd = Vx * sd, vxcd = Vx * cd;
    const double ex = fma(-Vx, rx, 1.0);
    const double t0r = yr * rx;
    const double k2 = fma(psid, wx, Vy);
    const double k3b = fma(psid, wy, k3);
    const double k1 = fma(cI1fxf, sd, psid);
    const double qI = c.cI1 * cd, qy = c.cm * cd, qx = c.cm * sd;
    const double fxfsd = c.fxf * sd;
    const double num = fma(yf, cd, -vxsd), den = fma(yf, sd, vxcd);
    const double px = fma(ex, ex, ex);
    const double Kx = fma(cmfxf, cd, k3b);
    const double r0 = (VAR == 3 || VAR == 4) ? fma(den, -0.01, 0.2) : rcp_seed(den);
    const double ta_r = fma(t0r, px, t0r);
    if (VAR != 1)
      bad |= ((hvx ^ hvx0) & brake_mask) | ((hi32(den) & 0x7ff00000) - 0x00100000) | ((hvx & 0x7ff00000) - 0x00100000);
    const double e = fma(-den, r0, 1.0), t0 = num * r0;
    const double atr = fabs(ta_r);
    const double ur = ta_r * atr, vr = fma(-c.c3_r, atr, c.c2_r), x1r = -c.C_ar * ta_r;
    const double pe = fma(e, e, e);
    const double cubic_r = fma(ur, vr, x1r);
    const double ta = fma(t0, pe, t0);
    double fyr, fyf;
    if (VAR == 2 || VAR == 4) fyr = cubic_r;
    else fyr = (fwd & (atr < c.thr_r)) ? cubic_r : with_opposite_sign(c.fymax_r, hi32(yr));
    const double at = fabs(ta);
    const double u = ta * at, v = fma(-c.c3_f, at, c.c2_f), x1 = -c.C_af * ta;
    const double Kp = fma(-c.cI2, fyr, k1);
    const double Ky = fma(c.cm, fxfsd + fyr, k2);
    const double cubic = fma(u, v, x1);
    if (VAR == 2 || VAR == 4) fyf = cubic;
    else fyf = ((hi32(den) >= 0) & (at < c.thr_f)) ? cubic : with_opposite_sign(c.fymax_f, fwd ? hi32(num) : hi32(yf));
    psid = fma(qI, fyf, Kp), Vy = fma(qy, fyf, Ky), Vx = fma(-qx, fyf, Kx);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = Vx + Vy + psid + sd + cd + bad;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// dependent-latency probes: OP 0 DFMA chain; 1 MUFU.RCP64H -> DFMA; 2 DSETP -> select -> DFMA; 3 hi-word LOP3 -> DFMA;
// 4 DMUL -> DFMA; 5 DSETP(|x|) -> FSEL pair only
template <int OP>
__global__ void latency_kernel(int iters, double seed, double *out, long long *cycles) {
  double x = seed + 1e-3 * threadIdx.x, y = 0.5;
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = fma(x, 0.999, 0.001);
    if (OP == 1) x = fma(rcp_seed(x), 0.5, 0.75);
    if (OP == 2) x = fma((x < 1.5) ? x : y, 0.999, 0.001);
    if (OP == 3) x = fma(with_opposite_sign(x, hi32(y)), -0.999, 0.001);
    if (OP == 4) x = fma(x * 0.999, 0.999, 0.002);
    if (OP == 5) x = (fabs(x) < 1.5) ? fma(x, 0.999, 0.001) : y;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  Consts c{1.6, 1.4, 180000.0, 220000.0, 0.01 * 1.6 / 2200.0, 0.01 * 1.4 / 2200.0, 0.01 / 1000.0, 1.0 - 1e-5 * 3.0, 1e-5 * 218.0,
           0.01, 2000.0, 1500.0, 5000.0, 5200.0, 0.08, 0.07, 6e7, 7e7, 2e9, 3e9, 0.0017, 0.9999986};
  double *out;
  long long *cyc, h = 0;
  cudaMalloc(&out, sizeof(double) * 4096);
  cudaMalloc(&cyc, sizeof(long long) * 64);
  const int iters = 20000;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
#define RUN_SUB(V, name)                                                                   \
  substep_kernel<V><<<1, 32>>>(c, iters, out, cyc);                                        \
  substep_kernel<V><<<1, 32>>>(c, iters, out, cyc);                                        \
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);                                   \
  printf("sub-step %-58s %8.1f cycles per sub-step (one warp alone)\n", name, (double)h / iters);
  RUN_SUB(0, "full (rollout_split.cu velocity warp)")
  RUN_SUB(1, "without validity / sign bookkeeping on the integer pipe")
  RUN_SUB(2, "tyre-force selects replaced by the cubic")
  RUN_SUB(3, "reciprocal seeds (MUFU.RCP64H) replaced by a linear guess")
  RUN_SUB(4, "no selects and no MUFU")
#define RUN_UNR(U, name)                                                                   \
  substep_kernel<0, U><<<1, 32>>>(c, iters, out, cyc);                                     \
  substep_kernel<0, U><<<1, 32>>>(c, iters, out, cyc);                                     \
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);                                   \
  printf("sub-step %-58s %8.1f cycles per sub-step (one warp alone)\n", name, (double)h / iters);
  RUN_UNR(1, "full, loop body = 1 sub-step  (~1.4 KB of code)")
  RUN_UNR(2, "full, loop body = 2 sub-steps")
  RUN_UNR(10, "full, loop body = 10 sub-steps (~14 KB)")
  RUN_UNR(20, "full, loop body = 20 sub-steps (~28 KB)")
  RUN_UNR(40, "full, loop body = 40 sub-steps (~56 KB)")
  RUN_UNR(100, "full, loop body = 100 sub-steps (~140 KB)")
#define RUN_LAT(O, name)                                                                   \
  latency_kernel<O><<<1, 32>>>(iters, 1.25, out, cyc);                                     \
  latency_kernel<O><<<1, 32>>>(iters, 1.25, out, cyc);                                     \
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);                                   \
  printf("latency  %-58s %8.1f cycles per iteration\n", name, (double)h / iters);
  RUN_LAT(0, "DFMA -> DFMA")
  RUN_LAT(4, "DMUL -> DFMA")
  RUN_LAT(1, "MUFU.RCP64H -> DFMA")
  RUN_LAT(2, "DSETP -> select -> DFMA")
  RUN_LAT(5, "DSETP(|x|) -> predicated DFMA / select")
  RUN_LAT(3, "integer sign transfer on the high word -> DFMA")
  if (cudaDeviceSynchronize() != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
