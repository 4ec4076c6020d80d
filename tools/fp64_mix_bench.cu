// fp64_mix_bench.cu — issue cost of the FP64-pipe instructions the rollout kernel is made of, on the SM sub-partition.
//
// Round 1 ended on an open question (profiles/README.md, item 16): the rollout kernel sits on a throughput plateau at
// ≈4.1 cycles per FP64 warp-instruction while a dependent-DFMA chain sustains 2.2. This standalone micro-benchmark
// measures, per opcode and operand kind, (a) the dependent-issue latency (one warp per scheduler, one chain) and
// (b) the sustained cost per warp-instruction with 1..4 warps per scheduler and 8 independent chains per thread:
//   DFMA r,r,r | DFMA with a constant-bank operand | DMUL | DADD | DSETP+FSEL | the rollout kernel's mix.
// One CTA per SM (dynamic shared memory forces it), 4·n warps per CTA -> n warps per scheduler.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/fp64_mix_bench.cu -o /tmp/fp64_mix && /tmp/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

constexpr int ITERS = 4096;

// KIND: 0 DFMA reg, 1 DFMA const operand, 2 DMUL, 3 DADD, 4 DSETP+select, 5 mix (4 DFMA, 3 DMUL, 1 DADD per 8)
template <int KIND, int CHAINS>
__global__ void bench(double *out, long long *cycles, double c0, double c1) {
  extern __shared__ double pad[];
  double v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  // per-thread values: plain register operands (a uniform value would be promoted to a uniform register)
  const double a = 1.0000001 + 1e-13 * threadIdx.x, b = 1e-12 + 1e-22 * threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
      if (KIND == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(c0), "d"(b));
      if (KIND == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(a));
      if (KIND == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(b));
      if (KIND == 4)
        asm volatile("{ .reg .pred p; setp.lt.f64 p, %0, %1; selp.f64 %0, %2, %0, p; }" : "+d"(v[i]) : "d"(c1), "d"(a));
      if (KIND == 5) {
        const int m = i & 7;
        if (m < 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(m & 1 ? c0 : a), "d"(b));
        else if (m < 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(a));
        else asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(b));
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + pad[0] * 0.0;
  if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int KIND, int CHAINS>
static double run(int warps_per_sched, int sms) {
  const int threads = 128 * warps_per_sched;  // 4 schedulers x n warps
  double *out;
  long long *cyc;
  cudaMalloc(&out, sizeof(double) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms * threads / 32);
  const size_t smem = 120 * 1024;  // one CTA per SM
  cudaFuncSetAttribute(bench<KIND, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<KIND, CHAINS><<<sms, threads, smem>>>(out, cyc, 1.0000001, 0.5);
  bench<KIND, CHAINS><<<sms, threads, smem>>>(out, cyc, 1.0000001, 0.5);
  cudaDeviceSynchronize();
  std::vector<long long> h(sms * threads / 32);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long c : h) mx = c > mx ? c : mx;
  cudaFree(out), cudaFree(cyc);
  // cycles per warp-instruction seen by one scheduler: all its warps' instructions / the slowest warp's cycles
  return (double)mx / ((double)ITERS * CHAINS * warps_per_sched);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %d MHz\n", prop.name, sms, prop.clockRate / 1000);
  const char *names[] = {"DFMA r,r,r", "DFMA r,c[],r", "DMUL", "DADD", "DSETP+SEL", "mix 4 DFMA(2 const)+3 DMUL+1 DADD"};
  printf("%-36s %10s | cycles per warp-instruction per scheduler, 8 chains, n warps/scheduler:\n", "opcode", "dep. lat.");
  printf("%-36s %10s | %8s %8s %8s %8s\n", "", "(1 chain)", "n=1", "n=2", "n=3", "n=4");
#define ROW(K)                                                                                              \
  printf("%-36s %10.2f | %8.2f %8.2f %8.2f %8.2f\n", names[K], run<K, 1>(1, sms), run<K, 8>(1, sms), run<K, 8>(2, sms), \
         run<K, 8>(3, sms), run<K, 8>(4, sms));
  ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5)
  return 0;
}
