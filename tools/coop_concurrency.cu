// coop_concurrency.cu — can a kernel launched on stream B start while a kernel on stream A spins waiting for it?
// (round 2 diagnostic for the peer-memory collectives between loop-back ranks, csrc/comm.cu)
//   stream A: spin kernel (waits for *flag == 1, gives up after ~2 s)
//   stream B: a kernel that sets *flag = 1, launched (a) normally, (b) cooperatively, (c) as one 8-CTA cluster
// Prints for each flavour whether the spinner saw the flag (concurrent) or timed out (serialised behind the spinner).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_concurrency tools/coop_concurrency.cu && ./coop_concurrency
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void spin_kernel(volatile int *flag, int *result) {
  const long long t0 = clock64();
  while (*flag == 0) {
    __nanosleep(100);
    if (clock64() - t0 > 4000000000LL) {
      *result = 0;  // timed out
      return;
    }
  }
  *result = 1;
}
__global__ void set_kernel(volatile int *flag) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1;
}
__global__ void set_coop_kernel(volatile int *flag) {
  cg::this_grid().sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1;
}
__global__ void __cluster_dims__(8, 1, 1) set_cluster_kernel(volatile int *flag) {
  cg::this_cluster().sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 1;
}

int main() {
  int *flag, *result;
  cudaMalloc(&flag, 4);
  cudaMalloc(&result, 4);
  cudaStream_t a, b;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  // load everything first (lazy module loading would serialise on its own)
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, spin_kernel);
  cudaFuncGetAttributes(&fa, set_kernel);
  cudaFuncGetAttributes(&fa, set_coop_kernel);
  cudaFuncGetAttributes(&fa, set_cluster_kernel);
  const char *names[] = {"plain launch", "cooperative launch (5 CTAs)", "cluster launch (8 CTAs)", "cooperative launch (148 CTAs)"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(flag, 0, 4);
      cudaMemset(result, 0xff, 4);
      cudaDeviceSynchronize();
      spin_kernel<<<1, 32, 0, a>>>(flag, result);
      void *args[] = {&flag};
      cudaError_t e = cudaSuccess;
      if (mode == 0) set_kernel<<<5, 256, 0, b>>>(flag);
      if (mode == 1) e = cudaLaunchCooperativeKernel((const void *)set_coop_kernel, dim3(5), dim3(256), args, 0, b);
      if (mode == 2) set_cluster_kernel<<<8, 256, 0, b>>>(flag);
      if (mode == 3) e = cudaLaunchCooperativeKernel((const void *)set_coop_kernel, dim3(148), dim3(256), args, 0, b);
      cudaDeviceSynchronize();
      int r = -1;
      cudaMemcpy(&r, result, 4, cudaMemcpyDeviceToHost);
      printf("%-32s rep %d: %s (%s)\n", names[mode], rep, r == 1 ? "ran CONCURRENTLY with the spinner" : "SERIALISED behind the spinner (timed out)",
             cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e));
    }
  }
  return 0;
}
