// sort.cu — G4: stable ascending arg-sort of the K trajectory costs.
//
// Replaces `order = sortperm(trajectory_cost)` (POL:455, 563). Julia's sortperm is a stable merge
// sort on isless (ties keep index order, −0.0 < +0.0, NaN last). Here every sample gets the
// composite key (order-preserving 64-bit image of the cost, sample index): composite keys are
// unique, so ANY comparison sort of them yields exactly the stable permutation. The sort is a merge
// sort like the reference's: one bitonic pass sorts 2048-key tiles in shared memory, then
// log2(K/2048) merge-path passes merge runs pairwise with every CTA producing one 2048-key output
// tile (6 launches at K = 65 536; round 1 started with a 26-launch LSD radix sort — see
// profiles/README.md for the before/after).
#include <cooperative_groups.h>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace mpopis {

namespace {

constexpr int TILE = 2048;
constexpr unsigned long long KEY_PAD = ~0ULL;

__device__ __forceinline__ unsigned long long key_of(double c) { return cost_key(c); }  // engine.cuh
__device__ __forceinline__ bool lt(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// Sorts tile blockIdx.x of (key_of(costs[i]), i) with a bitonic network in shared memory.
__global__ void __launch_bounds__(1024) tile_sort_kernel(const double *__restrict__ costs, int n,
                                                          unsigned long long *__restrict__ keys,
                                                          int *__restrict__ vals, const int *stop) {
  if (stop && *stop) return;
  __shared__ unsigned long long sk[TILE];
  __shared__ int sv[TILE];
  const int base = blockIdx.x * TILE;
  for (int e = threadIdx.x; e < TILE; e += 1024) {
    const int i = base + e;
    sk[e] = i < n ? key_of(costs[i]) : KEY_PAD;
    sv[e] = i < n ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= TILE; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair, bit j clear
      const int p = i | j;
      const bool up = (i & k) == 0;
      const unsigned long long ka = sk[i], kb = sk[p];
      const int va = sv[i], vb = sv[p];
      if (lt(kb, vb, ka, va) == up) sk[i] = kb, sv[i] = vb, sk[p] = ka, sv[p] = va;
      __syncthreads();
    }
  for (int e = threadIdx.x; e < TILE; e += 1024) {
    const int i = base + e;
    if (i < n) keys[i] = sk[e], vals[i] = sv[e];
  }
}

// number of elements taken from A among the first `diag` outputs of merge(A, B)
__device__ __forceinline__ int merge_path(const unsigned long long *ak, const int *av, int na,
                                          const unsigned long long *bk, const int *bv, int nb, int diag) {
  int lo = max(0, diag - nb), hi = min(diag, na);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int bi = diag - 1 - mid;
    if (lt(bk[bi], bv[bi], ak[mid], av[mid])) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

// Same partition point, found by one warp with a 33-ary search: every lane probes one split and the
// ballot of the (monotone) predicate narrows [lo, hi] 33-fold per round — 3 dependent global-memory
// round trips instead of 15 for the run lengths of K = 65 536 (the pass is latency-, not bandwidth-bound).
__device__ __forceinline__ int merge_path_warp(const unsigned long long *ak, const int *av, int na,
                                               const unsigned long long *bk, const int *bv, int nb, int diag) {
  const int lane = threadIdx.x & 31;
  int lo = max(0, diag - nb), hi = min(diag, na);
  while (hi > lo) {
    const int mid = lo + (int)(((long long)(hi - lo) * (lane + 1)) / 33);  // in [lo, hi)
    const int bi = diag - 1 - mid;
    const bool gt = !lt(bk[bi], bv[bi], ak[mid], av[mid]);  // partition point lies above `mid`
    const int c = __popc(__ballot_sync(0xffffffffu, gt));   // lanes 0..c-1 say "above"
    const int lo_c = __shfl_sync(0xffffffffu, mid, max(c - 1, 0)) + 1;
    const int hi_c = __shfl_sync(0xffffffffu, mid, min(c, 31));
    lo = c > 0 ? lo_c : lo;
    hi = c < 32 ? hi_c : hi;
  }
  return lo;
}

// Merges the pair of sorted runs (length L) that contains output tile `tile` and writes that tile.
__device__ __forceinline__ void merge_tile(const unsigned long long *__restrict__ kin, const int *__restrict__ vin,
                                           int n, int L, unsigned long long *__restrict__ kout,
                                           int *__restrict__ vout, int tile, unsigned long long *sk, int *sv,
                                           int *sa) {
  const int out0 = tile * TILE;
  const int pair_base = (out0 / (2 * L)) * (2 * L);
  const int na = min(L, n - pair_base), nb = max(0, min(L, n - pair_base - L));
  const unsigned long long *ak = kin + pair_base, *bk = kin + pair_base + L;
  const int *av = vin + pair_base, *bv = vin + pair_base + L;
  const int d0 = out0 - pair_base, d1 = min(d0 + TILE, na + nb);
  if (threadIdx.x < 64) {  // warp 0 -> start diagonal, warp 1 -> end diagonal
    const int wsel = threadIdx.x >> 5;
    const int r = merge_path_warp(ak, av, na, bk, bv, nb, wsel ? d1 : d0);
    if ((threadIdx.x & 31) == 0) sa[wsel] = r;
  }
  __syncthreads();
  const int a0 = sa[0], a1 = sa[1], b0 = d0 - a0, b1 = d1 - a1;
  const int la = a1 - a0, lb = b1 - b0;  // la + lb = d1 - d0 <= TILE
  for (int e = threadIdx.x; e < la; e += 256) sk[e] = ak[a0 + e], sv[e] = av[a0 + e];
  for (int e = threadIdx.x; e < lb; e += 256) sk[la + e] = bk[b0 + e], sv[la + e] = bv[b0 + e];
  __syncthreads();
  constexpr int VT = TILE / 256;
  const int diag = min(threadIdx.x * VT, la + lb);
  int ia = merge_path(sk, sv, la, sk + la, sv + la, lb, diag), ib = diag - ia;
#pragma unroll
  for (int q = 0; q < VT; ++q) {
    const int o = diag + q;
    if (o >= la + lb) break;
    bool takeA;
    if (ia >= la) takeA = false;
    else if (ib >= lb) takeA = true;
    else takeA = !lt(sk[la + ib], sv[la + ib], sk[ia], sv[ia]);
    const int src = takeA ? ia : la + ib;
    kout[out0 + o] = sk[src];
    vout[out0 + o] = sv[src];
    ia += takeA, ib += !takeA;
  }
}

__device__ __forceinline__ double cost_of(unsigned long long b) { return key_cost(b); }
// maximum(abs.(diff(...))) propagates NaN in Julia (then `NaN < 10e-3` is false: no break): fmax alone would drop it
__device__ __forceinline__ double gap_max(double mx, double d) { return (d != d || mx != mx) ? d + mx : fmax(mx, d); }

// All merge passes in ONE cooperative kernel (grid-wide barrier between passes) followed by the elite
// early-stop test maximum(abs.(diff(elite_traj_cost))) < 10e-3 (POL:458-461, 566-569) on the sorted keys:
// K = 65 536 needs 5 passes, which as separate launches cost 5 x 12 µs of mostly launch/drain latency.
__global__ void __launch_bounds__(256) merge_all_kernel(unsigned long long *kin, int *vin, unsigned long long *kout,
                                                         int *vout, int n, int m, int early_stop, int *stop_flag,
                                                         const int *stop) {
  if (stop && *stop) return;  // grid-uniform
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned long long sk[TILE];
  __shared__ int sv[TILE];
  __shared__ int sa[2];
  __shared__ double red[8];
  const int ntiles = (n + TILE - 1) / TILE;
  for (long long L = TILE; L < n; L <<= 1) {
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      __syncthreads();
      merge_tile(kin, vin, n, (int)L, kout, vout, t, sk, sv, sa);
    }
    grid.sync();
    unsigned long long *tk = kin;
    kin = kout, kout = tk;
    int *tv = vin;
    vin = vout, vout = tv;
  }
  if (blockIdx.x == 0 && stop_flag && m > 1) {
    double mx = -1.0;
    for (int j = threadIdx.x; j + 1 < m; j += 256) mx = gap_max(mx, fabs(cost_of(kin[j + 1]) - cost_of(kin[j])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = gap_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) mx = gap_max(mx, red[w]);
      if (early_stop && mx < 10e-3) *stop_flag = 1;
    }
  }
}

}  // namespace

// K <= TILE (the reference's own sizes, K = 150 / 375 / 20): ONE ordinary launch — a bitonic network over the next
// power of two >= K instead of the full 2048-entry tile, with the elite early-stop test at its end — replaces the
// tile sort + the cooperative merge kernel (whose only job would be that test).
__global__ void __launch_bounds__(1024) small_sort_kernel(const double *__restrict__ costs, int n, int tn,
                                                           unsigned long long *__restrict__ keys,
                                                           int *__restrict__ vals, int m, int early_stop,
                                                           int *stop_flag, const int *stop) {
  if (stop && *stop) return;
  __shared__ unsigned long long sk[TILE];
  __shared__ int sv[TILE];
  __shared__ double red[32];
  for (int e = threadIdx.x; e < tn; e += 1024) {
    sk[e] = e < n ? key_of(costs[e]) : KEY_PAD;
    sv[e] = e < n ? e : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= tn; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      if (t < (tn >> 1)) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair, bit j clear
        const int p = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long ka = sk[i], kb = sk[p];
        const int va = sv[i], vb = sv[p];
        if (lt(kb, vb, ka, va) == up) sk[i] = kb, sv[i] = vb, sk[p] = ka, sv[p] = va;
      }
      __syncthreads();
    }
  for (int e = threadIdx.x; e < n; e += 1024) keys[e] = sk[e], vals[e] = sv[e];
  if (stop_flag && m > 1) {  // maximum(abs.(diff(elite_traj_cost))) < 10e-3, POL:458-461, 566-569
    double mx = -1.0;
    for (int j = threadIdx.x; j + 1 < m; j += 1024) mx = gap_max(mx, fabs(cost_of(sk[j + 1]) - cost_of(sk[j])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = gap_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 32; ++w) mx = gap_max(mx, red[w]);
      if (early_stop && mx < 10e-3) *stop_flag = 1;
    }
  }
}

int sort_launches(int K) { return K <= TILE ? 1 : 2; }

// Sorts costs[0:K]; on return `order` holds the stable ascending permutation (0-based sample ids).
// keys_a/keys_b, vals_b: scratch of K entries. With stop_flag != nullptr the kernel also evaluates the elite
// early-stop test on the m smallest costs and raises *stop_flag. max_ctas: co-resident CTA budget of the
// cooperative merge kernel. Returns a cudaError_t.
int launch_sortperm(const double *costs, int K, unsigned long long *keys_a, unsigned long long *keys_b, int *order,
                    int *vals_b, int m, int early_stop, int *stop_flag, const int *stop, int max_ctas,
                    cudaStream_t s) {
  if (K <= TILE) {
    int tn = 32;
    while (tn < K) tn <<= 1;
    small_sort_kernel<<<1, 1024, 0, s>>>(costs, K, tn, keys_a, order, m, early_stop, stop_flag, stop);
    return (int)cudaGetLastError();
  }
  const int ntiles = (K + TILE - 1) / TILE;
  int passes = 0;
  for (long long L = TILE; L < K; L <<= 1) ++passes;
  // choose the starting buffer so that the final pass lands in (keys_a, order)
  unsigned long long *kin = (passes & 1) ? keys_b : keys_a, *kout = (passes & 1) ? keys_a : keys_b;
  int *vin = (passes & 1) ? vals_b : order, *vout = (passes & 1) ? order : vals_b;
  tile_sort_kernel<<<ntiles, 1024, 0, s>>>(costs, K, kin, vin, stop);
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  void *args[] = {(void *)&kin, (void *)&vin, (void *)&kout, (void *)&vout, (void *)&K,
                  (void *)&m,   (void *)&early_stop, (void *)&stop_flag, (void *)&stop};
  return (int)cudaLaunchCooperativeKernel((const void *)merge_all_kernel, dim3(grid), dim3(256), args, 0, s);
}

int sort_max_ctas(int num_sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_all_kernel, 256, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return per_sm * num_sms;
}

}  // namespace mpopis
