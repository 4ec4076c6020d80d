// sort.cu — G4: stable ascending arg-sort of the K trajectory costs.
//
// Replaces `order = sortperm(trajectory_cost)` (POL:455, 563). Julia's sortperm is a stable merge
// sort on isless (ties keep index order, −0.0 < +0.0, NaN last). An LSD radix sort over the
// order-preserving 64-bit image of the doubles is stable and induces exactly that total order, so
// given identical costs the permutation is bit-identical to the reference's (tests/test_parity_*).
// 8 passes of 8 bits: per pass a per-block digit histogram, an exclusive scan over (digit, block)
// and a stable scatter — written here rather than calling a library sort.
#include "engine.cuh"

namespace mpopis {

__device__ __forceinline__ unsigned long long key_of(double c) {
  unsigned long long b = (unsigned long long)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

constexpr int RS_BLOCK = 256;  // threads per block
constexpr int RS_ITEMS = 8;    // keys per thread (contiguous per thread -> stable ranking)
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;

__global__ void sort_init_kernel(const double *__restrict__ costs, int K, unsigned long long *__restrict__ keys,
                                 int *__restrict__ vals, const int *stop) {
  if (stop && *stop) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) keys[i] = key_of(costs[i]), vals[i] = i;
}

// hist[d * nblocks + b] = number of keys of tile b whose digit is d
__global__ void __launch_bounds__(RS_BLOCK) radix_hist_kernel(const unsigned long long *__restrict__ keys, int K,
                                                               int shift, int nblocks, int *__restrict__ hist,
                                                               const int *stop) {
  if (stop && *stop) return;
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
  for (int q = 0; q < RS_ITEMS; ++q) {
    const int i = base + q * RS_BLOCK + threadIdx.x;
    if (i < K) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist (256 * nblocks ints), single CTA
__global__ void __launch_bounds__(1024) radix_scan_kernel(int *__restrict__ hist, int n, const int *stop) {
  if (stop && *stop) return;
  __shared__ int seg[1024];
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(n, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += hist[i];
  seg[threadIdx.x] = s;
  __syncthreads();
  // Hillis–Steele inclusive scan over the 1024 segment sums
  for (int off = 1; off < 1024; off <<= 1) {
    int v = threadIdx.x >= off ? seg[threadIdx.x - off] : 0;
    __syncthreads();
    seg[threadIdx.x] += v;
    __syncthreads();
  }
  int run = seg[threadIdx.x] - s;
  for (int i = beg; i < end; ++i) {
    const int v = hist[i];
    hist[i] = run;
    run += v;
  }
}

// stable scatter: within a tile, keys with equal digit keep their input order. One warp-serial
// ranking per digit would be slow; instead each thread ranks its keys with a per-digit running
// counter built from a ballot-free two-level count: (1) per-thread-chunk digit counts in shared
// memory laid out [thread][...] are too large, so the tile is processed in RS_ITEMS rounds of
// RS_BLOCK consecutive keys; in each round a key's rank among equal digits is the number of
// lower-indexed threads holding the same digit (match_any + popc per warp, plus per-warp digit
// offsets accumulated in shared memory).
__global__ void __launch_bounds__(RS_BLOCK) radix_scatter_kernel(const unsigned long long *__restrict__ keys_in,
                                                                  const int *__restrict__ vals_in, int K, int shift,
                                                                  int nblocks, const int *__restrict__ hist,
                                                                  unsigned long long *__restrict__ keys_out,
                                                                  int *__restrict__ vals_out, const int *stop) {
  if (stop && *stop) return;
  __shared__ int digit_base[256];          // running output offset per digit for this tile
  __shared__ int warp_cnt[RS_BLOCK / 32][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  digit_base[threadIdx.x] = hist[threadIdx.x * nblocks + blockIdx.x];
  const int base = blockIdx.x * RS_TILE;
  for (int q = 0; q < RS_ITEMS; ++q) {
    for (int d = lane; d < 256; d += 32) warp_cnt[wid][d] = 0;
    __syncthreads();
    const int i = base + q * RS_BLOCK + threadIdx.x;
    const bool ok = i < K;
    unsigned long long key = 0;
    int val = 0, d = 0;
    if (ok) key = keys_in[i], val = vals_in[i], d = (int)((key >> shift) & 255);
    // rank within the warp among lanes with the same digit
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    unsigned same = __match_any_sync(0xffffffffu, ok ? d : (256 + lane));
    same &= act;
    const int rank_in_warp = __popc(same & ((1u << lane) - 1));
    if (ok && rank_in_warp == 0) warp_cnt[wid][d] = __popc(same);
    __syncthreads();
    if (ok) {
      int off = digit_base[d];
      for (int w2 = 0; w2 < wid; ++w2) off += warp_cnt[w2][d];
      const int dst = off + rank_in_warp;
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
    __syncthreads();
    {  // advance the per-digit base by this round's totals
      const int dd = threadIdx.x;
      int tot = 0;
#pragma unroll
      for (int w2 = 0; w2 < RS_BLOCK / 32; ++w2) tot += warp_cnt[w2][dd];
      digit_base[dd] += tot;
    }
    __syncthreads();
  }
}

__global__ void sorted_costs_kernel(const unsigned long long *__restrict__ keys, int m, double *__restrict__ out,
                                    const int *stop) {
  if (stop && *stop) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  unsigned long long b = keys[i];
  b = (b >> 63) ? (b & 0x7fffffffffffffffULL) : ~b;
  out[i] = __longlong_as_double((long long)b);
}

int sort_nblocks(int K) { return (K + RS_TILE - 1) / RS_TILE; }
size_t sort_hist_ints(int K) { return (size_t)256 * sort_nblocks(K); }

// Sorts costs[0:K]; on return `order` holds the stable ascending permutation (0-based sample ids)
// and sorted_costs[0:m] the m smallest costs in order. keys_a/keys_b, vals_b: scratch of K entries.
void launch_sortperm(const double *costs, int K, int m, unsigned long long *keys_a, unsigned long long *keys_b,
                     int *order, int *vals_b, int *hist, double *sorted_costs, const int *stop, cudaStream_t s) {
  const int nb = sort_nblocks(K);
  sort_init_kernel<<<(K + 255) / 256, 256, 0, s>>>(costs, K, keys_a, order, stop);
  unsigned long long *kin = keys_a, *kout = keys_b;
  int *vin = order, *vout = vals_b;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 8 * pass;
    radix_hist_kernel<<<nb, RS_BLOCK, 0, s>>>(kin, K, shift, nb, hist, stop);
    radix_scan_kernel<<<1, 1024, 0, s>>>(hist, 256 * nb, stop);
    radix_scatter_kernel<<<nb, RS_BLOCK, 0, s>>>(kin, vin, K, shift, nb, hist, kout, vout, stop);
    unsigned long long *tk = kin;
    kin = kout, kout = tk;
    int *tv = vin;
    vin = vout, vout = tv;
  }
  // 8 passes = even number of swaps: results are back in keys_a / order
  sorted_costs_kernel<<<(m + 255) / 256, 256, 0, s>>>(kin, m, sorted_costs, stop);
}

}  // namespace mpopis
