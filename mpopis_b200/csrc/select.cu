// select.cu — G4 for :cemppi without a sort: elite SELECTION + the early-stop test + the elite gather.
//
// Replaces, for the cross-entropy policy (POL:455-461),
//     order = sortperm(trajectory_cost); elite = E[:, order[1:m_elite]]
//     if maximum(abs.(diff(trajectory_cost[order[1:m_elite]]))) < 10e-3; break; end
// The CE update (POL:464-465) uses the elite SET only — mean and covariance do not depend on the order of the
// columns — so a full stable sort of K costs (round 1: tile sort + merge passes, 86 µs per AIS iteration at
// K = 65 536, and in the sharded case an all-gather of every rank's sorted run + chains of binary searches) is
// replaced by
//   1. an exact radix SELECT of the m-th smallest composite key (cost image under Base.isless, global sample
//      index): 11-bit digits, most significant first, candidates narrowed per pass until <= 256 remain, then
//      ranked directly. The elite set {key <= τ} is exactly order[1:m] (ties by index, like the stable sort);
//   2. the early-stop test WITHOUT sorting the elites: with c₁ the smallest and c_m the largest elite cost,
//        - a NaN or non-finite elite cost makes the reference's maximum NaN/Inf                 -> no stop;
//        - (c_m − c₁)·200 >= 2m + 2 forces a gap >= 10e-3 among m values (pigeonhole)           -> no stop;
//        - otherwise the elites fall into <= 2m + 2 buckets of width 0.005 (two costs in one bucket differ by
//          < 0.005, so only gaps BETWEEN consecutive non-empty buckets can reach 10e-3); per bucket the min and
//          max key are kept (64-bit atomics) and the gaps min(next) − max(prev) are evaluated with the
//          reference's own subtraction. O(m), exact in the `< 10e-3` decision;
//   3. the ids of the elites owned by this shard (global ids k0 .. k0+Kloc), compacted in index order.
// Everything runs in ONE cooperative kernel on the (all-gathered) cost vector; in the sharded configuration every
// rank executes it redundantly on identical input, so all ranks agree on τ and on the stop decision bit for bit
// with no collective beyond the all-gather of the costs (8 B per sample).
// A second kernel gathers the local elite columns into X and accumulates Σx, Σx² per row on the way.
#include <cooperative_groups.h>
#include <math_constants.h>

#include <cstdlib>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace mpopis {

namespace {

constexpr int NBIN = 2048;       // 11-bit digits
constexpr int NCAND = 256;       // direct ranking below this many candidates
constexpr int MAXCTA = SELECT_MAX_CTAS;
constexpr unsigned long long KMAX = ~0ULL;

struct Ws {  // device workspace; zero/clean between launches (the kernel restores it before it ends)
  unsigned hist[3][NBIN];
  unsigned long long min_key;  // KMAX when clean
  unsigned ncand;              // fill of cand_*
  unsigned pad0;
  unsigned long long cand_k[NCAND];
  unsigned cand_i[NCAND];
  int cta_cnt[MAXCTA];                     // local elites per CTA slice
  unsigned long long part_first[MAXCTA];   // bucket scan partials: key of the first non-empty bucket's min
  unsigned long long part_last[MAXCTA];    //                      key of the last non-empty bucket's max
  int part_flags[MAXCTA];                  // bit 0: any non-empty bucket, bit 1: a gap >= 10e-3 seen
  unsigned long long cta_min1[MAXCTA], cta_min2[MAXCTA];  // the two smallest keys of each CTA's slice (pass 0)
};
static_assert(sizeof(Ws) <= SELECT_WS_BYTES, "SELECT_WS_BYTES too small");

// 96-bit composite value = (key << 32) | idx
__device__ __forceinline__ unsigned digit96(unsigned long long k, unsigned i, int shift, unsigned mask) {
  if (shift >= 32) return (unsigned)(k >> (shift - 32)) & mask;
  return (unsigned)(((k << 32) | i) >> shift) & mask;
}
// do (k, i) and the prefix (pk, pi) agree on all bits >= low?
__device__ __forceinline__ bool match96(unsigned long long k, unsigned i, unsigned long long pk, unsigned pi, int low) {
  if (low >= 96) return true;
  if (low >= 32) return ((k ^ pk) >> (low - 32)) == 0;
  return k == pk && ((i ^ pi) >> low) == 0;
}
__device__ __forceinline__ bool le96(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
  return ka < kb || (ka == kb && ia <= ib);
}

// histogram increment aggregated over the warp: lanes with the same digit elect one lane that adds their count. The
// most significant digits of real cost vectors are nearly constant (sign, exponent), so un-aggregated every lane of every
// warp would serialise on ONE shared-memory word. All 32 lanes must call this (inactive ones with valid = false).
__device__ __forceinline__ void hist_add(unsigned *hist, unsigned digit, bool valid) {
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? digit : 0xffffffffu);
  if (valid && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[digit], (unsigned)__popc(peers));
}

struct Seg {  // summary of a run of buckets
  unsigned long long first, last;
  int flags;  // bit 0 any, bit 1 gap >= 10e-3
};
__device__ __forceinline__ Seg seg_join(const Seg &a, const Seg &b) {
  if (!(a.flags & 1)) return b;
  if (!(b.flags & 1)) return a;
  Seg r;
  r.first = a.first, r.last = b.last;
  r.flags = a.flags | b.flags;
  // the reference's own arithmetic: abs(c[j+1] − c[j]) < 10e-3 (POL:459)
  if (!(fabs(key_cost(b.first) - key_cost(a.last)) < 10e-3)) r.flags |= 2;
  return r;
}

template <int ST>
__global__ void __launch_bounds__(ST) ce_select_kernel(const double *__restrict__ costs, int Ktot, int m, long long k0,
                                                        int Kloc, int early_stop, Ws *ws,
                                                        unsigned long long *__restrict__ bmin,
                                                        unsigned long long *__restrict__ bmax, long long nb_cap,
                                                        int *__restrict__ eidx, int *m_loc, double *tau_out,
                                                        int *stop_flag, const int *stop) {
  if (stop && *stop) return;  // grid-uniform: nobody writes the flag before the last grid barrier
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned sh[NBIN];
  __shared__ unsigned long long s64[ST / 32 * 2 + 4];  // per-warp pairs; [0] doubles as the τ hand-over
  __shared__ int s32[ST / 32 + 8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nc = gridDim.x, b = blockIdx.x;
  const int S = (Ktot + nc - 1) / nc;
  const int ibeg = min(Ktot, b * S), iend = min(Ktot, ibeg + S);

  // ---- 1. radix select of rank m-1 over (key, idx) ---------------------------------------------------------
  unsigned long long pk = 0, tk = 0;
  unsigned pi = 0, ti = 0;
  long long need = m - 1;
  int low = 96;  // bits >= low of the prefix are fixed
  unsigned long long mn = KMAX, mn2 = KMAX;  // the two smallest keys this thread has seen (mn <= mn2, duplicates kept)
  bool have_tau = false;
  for (int pass = 0; pass < 9; ++pass) {
    const int shift = pass < 8 ? 85 - 11 * pass : 0;
    const unsigned mask = pass < 8 ? 0x7FFu : 0xFFu;
    for (int e = tid; e < NBIN; e += ST) sh[e] = 0;
    __syncthreads();
    for (int i0 = ibeg; i0 < iend; i0 += ST) {
      const int i = i0 + tid;
      const unsigned long long k = i < iend ? cost_key(costs[i]) : KMAX;
      if (pass == 0) {
        mn2 = k < mn ? mn : (k < mn2 ? k : mn2);
        mn = k < mn ? k : mn;
      }
      hist_add(sh, digit96(k, (unsigned)i, shift, mask), i < iend && match96(k, (unsigned)i, pk, pi, low));
    }
    __syncthreads();
    unsigned *gh = ws->hist[pass % 3];
    for (int e = tid; e < NBIN; e += ST)
      if (sh[e]) atomicAdd(&gh[e], sh[e]);
    if (b == 0) {  // clean the histogram of the next pass (last read two barriers ago)
      unsigned *nh = ws->hist[(pass + 1) % 3];
      for (int e = tid; e < NBIN; e += ST) nh[e] = 0;
    }
    if (pass == 0) {  // the two smallest keys: the global minimum (smallest elite cost) and its successor
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a1 = __shfl_xor_sync(0xffffffffu, mn, o), a2 = __shfl_xor_sync(0xffffffffu, mn2, o);
        const unsigned long long lo = a1 < mn ? a1 : mn, hi = a1 < mn ? mn : a1;  // merge the pairs (mn, mn2), (a1, a2)
        const unsigned long long c2 = a2 < mn2 ? a2 : mn2;
        mn = lo, mn2 = hi < c2 ? hi : c2;
      }
      if (lane == 0) s64[2 * wid] = mn, s64[2 * wid + 1] = mn2;
      __syncthreads();
      if (tid == 0) {
        unsigned long long m1 = KMAX, m2 = KMAX;
        for (int w = 0; w < ST / 32; ++w) {
          const unsigned long long a1 = s64[2 * w], a2 = s64[2 * w + 1];
          const unsigned long long lo = a1 < m1 ? a1 : m1, hi = a1 < m1 ? m1 : a1, c2 = a2 < m2 ? a2 : m2;
          m1 = lo, m2 = hi < c2 ? hi : c2;
        }
        ws->cta_min1[b] = m1, ws->cta_min2[b] = m2;
        if (m1 != KMAX) atomicMin(&ws->min_key, m1);
      }
    }
    grid.sync();
    // every CTA finds the bin of rank `need` in the global histogram (redundantly: no broadcast needed)
    unsigned loc[NBIN / ST], tot = 0;
#pragma unroll
    for (int q = 0; q < NBIN / ST; ++q) loc[q] = __ldcg(&gh[tid * (NBIN / ST) + q]), tot += loc[q];
    unsigned incl = tot;  // inclusive scan over the 256 thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s32[wid] = (int)incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += (unsigned)s32[w];
    const unsigned excl = wbase + incl - tot;
    if ((long long)excl <= need && need < (long long)(excl + tot)) {  // exactly one thread
      unsigned run = excl;
      int q = 0;
      for (; q < NBIN / ST - 1; ++q) {
        if (need < (long long)(run + loc[q])) break;
        run += loc[q];
      }
      s32[ST / 32 + 0] = tid * (NBIN / ST) + q;  // selected bin
      s32[ST / 32 + 1] = (int)run;               // candidates below it
      s32[ST / 32 + 2] = (int)loc[q];            // candidates inside it
    }
    __syncthreads();
    const unsigned bin = (unsigned)s32[ST / 32 + 0];
    need -= s32[ST / 32 + 1];
    const unsigned ncand = (unsigned)s32[ST / 32 + 2];
    if (shift >= 32) pk |= (unsigned long long)bin << (shift - 32);
    else {
      const unsigned long long v = (unsigned long long)bin << shift;  // may straddle key / index
      pk |= v >> 32, pi |= (unsigned)v;
    }
    low = shift;
    __syncthreads();
    if (ncand <= NCAND || pass == 8) {
      if (ncand == 1 && pass == 8) {
        tk = pk, ti = pi, have_tau = true;
      }
      break;
    }
  }
  if (!have_tau) {
    // collect the <= 256 remaining candidates and rank them directly
    for (int i = ibeg + tid; i < iend; i += ST) {
      const unsigned long long k = cost_key(costs[i]);
      if (match96(k, (unsigned)i, pk, pi, low)) {
        const unsigned slot = atomicAdd(&ws->ncand, 1u);
        if (slot < NCAND) ws->cand_k[slot] = k, ws->cand_i[slot] = (unsigned)i;
      }
    }
    grid.sync();
    const unsigned n = min(__ldcg(&ws->ncand), (unsigned)NCAND);
    unsigned long long *ck = reinterpret_cast<unsigned long long *>(sh);  // 256 x 8 B
    unsigned *ci = sh + 2 * NCAND;                                       // 256 x 4 B (sh holds 2048 words)
    if (tid < (int)n) ck[tid] = __ldcg(&ws->cand_k[tid]), ci[tid] = __ldcg(&ws->cand_i[tid]);
    __syncthreads();
    if (tid < (int)n) {
      const unsigned long long k = ck[tid];
      const unsigned i = ci[tid];
      int below = 0;
      for (unsigned j = 0; j < n; ++j) below += (ck[j] < k || (ck[j] == k && ci[j] < i)) ? 1 : 0;
      if (below == (int)need) s64[0] = k, s32[ST / 32 + 3] = (int)i;
    }
    __syncthreads();
    tk = s64[0], ti = (unsigned)s32[ST / 32 + 3];
    __syncthreads();
  }

  // ---- 2. mark: bucket min/max of all elites, count of the elites this shard owns ------------------------------
  const unsigned long long mnk = __ldcg(&ws->min_key);
  const double c1 = key_cost(mnk), cm = key_cost(tk);
  // The two smallest costs are the first two elites (m >= 2). The best samples of a cost vector are sparse, so their gap
  // alone usually exceeds 10e-3 and decides "no stop" — then no bucket is touched. Every CTA merges the per-CTA pairs.
  unsigned long long g1 = KMAX, g2 = KMAX;
  for (int q = tid; q < nc; q += ST) {
    const unsigned long long a1 = __ldcg(&ws->cta_min1[q]), a2 = __ldcg(&ws->cta_min2[q]);
    const unsigned long long lo = a1 < g1 ? a1 : g1, hi = a1 < g1 ? g1 : a1, c2 = a2 < g2 ? a2 : g2;
    g1 = lo, g2 = hi < c2 ? hi : c2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a1 = __shfl_xor_sync(0xffffffffu, g1, o), a2 = __shfl_xor_sync(0xffffffffu, g2, o);
    const unsigned long long lo = a1 < g1 ? a1 : g1, hi = a1 < g1 ? g1 : a1, c2 = a2 < g2 ? a2 : g2;
    g1 = lo, g2 = hi < c2 ? hi : c2;
  }
  __syncthreads();
  if (lane == 0) s64[2 * wid] = g1, s64[2 * wid + 1] = g2;
  __syncthreads();
  g1 = KMAX, g2 = KMAX;
  for (int w = 0; w < ST / 32; ++w) {
    const unsigned long long a1 = s64[2 * w], a2 = s64[2 * w + 1];
    const unsigned long long lo = a1 < g1 ? a1 : g1, hi = a1 < g1 ? g1 : a1, c2 = a2 < g2 ? a2 : g2;
    g1 = lo, g2 = hi < c2 ? hi : c2;
  }
  const bool first_gap_decides = m > 1 && g2 != KMAX && !(fabs(key_cost(g2) - key_cost(g1)) < 10e-3);  // POL:459 on c₂ − c₁
  long long nb = 0;  // 0: the stop test is decided without buckets (no stop)
  if (early_stop && m > 1 && !first_gap_decides && tk < COST_KEY_NAN && isfinite(c1) && isfinite(cm)) {
    const double span = (cm - c1) * 200.0;
    if (span < (double)(2LL * m + 2) && span + 1.0 <= (double)nb_cap) nb = (long long)span + 1;
  }
  int cnt = 0;
  for (int i0 = ibeg; i0 < iend; i0 += ST) {
    const int i = i0 + tid;
    bool elite = false;
    unsigned long long k = 0;
    if (i < iend) {
      k = cost_key(costs[i]);
      elite = le96(k, (unsigned)i, tk, ti);
    }
    if (elite && i >= k0 && i < k0 + Kloc) ++cnt;
    if (nb > 0) {
      long long bk = -1;
      if (elite) {
        bk = (long long)((key_cost(k) - c1) * 200.0);
        bk = bk < 0 ? 0 : (bk >= nb ? nb - 1 : bk);
      }
      // a warp whose lanes all hit one bucket (the converged case the test exists for) issues one atomic pair
      const long long b0 = __shfl_sync(0xffffffffu, bk, 0);
      if (__all_sync(0xffffffffu, bk == b0)) {
        if (b0 >= 0) {
          unsigned long long lo = k, hi = k;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), c = __shfl_xor_sync(0xffffffffu, hi, o);
            lo = a < lo ? a : lo, hi = c > hi ? c : hi;
          }
          if (lane == 0) atomicMin(&bmin[b0], lo), atomicMax(&bmax[b0], hi);
        }
      } else if (bk >= 0) {
        atomicMin(&bmin[bk], k), atomicMax(&bmax[bk], k);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s32[wid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < ST / 32; ++w) t += s32[w];
    ws->cta_cnt[b] = t;
  }
  grid.sync();

  // ---- 3a. compaction: ids of the local elites in index order ---------------------------------------------------
  {
    int off = 0;
    for (int q = tid; q < b; q += ST) off += __ldcg(&ws->cta_cnt[q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
    __syncthreads();
    if (lane == 0) s32[wid] = off;
    __syncthreads();
    off = 0;
    for (int w = 0; w < ST / 32; ++w) off += s32[w];
    __syncthreads();
    if (b == nc - 1 && tid == 0) *m_loc = off + __ldcg(&ws->cta_cnt[b]);
    for (int i0 = ibeg; i0 < iend; i0 += ST) {
      const int i = i0 + tid;
      bool own = false;
      if (i < iend && i >= k0 && i < k0 + Kloc) own = le96(cost_key(costs[i]), (unsigned)i, tk, ti);
      const unsigned bal = __ballot_sync(0xffffffffu, own);
      if (lane == 0) s32[wid] = __popc(bal);
      __syncthreads();
      int base = off, tot = 0;
      for (int w = 0; w < ST / 32; ++w) {
        if (w < wid) base += s32[w];
        tot += s32[w];
      }
      if (own) eidx[base + __popc(bal & ((1u << lane) - 1u))] = (int)(i - k0);
      off += tot;
      __syncthreads();
    }
  }
  // ---- 3b. bucket scan: gaps between consecutive non-empty buckets; buckets are left clean ----------------------
  if (nb > 0) {
    const long long Q = (nb + nc - 1) / nc;
    const long long q0 = min(nb, (long long)b * Q), q1 = min(nb, q0 + Q);
    const long long per = (q1 - q0 + ST - 1) / ST;
    const long long t0 = min(q1, q0 + tid * per), t1 = min(q1, t0 + per);
    Seg sg{0, 0, 0};
    for (long long q = t0; q < t1; ++q) {
      const unsigned long long lo = __ldcg(&bmin[q]), hi = __ldcg(&bmax[q]);
      if (lo != KMAX) {
        bmin[q] = KMAX, bmax[q] = 0;
        sg = seg_join(sg, Seg{lo, hi, 1});
      }
    }
    // ordered combine: lanes of a warp, then the warps
    __shared__ Seg segs[ST];
    segs[tid] = sg;
    __syncthreads();
    if (lane == 0) {
      Seg acc = segs[tid];
      for (int l = 1; l < 32; ++l) acc = seg_join(acc, segs[tid + l]);
      segs[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      Seg acc = segs[0];
      for (int w = 1; w < ST / 32; ++w) acc = seg_join(acc, segs[32 * w]);
      ws->part_first[b] = acc.first, ws->part_last[b] = acc.last, ws->part_flags[b] = acc.flags;
    }
  }
  // restore the workspace for the next launch (nobody reads these any more)
  if (b == 0) {
    for (int e = tid; e < 3 * NBIN; e += ST) (&ws->hist[0][0])[e] = 0;
    if (tid == 0) ws->ncand = 0, ws->min_key = KMAX;
  }
  if (tau_out && b == 0 && tid == 0) {
    tau_out[0] = key_cost(tk), tau_out[1] = (double)ti, tau_out[2] = c1, tau_out[3] = (double)nb;
  }
  if (nb == 0) return;  // grid-uniform
  grid.sync();
  if (b == 0 && tid == 0) {
    Seg acc{0, 0, 0};
    for (int q = 0; q < nc; ++q)
      acc = seg_join(acc, Seg{__ldcg(&ws->part_first[q]), __ldcg(&ws->part_last[q]), __ldcg(&ws->part_flags[q])});
    if ((acc.flags & 1) && !(acc.flags & 2)) *stop_flag = 1;  // maximum(abs.(diff(elite costs))) < 10e-3
  }
}

// (Round 2 also built the same selection inside ONE thread-block cluster of 8 CTAs x 1024 threads — cluster.sync()
// barriers, histograms and candidates in distributed shared memory. ncu at K = 65 536: 60 µs against 36 µs for the
// cooperative kernel above: 8 SMs pull the keys and walk the buckets where 64 do; the grid barriers were never the cost,
// the dependent L2 round trips of a latency-bound kernel are. It was removed: profiles/README.md, round 2.)

__global__ void select_ws_init_kernel(Ws *ws, unsigned long long *bmin, unsigned long long *bmax, long long nb_cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb_cap) bmin[i] = KMAX, bmax[i] = 0;
  if (i < 3 * NBIN) (&ws->hist[0][0])[i] = 0;
  if (i == 0) ws->ncand = 0, ws->min_key = KMAX;
}

// X[r][j] = E[r][eidx[j]] for j < *m_loc (coalesced stores, 8-byte gathers out of an L2-resident E), and the
// per-chunk partial sums Σ_j X[r][j], Σ_j X[r][j]² (block-reduced, written per (chunk, row): deterministic).
// grid (chunks of GS_COLS columns, cs rows); chunks beyond *m_loc write zero partials.
constexpr int GS_COLS = 1024;
__global__ void __launch_bounds__(256) elite_gather_sums_kernel(const double *__restrict__ E, long long ldk, int cs,
                                                                 const int *__restrict__ eidx, const int *m_loc,
                                                                 double *__restrict__ X, long long ldx,
                                                                 double *__restrict__ partial, const int *stop) {
  if (stop && *stop) return;
  __shared__ double red[2][8];
  const int m = *m_loc, r = blockIdx.y, c = blockIdx.x;
  const int j0 = c * GS_COLS;
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int q = 0; q < GS_COLS / 256; ++q) {
    const int j = j0 + q * 256 + threadIdx.x;
    if (j < m) {
      const double x = E[(size_t)r * ldk + eidx[j]];
      X[(size_t)r * ldx + j] = x;
      s1 += x, s2 = fma(x, x, s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1, red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    partial[((size_t)c * 2 + threadIdx.x) * cs + r] = t;
  }
}

// sums = [Σx (cs) | n | Σx² (cs)] from the chunk partials; with `finalize` (single GPU, or after the all-reduce in the
// sharded case with nchunks = 0) also μ = Σx/n, pol.U += μ (POL:465) and 1/σ_i for the :ss standardisation
// (σ_i² = Σx_i²/n − μ_i²: the elites' noise mean is a fraction of their spread, so the cancellation costs a few ulps at
// most, and it only feeds the shrinkage intensity λ̂). One warp per row: the lanes fetch the chunk partials in parallel
// (a per-thread loop over the chunks was a chain of dependent L2 round trips, 12 µs for 13 chunks) and combine them in
// a fixed shuffle tree, so the sums are deterministic.
__global__ void __launch_bounds__(256) ce_sums_kernel(const double *__restrict__ partial, int nchunks, int cs,
                                                       const int *m_loc, double *__restrict__ sums, int finalize,
                                                       int standardise, double *__restrict__ mu, double *__restrict__ U,
                                                       double *__restrict__ dinv, const int *stop) {
  if (stop && *stop) return;
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r > cs) return;  // warp-uniform
  double s1 = 0.0, s2 = 0.0;
  if (nchunks > 0 && r < cs) {
    for (int c = lane; c < nchunks; c += 32) s1 += partial[((size_t)c * 2) * cs + r], s2 += partial[((size_t)c * 2 + 1) * cs + r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane != 0) return;
  if (nchunks > 0) {
    if (r == cs) sums[cs] = (double)*m_loc;
    else sums[r] = s1, sums[cs + 1 + r] = s2;
  } else if (r < cs) {
    s1 = sums[r], s2 = sums[cs + 1 + r];
  }
  if (!finalize || r == cs) return;
  const double n = nchunks > 0 ? (double)*m_loc : sums[cs];
  const double m1 = s1 / n;
  mu[r] = m1;
  U[r] = U[r] + m1;
  dinv[r] = standardise ? 1.0 / sqrt(s2 / n - m1 * m1) : 1.0;
}

}  // namespace

int select_max_ctas(int num_sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ce_select_kernel<256>, 256, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int c = per_sm * num_sms;
  return c < MAXCTA ? c : MAXCTA;
}

long long select_bucket_capacity(int m) { return 2LL * m + 4; }

void launch_select_init(void *ws, unsigned long long *bmin, unsigned long long *bmax, long long nb_cap, cudaStream_t s) {
  const long long n = nb_cap > 3 * NBIN ? nb_cap : 3 * NBIN;
  select_ws_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((Ws *)ws, bmin, bmax, nb_cap);
}

int launch_ce_select(const double *costs, int Ktot, int m, long long k0, int Kloc, int early_stop, void *ws,
                     unsigned long long *bmin, unsigned long long *bmax, long long nb_cap, int *eidx, int *m_loc,
                     double *tau_out, int *stop_flag, const int *stop, int max_ctas, cudaStream_t s) {
  // Small K: 256-thread CTAs of 1024 keys. Large K (sharded policies select on the GATHERED costs, K = 2^18 .. 2^20):
  // 1024-thread CTAs, at most one per SM — the dense radix pass flushes up to 2048 bins per CTA with global atomics and
  // every phase ends in a grid barrier, so fewer, fatter CTAs (8-GPU trace: 170-214 µs per iteration with 512-592 small
  // CTAs). Keys per CTA, select phase of a control step (9 selections) on one GPU, profiles/r2_multi_gpu.md:
  //   K = 262 144: 8192 -> 0.61 ms, 4096 -> 0.46, 2048 -> 0.38 (256-thread path: 0.38);  K = 524 288: 0.61 / 0.45 / 0.46 (0.48)
  static int big_thr = -1, big_keys = 2048;  // MPOPIS_SELECT_BIG / MPOPIS_SELECT_KEYS: A/B of the CTA shape (tools)
  if (big_thr < 0) {
    const char *e = getenv("MPOPIS_SELECT_BIG"), *k = getenv("MPOPIS_SELECT_KEYS");
    big_thr = e ? atoi(e) : 131072;
    if (k && atoi(k) >= 1024) big_keys = atoi(k);
  }
  const bool big = Ktot > big_thr;
  int grid = big ? (Ktot + big_keys - 1) / big_keys : (Ktot + 1023) / 1024;
  const int cap = big ? (max_ctas / 4 > 148 ? 148 : max_ctas / 4) : max_ctas;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  Ws *w = (Ws *)ws;
  void *args[] = {(void *)&costs, (void *)&Ktot,  (void *)&m,     (void *)&k0,    (void *)&Kloc,
                  (void *)&early_stop, (void *)&w, (void *)&bmin,  (void *)&bmax,  (void *)&nb_cap,
                  (void *)&eidx,  (void *)&m_loc, (void *)&tau_out, (void *)&stop_flag, (void *)&stop};
  if (big) return (int)cudaLaunchCooperativeKernel((const void *)ce_select_kernel<1024>, dim3(grid), dim3(1024), args, 0, s);
  return (int)cudaLaunchCooperativeKernel((const void *)ce_select_kernel<256>, dim3(grid), dim3(256), args, 0, s);
}

int elite_gather_nchunks(int m_max) { return (m_max + GS_COLS - 1) / GS_COLS; }

void launch_elite_gather_sums(const double *E, long long ldk, int cs, const int *eidx, const int *m_loc, int m_max,
                              double *X, long long ldx, double *partial, const int *stop, cudaStream_t s) {
  dim3 grid(elite_gather_nchunks(m_max), cs);
  elite_gather_sums_kernel<<<grid, 256, 0, s>>>(E, ldk, cs, eidx, m_loc, X, ldx, partial, stop);
}

void launch_ce_sums(const double *partial, int nchunks, int cs, const int *m_loc, double *sums, int finalize,
                    int standardise, double *mu, double *U, double *dinv, const int *stop, cudaStream_t s) {
  ce_sums_kernel<<<(cs + 1 + 7) / 8, 256, 0, s>>>(partial, nchunks, cs, m_loc, sums, finalize, standardise, mu, U, dinv,
                                                 stop);
}

}  // namespace mpopis
