// comm.cu — the communicator of a sample-sharded policy: NCCL, or an in-process loop-back.
//
// The K rollouts shard by sample; per AIS iteration the engine needs exactly two kinds of exchange: an in-place
// all-gather of the cost vector (8 B per sample) and all-reduce(sum) of a few moment vectors (cs .. cs² doubles).
//   * NCCL (production): one process per GPU, libnccl.so.2 dlopen'ed (the library has no link-time NCCL dependency;
//     a host that already loaded torch's copy shares it), calls enqueued on the engine's stream.
//   * loop-back (verification): G handles created with world_size = G on ONE device in ONE process, each driven by its
//     own host thread; a collective is a host barrier between the threads plus plain device kernels/copies that read
//     the peers' buffers directly (same address space). The sharded code path — selection on the gathered costs,
//     ownership-compacted elite moments, fixed-order reductions — is then exactly the multi-GPU one, and it runs on a
//     single-GPU box (tests/test_gpu_loopback.py). The reduction order is rank 0..G-1 on every rank, so all virtual
//     ranks hold bit-identical results, like NCCL's.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "comm.cuh"

namespace mpopis {

namespace {

thread_local char g_cerr[512] = "";
int cfail(const char *fmt, const char *a = "", const char *b = "") {
  snprintf(g_cerr, sizeof g_cerr, fmt, a, b);
  return -1;
}

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
#define SYM(field, name) field = (decltype(field))dlsym(lib, name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather;
  }
} g_nccl;

#define NC(call)                                                                                    \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess) return cfail("%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)
#define CUC(call)                                                                     \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) return cfail("%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

struct PtrList {
  const double *p[COMM_MAX_WORLD];
};
// out[i] = Σ_r src[r][i] in rank order (identical on every virtual rank)
__global__ void loop_sum_kernel(double *__restrict__ out, const PtrList src, int G, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = src.p[0][i];
  for (int r = 1; r < G; ++r) s += src.p[r][i];
  out[i] = s;
}

}  // namespace

struct LoopGroup {
  int world = 0, device = -1, attached = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  bool broken = false;
  const double *pub[COMM_MAX_WORLD] = {};
  char *peer_region[COMM_MAX_WORLD] = {};
  double *peer_gather[COMM_MAX_WORLD] = {};
  // generation barrier between the host threads of the virtual ranks; a rank that never arrives (it returned an
  // error before the collective) breaks the group instead of hanging the others forever
  int barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return -1;
    const unsigned long long g = gen;
    if (++arrived == world) {
      arrived = 0, ++gen;
      cv.notify_all();
      return 0;
    }
    if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || broken; })) broken = true, cv.notify_all();
    return broken ? -1 : 0;
  }
};

const char *comm_error() { return g_cerr; }

int comm_unique_id(void *out128) {
  if (!g_nccl.load()) return cfail("libnccl.so.2 not found: %s", dlerror());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init_nccl(Comm &c, const void *id128) {
  if (c.world == 1) return 0;
  if (!g_nccl.load()) return cfail("libnccl.so.2 not found: %s", dlerror());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t nc = nullptr;
  NC(g_nccl.CommInitRank(&nc, c.world, id, c.rank));
  c.nccl = nc;
  return 0;
}

LoopGroup *loop_group_create(int world) {
  if (world < 1 || world > COMM_MAX_WORLD) return nullptr;
  LoopGroup *g = new LoopGroup();
  g->world = world;
  return g;
}
void loop_group_destroy(LoopGroup *g) { delete g; }

int comm_init_loopback(Comm &c, LoopGroup *g, int device, size_t max_doubles) {
  if (!g) return cfail("null loop-back group");
  std::lock_guard<std::mutex> lk(g->mu);
  if (g->world != c.world) return cfail("loop-back group was created for a different world size");
  if (g->device >= 0 && g->device != device) return cfail("all virtual ranks of a loop-back group must share one device");
  g->device = device;
  g->attached += 1;
  CUC(cudaMalloc((void **)&c.loop_tmp, sizeof(double) * max_doubles));
  c.loop_tmp_n = max_doubles;
  c.loop = g;
  return 0;
}

void comm_destroy(Comm &c) {
  if (c.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl);
  for (void *&p : c.peer_opened)
    if (p) cudaIpcCloseMemHandle(p), p = nullptr;
  if (c.peer_region) cudaFree(c.peer_region);
  c.peer_region = nullptr, c.peer = false;
  if (c.loop_tmp) cudaFree(c.loop_tmp);
  c.nccl = nullptr, c.loop = nullptr, c.loop_tmp = nullptr;
}

// =====================================================================================================================
// Peer-memory collectives: one kernel per exchange, stores into the peers' HBM over NVLink / NVSwitch
// =====================================================================================================================
// Every rank owns a `region` = PeerCtrl + 2 (parity) x world all-reduce slots, and its full-length cost vector; every
// rank holds pointers to all of them (PeerTable). Per control step the engine runs ~3 tiny exchanges per AIS iteration
// (8 B x K_loc costs, 2cs+1 and cs²+1 moment sums): their cost is latency, not bytes, and NCCL's launch + protocol
// latency (≈35 µs each on 8 GPUs, profiles/r2_multi_gpu.md) was most of the weak-scaling loss. Here:
//   all-reduce  push my vector into slot[parity][my rank] of EVERY rank as self-validating 16-byte lines {lo, flag, hi,
//               flag} (flag = epoch + 1: no fence, no separate flag, one NVLink trip) -> spin on the lines of my own
//               slots and sum them in rank order (bit-identical on all ranks, no atomics on data) -> last CTA bumps the
//               device-resident epoch. Two parities suffice: a rank can only reach epoch e+2 after it received
//               everyone's e+1 contribution, which each rank sends only after it finished reading epoch e.
//   all-gather  store my K_loc costs straight into segment [my rank] of every rank's cost vector, flag, wait. One
//               buffer suffices because an all-reduce always separates two gathers (the step ends in one) and the
//               consumers of the gathered costs are stream-ordered before this rank's contribution to it.
// Epochs live in device memory and are advanced by the kernels, so a captured CUDA graph replays them unchanged.
// A spin that sees no peer for ~30 s sets *peer_err (sticky, surfaces as MPOPIS_ERR_NCCL) instead of hanging the GPU.
namespace {

constexpr size_t PEER_CTRL_BYTES = 1024;
struct PeerCtrl {
  unsigned epoch_ar, epoch_ag, cnt_push, cnt_exit, pad[4];
  unsigned flags_ar[2][COMM_PEER_MAX];
  unsigned flags_ag[COMM_PEER_MAX];
};
static_assert(sizeof(PeerCtrl) <= PEER_CTRL_BYTES, "control block fits its reservation");

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *flag reaches `want` (epochs only grow); false on timeout. Once a collective of this handle has timed out
// (*err set, sticky until the host reads it) later ones give up after ~1 ms instead of 30 s each.
__device__ bool wait_flag(const unsigned *flag, unsigned want, const int *err) {
  const long long t0 = clock64();
  const long long limit = (err && *(const volatile int *)err == COMM_PEER_TIMEOUT) ? 2000000LL : 60000000000LL;
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(40);
    if (clock64() - t0 > limit) return false;
  }
  return true;
}

// All-reduce, "LL" wire format (the idea of NCCL's low-latency protocol): every double travels as ONE 16-byte store
// {lo, flag, hi, flag} with flag = epoch + 1, so the data validates itself — no fence, no separate flag store, no second
// NVLink round trip. Each 8-byte half is written atomically, hence a line whose two flags match is complete. The
// receiver spins on the lines of its own slots and sums them in rank order.
struct __align__(16) LLLine {
  unsigned lo, f0, hi, f1;
};
__device__ __forceinline__ LLLine *peer_slot(char *region, int par, int world, int src, size_t slot_n) {
  return (LLLine *)(region + PEER_CTRL_BYTES) + ((size_t)par * world + src) * slot_n;
}
__device__ __forceinline__ void ll_store(LLLine *dst, double v, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((unsigned)__double2loint(v)), "r"(flag),
               "r"((unsigned)__double2hiint(v)), "r"(flag)
               : "memory");
}
// false on timeout (≈30 s; ≈1 ms once a collective of this handle has already timed out)
__device__ __forceinline__ bool ll_load(const LLLine *src, unsigned flag, double *v, const int *err) {
  unsigned lo, f0, hi, f1;
  long long t0 = 0;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
    if (f0 == flag && f1 == flag) break;
    if ((spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      const long long limit = (err && *(const volatile int *)err == COMM_PEER_TIMEOUT) ? 2000000LL : 60000000000LL;
      if (now - t0 > limit) return false;
    }
  }
  *v = __hiloint2double((int)hi, (int)lo);
  return true;
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(double *__restrict__ buf, int n, const PeerTable tab, int world,
                                                              int rank, size_t slot_n, int *err) {
  PeerCtrl *me = (PeerCtrl *)tab.region[rank];
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile unsigned *)&me->epoch_ar;
  __syncthreads();
  const unsigned e = s_epoch, flag = e + 1;
  const int par = e & 1;
  const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = i0; i < n; i += stride) {
    const double v = buf[i];
    for (int p = 0; p < world; ++p)
      if (p != rank) ll_store(peer_slot(tab.region[p], par, world, rank, slot_n) + i, v, flag);
  }
  bool ok = true;
  for (int i = i0; i < n; i += stride) {
    const double own = buf[i];
    double s = 0.0;
    for (int q = 0; q < world; ++q) {  // rank order: bit-identical on every rank
      double v = own;
      if (q != rank) ok = ll_load(peer_slot(tab.region[rank], par, world, q, slot_n) + i, flag, &v, err) && ok;
      s = q == 0 ? v : s + v;
    }
    buf[i] = s;
  }
  if (!ok && err) atomicCAS(err, 0, COMM_PEER_TIMEOUT);
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&me->cnt_exit, 1u) == gridDim.x - 1) {
    me->cnt_exit = 0;
    __threadfence();
    *(volatile unsigned *)&me->epoch_ar = e + 1;
  }
}

__global__ void __launch_bounds__(512) peer_allgather_kernel(const double *__restrict__ base, size_t n_per, const PeerTable tab,
                                                              int world, int rank, int *err) {
  PeerCtrl *me = (PeerCtrl *)tab.region[rank];
  __shared__ unsigned s_epoch;
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_epoch = *(volatile unsigned *)&me->epoch_ag;
  __syncthreads();
  const unsigned e = s_epoch;
  const size_t off = (size_t)rank * n_per, stride = (size_t)gridDim.x * blockDim.x;
  if (n_per % 2 == 0) {  // 16-byte stores
    const double2 *src = (const double2 *)(base + off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per / 2; i += stride) {
      const double2 v = src[i];
      for (int p = 0; p < world; ++p)
        if (p != rank) ((double2 *)(tab.gather[p] + off))[i] = v;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += stride) {
      const double v = base[off + i];
      for (int p = 0; p < world; ++p)
        if (p != rank) tab.gather[p][off + i] = v;
    }
  }
  __syncthreads();  // the CTA's stores happen-before thread 0's system-scope fence (cumulative): one fence per CTA
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(&me->cnt_push, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // the last CTA to finish its stores publishes the segment and waits for everyone else's
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int p = 0; p < world; ++p) st_release_sys(&((PeerCtrl *)tab.region[p])->flags_ag[rank], e + 1);
  }
  if (threadIdx.x < world && threadIdx.x != rank && !wait_flag(&me->flags_ag[threadIdx.x], e + 1, err) && err)
    atomicCAS(err, 0, COMM_PEER_TIMEOUT);
  __syncthreads();
  if (threadIdx.x == 0) {
    me->cnt_push = 0;
    __threadfence();
    *(volatile unsigned *)&me->epoch_ag = e + 1;
  }
}

}  // namespace

int comm_peer_alloc(Comm &c, size_t slot_doubles) {
  if (c.world == 1 || c.peer_region) return 0;
  if (c.world > COMM_PEER_MAX) return cfail("peer-memory collectives support at most 16 ranks");
  const size_t bytes = PEER_CTRL_BYTES + sizeof(LLLine) * 2 * (size_t)c.world * slot_doubles;
  CUC(cudaMalloc((void **)&c.peer_region, bytes));
  CUC(cudaMemset(c.peer_region, 0, bytes));
  CUC(cudaDeviceSynchronize());  // zeroed before any peer can learn the address
  cudaFuncAttributes fa;         // load both kernels now, not under a spinning peer
  CUC(cudaFuncGetAttributes(&fa, peer_allreduce_kernel));
  CUC(cudaFuncGetAttributes(&fa, peer_allgather_kernel));
  c.peer_slot_n = slot_doubles;
  return 0;
}

int comm_peer_export(Comm &c, double *gather_base, void *out128) {
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == COMM_PEER_HANDLE, "two IPC handles per rank");
  if (!c.peer_region) return cfail("comm_peer_alloc first");
  cudaIpcMemHandle_t hd[2];
  CUC(cudaIpcGetMemHandle(&hd[0], c.peer_region));
  CUC(cudaIpcGetMemHandle(&hd[1], gather_base));
  memcpy(out128, hd, sizeof hd);
  return 0;
}

int comm_peer_attach(Comm &c, double *gather_base, const void *all_handles) {
  if (!c.peer_region) return cfail("comm_peer_alloc first");
  if (c.peer) return cfail("peer-memory collectives already attached");
  for (int r = 0; r < c.world; ++r) {
    if (r == c.rank) {
      c.peer_tab.region[r] = c.peer_region, c.peer_tab.gather[r] = gather_base;
      continue;
    }
    cudaIpcMemHandle_t hd[2];
    memcpy(hd, (const char *)all_handles + (size_t)r * COMM_PEER_HANDLE, sizeof hd);
    void *a = nullptr, *b = nullptr;
    CUC(cudaIpcOpenMemHandle(&a, hd[0], cudaIpcMemLazyEnablePeerAccess));
    c.peer_opened[2 * r] = a;
    CUC(cudaIpcOpenMemHandle(&b, hd[1], cudaIpcMemLazyEnablePeerAccess));
    c.peer_opened[2 * r + 1] = b;
    c.peer_tab.region[r] = (char *)a, c.peer_tab.gather[r] = (double *)b;
  }
  c.peer = true;
  return 0;
}

// CUDA's lazy module loading may synchronise the context when a kernel is first used. Between PROCESSES that is harmless
// (a rank's spinning kernel only needs the peer's kernel of the same collective, launched before any later load), but
// the virtual ranks of a loop-back group share one context: rank A's kernel spinning on rank B's flag would block the
// very load rank B's launch is waiting for. Loop-back peer mode therefore requires CUDA_MODULE_LOADING=EAGER.
static int module_loading_is_eager() {
  void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return -1;
  typedef int (*fn_t)(int *);
  fn_t fn = (fn_t)dlsym(lib, "cuModuleGetLoadingMode");
  int mode = 0;
  if (!fn || fn(&mode) != 0) return -1;
  return mode == 1;  // CU_MODULE_EAGER_LOADING = 0x1, CU_MODULE_LAZY_LOADING = 0x2
}

int comm_peer_attach_loopback(Comm &c, double *gather_base) {
  if (!c.loop) return cfail("not a loop-back communicator");
  if (!c.peer_region) return cfail("comm_peer_alloc first");
  if (module_loading_is_eager() != 1)
    return cfail("peer-memory collectives between loop-back ranks need CUDA_MODULE_LOADING=EAGER in the environment "
                 "(virtual ranks share one CUDA context: a lazy kernel load would wait for the spinning peer)");
  LoopGroup *g = c.loop;
  g->peer_region[c.rank] = c.peer_region, g->peer_gather[c.rank] = gather_base;
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  for (int r = 0; r < c.world; ++r) c.peer_tab.region[r] = g->peer_region[r], c.peer_tab.gather[r] = g->peer_gather[r];
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  c.peer = true;
  return 0;
}

int comm_host_barrier(Comm &c) {
  if (!c.loop) return 0;
  if (c.loop->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  return 0;
}

int comm_allreduce_sum(Comm &c, double *buf, size_t n, cudaStream_t st) {
  if (c.world == 1) return 0;
  if (c.peer && n <= c.peer_slot_n) {
    const unsigned grid = (unsigned)std::min<size_t>(8, (n + 1023) / 1024);
    peer_allreduce_kernel<<<grid, 512, 0, st>>>(buf, (int)n, c.peer_tab, c.world, c.rank, c.peer_slot_n, c.peer_err);
    CUC(cudaGetLastError());
    return 0;
  }
  if (c.nccl) {
    NC(g_nccl.AllReduce(buf, buf, n, ncclFloat64, ncclSum, (ncclComm_t)c.nccl, st));
    return 0;
  }
  if (!c.loop) return cfail("sharded handle without a communicator: call comm_init first");
  LoopGroup *g = c.loop;
  if (n > c.loop_tmp_n) return cfail("loop-back all-reduce larger than its scratch");
  CUC(cudaStreamSynchronize(st));  // this rank's contribution is complete
  g->pub[c.rank] = buf;
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  PtrList pl;
  for (int r = 0; r < c.world; ++r) pl.p[r] = g->pub[r];
  loop_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c.loop_tmp, pl, c.world, n);
  CUC(cudaStreamSynchronize(st));
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");  // all reads done
  CUC(cudaMemcpyAsync(buf, c.loop_tmp, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int comm_allgather_f64(Comm &c, double *base, size_t n_per_rank, cudaStream_t st) {
  if (c.world == 1) return 0;
  if (c.peer && base == c.peer_tab.gather[c.rank]) {
    const unsigned grid = (unsigned)std::min<size_t>(32, (n_per_rank / 2 + 2047) / 2048);
    peer_allgather_kernel<<<grid, 512, 0, st>>>(base, n_per_rank, c.peer_tab, c.world, c.rank, c.peer_err);
    CUC(cudaGetLastError());
    return 0;
  }
  if (c.nccl) {
    NC(g_nccl.AllGather(base + (size_t)c.rank * n_per_rank, base, n_per_rank, ncclFloat64, (ncclComm_t)c.nccl, st));
    return 0;
  }
  if (!c.loop) return cfail("sharded handle without a communicator: call comm_init first");
  LoopGroup *g = c.loop;
  CUC(cudaStreamSynchronize(st));
  g->pub[c.rank] = base;
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  for (int r = 0; r < c.world; ++r)
    if (r != c.rank)  // peers only ever write the segments they do not own: no race with their own gathers
      CUC(cudaMemcpyAsync(base + (size_t)r * n_per_rank, g->pub[r] + (size_t)r * n_per_rank, sizeof(double) * n_per_rank,
                          cudaMemcpyDeviceToDevice, st));
  CUC(cudaStreamSynchronize(st));
  if (g->barrier()) return cfail("loop-back group broken (a virtual rank left the collective sequence)");
  return 0;
}

}  // namespace mpopis
