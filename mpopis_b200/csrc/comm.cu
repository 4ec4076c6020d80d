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
  if (c.loop_tmp) cudaFree(c.loop_tmp);
  c.nccl = nullptr, c.loop = nullptr, c.loop_tmp = nullptr;
}

int comm_allreduce_sum(Comm &c, double *buf, size_t n, cudaStream_t st) {
  if (c.world == 1) return 0;
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
