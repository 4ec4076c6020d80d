// comm.cuh — communicator of a sample-sharded policy: NCCL (one process per GPU) or an in-process loop-back group
// of virtual ranks on one device (verification of the sharded code path on a single GPU). See comm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mpopis {

constexpr int COMM_MAX_WORLD = 64;
constexpr int COMM_PEER_MAX = 16;      // widest NVSwitch domain the peer-memory collectives are built for
constexpr int COMM_PEER_TIMEOUT = 9000;  // value written to Comm::peer_err
constexpr int COMM_PEER_HANDLE = 128;  // bytes one rank exports: two cudaIpcMemHandle_t (control region, cost vector)

// Peer-memory collectives (comm.cu, "peer" section): every rank maps every other rank's control region and cost vector
// (CUDA IPC between processes, plain pointers between the virtual ranks of a loop-back group) and the exchange is ONE
// kernel per collective that stores into the peers' memory over NVLink and spins on arrival flags in its own.
struct PeerTable {
  char *region[COMM_PEER_MAX];    // PeerCtrl + all-reduce slots of each rank
  double *gather[COMM_PEER_MAX];  // each rank's full-length cost vector
};

struct LoopGroup;  // opaque: shared by the virtual ranks of one loop-back group

struct Comm {
  int world = 1, rank = 0;
  void *nccl = nullptr;      // ncclComm_t
  LoopGroup *loop = nullptr;
  double *loop_tmp = nullptr;  // loop-back all-reduce scratch
  size_t loop_tmp_n = 0;
  bool peer = false;           // peer-memory collectives attached (then NCCL / the host barrier serve only oversize calls)
  char *peer_region = nullptr;
  size_t peer_slot_n = 0;      // doubles one all-reduce slot holds
  PeerTable peer_tab = {};
  void *peer_opened[2 * COMM_PEER_MAX] = {};
  int *peer_err = nullptr;     // device int the kernels set to COMM_PEER_TIMEOUT when a peer never arrives (sticky)
  bool host_synchronous() const { return loop != nullptr && !peer; }  // host-barrier collectives: no graph capture
};

// all return 0 on success, -1 on failure with comm_error() describing it
const char *comm_error();
int comm_unique_id(void *out128);
int comm_init_nccl(Comm &c, const void *id128);
LoopGroup *loop_group_create(int world);
void loop_group_destroy(LoopGroup *g);
int comm_init_loopback(Comm &c, LoopGroup *g, int device, size_t max_doubles);
void comm_destroy(Comm &c);
// peer-memory collectives: alloc (every rank) -> export 128 B -> exchange out of band -> attach (world x 128 B).
// A loop-back group exchanges the pointers itself: comm_peer_attach_loopback (collective over the virtual ranks).
int comm_peer_alloc(Comm &c, size_t slot_doubles);
int comm_peer_export(Comm &c, double *gather_base, void *out128);
int comm_peer_attach(Comm &c, double *gather_base, const void *all_handles);
int comm_peer_attach_loopback(Comm &c, double *gather_base);
int comm_host_barrier(Comm &c);  // loop-back groups only (no-op otherwise): every virtual rank's thread has arrived
int comm_allreduce_sum(Comm &c, double *buf, size_t n, cudaStream_t st);
// in place: rank r's n_per_rank doubles live at base + r * n_per_rank
int comm_allgather_f64(Comm &c, double *base, size_t n_per_rank, cudaStream_t st);

}  // namespace mpopis
