// comm.cuh — communicator of a sample-sharded policy: NCCL (one process per GPU) or an in-process loop-back group
// of virtual ranks on one device (verification of the sharded code path on a single GPU). See comm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mpopis {

constexpr int COMM_MAX_WORLD = 64;

struct LoopGroup;  // opaque: shared by the virtual ranks of one loop-back group

struct Comm {
  int world = 1, rank = 0;
  void *nccl = nullptr;      // ncclComm_t
  LoopGroup *loop = nullptr;
  double *loop_tmp = nullptr;  // loop-back all-reduce scratch
  size_t loop_tmp_n = 0;
  bool host_synchronous() const { return loop != nullptr; }  // loop-back collectives block the host: no graph capture
};

// all return 0 on success, -1 on failure with comm_error() describing it
const char *comm_error();
int comm_unique_id(void *out128);
int comm_init_nccl(Comm &c, const void *id128);
LoopGroup *loop_group_create(int world);
void loop_group_destroy(LoopGroup *g);
int comm_init_loopback(Comm &c, LoopGroup *g, int device, size_t max_doubles);
void comm_destroy(Comm &c);
int comm_allreduce_sum(Comm &c, double *buf, size_t n, cudaStream_t st);
// in place: rank r's n_per_rank doubles live at base + r * n_per_rank
int comm_allgather_f64(Comm &c, double *base, size_t n_per_rank, cudaStream_t st);

}  // namespace mpopis
