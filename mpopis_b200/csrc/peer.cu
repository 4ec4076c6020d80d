// peer.cu — one-shot all-reduce(sum) over NVLink/NVSwitch peer memory for the small per-iteration
// statistics of a sharded policy ([Σx, n, early-stop statistics], the cs x cs scatter matrix, the shrinkage
// scalar: 0.8 – 720 KB).
//
// Why not NCCL: these messages are latency-, not bandwidth-bound. The phase trace (MPOPIS_TRACE=1,
// profiles/README.md) measured 20–70 µs per ncclAllReduce at 4 ranks, four times per AIS iteration — more
// than the kernels whose results they carry. Every B200 of the box reaches every peer through NVSwitch, so
// each rank publishes its vector into its OWN mailbox (device memory exported with cudaIpcGetMemHandle and
// mapped by all peers) and then reads all G mailboxes directly, summing in rank order: one kernel, a few µs,
// and the result is bitwise identical on every rank (NCCL's ring/tree order is not).
//
// Protocol (per call, sequence number seq = 1, 2, ...; slot = seq & 1):
//   publish  every CTA copies its slice of `buf` into mailbox.data[slot]; __threadfence_system(); the last CTA
//            to arrive (atomic counter) stores seq into mailbox.flag[slot].
//   wait     one thread per CTA spins (volatile loads over NVLink) until every peer's flag[slot] >= seq.
//   reduce   out[i] = Σ_r peer[r].data[slot][i], r = 0..G-1 in order, 8-byte peer loads.
// A slot is reused two sequence numbers later; a rank can only get there after every peer has published
// seq+1, i.e. has finished reading seq — so two slots suffice and no trailing barrier is needed. The spin has
// a wall-clock budget; on expiry the kernel raises the handle's info flag instead of hanging the GPU.
#include "engine.cuh"

namespace mpopis {

__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerMailboxes pm, double *__restrict__ buf, int n,
                                                              unsigned long long seq, int *info,
                                                              const int *stop) {
  // NOTE: no early return on `stop`: every rank must take part in every collective (stop is identical on all
  // ranks, but the mailboxes' sequence numbers must advance in lock-step).
  (void)stop;
  const int slot = (int)(seq & 1ULL);
  double *mine = pm.data[pm.rank] + (size_t)slot * pm.capacity;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < n; i += nth) mine[i] = buf[i];
  __syncthreads();
  __shared__ int ok;
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative: orders the CTA's stores (observed through the barrier) before the flag
    unsigned int *arrive = pm.arrive + slot;
    if (atomicAdd(arrive, 1u) == gridDim.x - 1) {  // last CTA of this rank: everything is published
      *arrive = 0u;
      __threadfence_system();
      *(volatile unsigned long long *)(pm.flag[pm.rank] + slot) = seq;
    }
    ok = 1;
    const long long t0 = clock64();
    for (int r = 0; r < pm.world; ++r) {
      const volatile unsigned long long *f = (const volatile unsigned long long *)(pm.flag[r] + slot);
      while (*f < seq) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer never arrived
          ok = 0;
          break;
        }
      }
      if (!ok) break;
    }
    __threadfence_system();
  }
  __syncthreads();
  if (!ok) {
    if (threadIdx.x == 0) atomicCAS(info, 0, 3000);
    return;
  }
  for (int i = tid; i < n; i += nth) {
    double s = 0.0;
    for (int r = 0; r < pm.world; ++r) s += ((const volatile double *)(pm.data[r] + (size_t)slot * pm.capacity))[i];
    buf[i] = s;
  }
}

void launch_peer_allreduce(const PeerMailboxes &pm, double *buf, int n, unsigned long long seq, int *info,
                           const int *stop, cudaStream_t s) {
  int grid = (n + 1023) / 1024;
  if (grid > 32) grid = 32;
  if (grid < 1) grid = 1;
  peer_allreduce_kernel<<<grid, 256, 0, s>>>(pm, buf, n, seq, info, stop);
}

}  // namespace mpopis
