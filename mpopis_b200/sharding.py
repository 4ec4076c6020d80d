"""Sample-sharding of one policy across the GPUs of a box.

The K rollouts are independent given (state, U, Σ′): rank g owns the contiguous samples
[g·K/G, (g+1)·K/G). Per AIS iteration the engine all-gathers the K costs (so every rank performs the
identical stable selection / weights) and all-reduces the moment sums; the Philox counter is the GLOBAL
sample index, so the draws do not depend on G. This module holds the host-side plumbing and a numpy
emulation of the device's sharded reduction (ownership-masked partial sums -> all-reduce -> finalise)
that the CPU test-suite runs over gloo with world_size = 2.
"""
from __future__ import annotations

import numpy as np


def shard_range(K: int, rank: int, world: int) -> tuple[int, int]:
    """(k0, K_local) of `rank`; the engine requires K % world == 0."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank out of range")
    if K % world:
        raise ValueError("num_samples must be divisible by world_size")
    kl = K // world
    return rank * kl, kl


def broadcast_comm_id(dist, rank: int, src: int = 0) -> bytes:
    """ncclUniqueId of the engine's own communicator, created on `src` and broadcast over an existing
    torch.distributed process group (any backend)."""
    from . import _lib
    box = [_lib.comm_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def connect(engine, dist, rank: int, world: int, peer: bool | None = None) -> str:
    """Join `engine` (created with rank / world_size) to its peers over an initialised torch.distributed group:
    NCCL communicator first, then — when every rank sits on one node with peer access (an NVLink / NVSwitch box) —
    the peer-memory collectives of csrc/comm.cu. `peer`: None = decide (MPOPIS_COMM_PEER=0 disables), True = require,
    False = NCCL only. Returns "peer" or "nccl": the transport of the per-iteration exchanges."""
    import os
    import socket
    if world == 1:
        return "none"
    engine.comm_init(broadcast_comm_id(dist, rank))
    if peer is None:
        peer = os.environ.get("MPOPIS_COMM_PEER", "1") != "0"
    if not peer or world > 16:
        return "nccl"
    import torch
    dev = engine.device
    ok = all(d == dev or torch.cuda.can_device_access_peer(dev, d) for d in range(torch.cuda.device_count()))
    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), bool(ok)))
    if len({h for h, _ in info}) != 1 or not all(o for _, o in info):  # the same verdict on every rank
        return "nccl"
    blobs = [None] * world
    dist.all_gather_object(blobs, engine.comm_peer_export())
    engine.comm_peer_attach(blobs)
    dist.barrier()  # nobody stores into a region before everyone has mapped it
    return "peer"


def run_virtual_ranks(fns):
    """Run one callable per virtual rank of a loop-back group, each on its own host thread (the collectives of a
    loop-back group block until every rank has entered them). Returns the results in rank order; the first
    exception is re-raised after all threads have finished."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(fns)) as ex:
        futs = [ex.submit(f) for f in fns]
        out, err = [], None
        for f in futs:
            try:
                out.append(f.result())
            except Exception as e:  # keep draining: the other ranks unblock through the group's timeout/break
                out.append(None)
                err = err or e
    if err is not None:
        raise err
    return out


def sharded_elite_moments(E_local: np.ndarray, k0: int, order: np.ndarray, m: int, allreduce, method: str = "mle"):
    """numpy emulation of csrc/stats.cu for the CE update on one shard (POL:455-465).

    E_local: cs x K_local (this rank's columns, global ids k0..k0+K_local); order: global stable
    arg-sort of the gathered costs; allreduce(x) sums an array over ranks. Returns (mean, Σ′ without
    the 1e-8 ridge, λ̂)."""
    cs, kl = E_local.shape
    idx = order[:m] - k0
    own = (idx >= 0) & (idx < kl)
    X = np.where(own[None, :], E_local[:, np.clip(idx, 0, kl - 1)], 0.0)   # gather_cols_kernel
    w = own.astype(np.float64)                                             # ownership mask
    sums = allreduce(np.concatenate([X @ w, [w.sum()]]))                   # rowsum + count
    n = sums[-1]
    mu = sums[:-1] / n
    Xc = (X - mu[:, None]) * w[None, :]
    Sraw = allreduce(Xc @ Xc.T)                                            # syrk partials
    S = Sraw / n
    lam = 0.0
    if method in ("lw", "ss"):
        d = 1.0 / np.sqrt(np.diag(S)) if method == "ss" else np.ones(cs)
        Z = Xc * d[:, None]
        z2 = Z ** 2
        q = allreduce(np.array([np.sum(z2.sum(axis=0) ** 2 - (z2 ** 2).sum(axis=0))]))[0]   # shrink_q
        R = S * np.outer(d, d)
        off = ~np.eye(cs, dtype=bool)
        r2 = np.sum(R[off] ** 2)
        lam = float(np.clip((q - n * r2) * n / ((n - 1) * n * n) / r2, 0.0, 1.0))
        S = (1 - lam) * S + lam * np.diag(np.diag(S))
    elif method in ("rblw", "oas"):
        tr, tr2 = np.trace(S), np.sum(S ** 2)
        if method == "rblw":
            lam = ((n - 2) / n * tr2 + tr * tr) / ((n + 2) * (tr2 - tr * tr / cs))
        else:
            lam = ((1 - 2 / cs) * tr2 + tr * tr) / ((n + 1 - 2 / cs) * (tr2 - tr * tr / cs))
        lam = float(np.clip(lam, 0.0, 1.0))
        S = (1 - lam) * S + lam * tr / cs * np.eye(cs)
    return mu, S, lam


def sharded_weighted_sum(E_local: np.ndarray, w_global: np.ndarray, k0: int, allreduce):
    """Σ_k w_k E[:,k] and Σ_k w_k over all shards (the final control, POL:226-231)."""
    kl = E_local.shape[1]
    wl = w_global[k0:k0 + kl]
    return allreduce(np.concatenate([E_local @ wl, [wl.sum()]]))
