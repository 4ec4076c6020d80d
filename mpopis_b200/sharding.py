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


# ---- numpy restatement of csrc/select.cu (elite selection without a sort) ------------------------------------------
COST_KEY_NAN = 0xFFFFFFFFFFFFFFFE


def cost_keys(costs: np.ndarray) -> np.ndarray:
    """Order-preserving uint64 image of a Float64 under Base.isless (engine.cuh: cost_key): −0.0 < +0.0, NaN last."""
    c = np.ascontiguousarray(costs, dtype=np.float64)
    b = c.view(np.uint64)
    k = np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))
    return np.where(np.isnan(c), np.uint64(COST_KEY_NAN), k)


def key_costs(keys: np.ndarray) -> np.ndarray:
    k = np.asarray(keys, dtype=np.uint64)
    b = np.where(k >> np.uint64(63), k & np.uint64((1 << 63) - 1), ~k)
    return np.where(k >= np.uint64(COST_KEY_NAN), np.nan, b.view(np.float64))


def ce_select_emulation(costs: np.ndarray, m: int, k0: int = 0, kloc: int | None = None, early_stop: bool = True):
    """The algorithm of ce_select_kernel on the host, step by step (digits, prefixes, buckets as the kernel forms them):
    returns (ids of the elites inside the window [k0, k0 + kloc) in index order, stop decision, τ as (key, index)).

    1. exact radix select of the m-th smallest 96-bit composite (cost key << 32 | sample index): 11-bit digits from the
       top, the candidate set narrowed per pass, ranked directly once <= 256 remain;
    2. early stop (POL:458-461) without sorting the elites: decided "no" by the gap of the two smallest costs, by a
       NaN / non-finite elite cost or by the pigeonhole bound (c_m − c₁)·200 >= 2m + 2; otherwise from the gaps between
       consecutive non-empty buckets of width 0.005 (min / max key per bucket, the reference's own subtraction);
    3. ownership: the elites whose global id falls into the shard's window."""
    K = len(costs)
    kloc = K if kloc is None else kloc
    keys = cost_keys(costs)
    comp = [(int(keys[i]) << 32) | i for i in range(K)]
    need, cand, low = m - 1, list(range(K)), 96
    tau = None
    for pas in range(9):
        shift, mask = (85 - 11 * pas, 0x7FF) if pas < 8 else (0, 0xFF)
        hist: dict[int, list[int]] = {}
        for i in cand:
            hist.setdefault((comp[i] >> shift) & mask, []).append(i)
        for d in sorted(hist):
            if need < len(hist[d]):
                cand = hist[d]
                break
            need -= len(hist[d])
        low = shift
        if len(cand) <= 256:  # direct ranking
            tau = sorted(comp[i] for i in cand)[need]
            break
    assert tau is not None
    tk, ti = tau >> 32, tau & 0xFFFFFFFF
    elite = np.array([comp[i] <= tau for i in range(K)])
    assert int(elite.sum()) == m
    # ---- stop decision ----
    order2 = np.sort(keys)[:2]
    g1, g2 = int(order2[0]), (int(order2[1]) if K > 1 else None)
    first_gap_decides = m > 1 and g2 is not None and not (abs(float(key_costs(np.array([g2]))[0]) - float(key_costs(np.array([g1]))[0])) < 10e-3)
    c1, cm = float(key_costs(np.array([g1], dtype=np.uint64))[0]), float(key_costs(np.array([tk], dtype=np.uint64))[0])
    stop = False
    if early_stop and m > 1 and not first_gap_decides and tk < COST_KEY_NAN and np.isfinite(c1) and np.isfinite(cm):
        span = (cm - c1) * 200.0
        nb_cap = 2 * m + 4
        if span < 2 * m + 2 and span + 1.0 <= nb_cap:
            nb = int(span) + 1
            bmin, bmax = [None] * nb, [None] * nb
            for i in np.nonzero(elite)[0]:
                bk = min(max(int((float(costs[i]) - c1) * 200.0), 0), nb - 1)
                k = int(keys[i])
                bmin[bk] = k if bmin[bk] is None else min(bmin[bk], k)
                bmax[bk] = k if bmax[bk] is None else max(bmax[bk], k)
            gap_seen, last = False, None
            for bk in range(nb):
                if bmin[bk] is None:
                    continue
                if last is not None:
                    a = float(key_costs(np.array([bmin[bk]], dtype=np.uint64))[0])
                    b = float(key_costs(np.array([last], dtype=np.uint64))[0])
                    if not (abs(a - b) < 10e-3):
                        gap_seen = True
                last = bmax[bk]
            stop = not gap_seen
    ids = np.nonzero(elite)[0]
    return ids[(ids >= k0) & (ids < k0 + kloc)], stop, (tk, ti)
