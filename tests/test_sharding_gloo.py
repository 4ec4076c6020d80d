"""N > 1 path on CPU: world_size-2 gloo processes run the numpy emulation of the engine's sharded
reductions (ownership-masked partial moments -> all-reduce -> finalise) and compare with the unsharded
oracle. Also the shard arithmetic and the comm-id broadcast plumbing (with a stub id, no NCCL here)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, method, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from mpopis_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    rng = np.random.default_rng(123)  # same data on every rank
    cs, K, m = 20, 96, 24
    E = rng.standard_normal((cs, K)) * rng.uniform(0.2, 1.5, (cs, 1))
    costs = np.round(rng.normal(0, 10, K), 1)  # ties on purpose
    k0, kl = sharding.shard_range(K, rank, world)
    # all-gather of the local costs
    parts = [torch.zeros(kl, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(costs[k0:k0 + kl].copy()))
    gathered = torch.cat(parts).numpy()
    order = np.argsort(gathered, kind="stable")
    mu, S, lam = sharding.sharded_elite_moments(E[:, k0:k0 + kl], k0, order, m, allreduce, method)
    w = np.exp(-(gathered - gathered.min()) / 10.0)
    w /= w.sum()
    ws = sharding.sharded_weighted_sum(E[:, k0:k0 + kl], w, k0, allreduce)
    # comm-id broadcast plumbing (stub id: NCCL itself is not involved on the CPU box)
    box = [b"x" * 128 if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    np.savez(Path(out_dir) / f"r{rank}.npz", mu=mu, S=S, lam=lam, ws=ws, gathered=gathered, idlen=len(box[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("method", ["mle", "ss", "oas"])
def test_sharded_moments_match_unsharded_oracle(tmp_path, orc, method):
    from conftest import configure, engine_kwargs, make_env
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), method, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ("mu", "S", "lam", "ws", "gathered"):
        assert np.array_equal(r0[k], r1[k]), f"ranks disagree on {k}"  # every rank must reach the same Σ′
    assert int(r0["idlen"]) == 128
    # unsharded reference: the oracle's cov(method, elite') on the same elites
    rng = np.random.default_rng(123)
    cs, K, m = 20, 96, 24
    E = rng.standard_normal((cs, K)) * rng.uniform(0.2, 1.5, (cs, 1))
    costs = np.round(rng.normal(0, 10, K), 1)
    assert np.array_equal(r0["gathered"], costs)
    order = orc.sortperm(costs)
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("cemppi", env, 64, 10)), env, "cemppi")  # cs = 20
    mu, S = e.cov_estimate(E[:, order[:m]], method)
    np.testing.assert_allclose(r0["mu"], mu, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(r0["S"], S, rtol=1e-10, atol=1e-13)
    assert abs(float(r0["lam"]) - e.last_shrinkage()) < 1e-10
    w = e.weights(costs, 10.0)
    np.testing.assert_allclose(r0["ws"][:-1], E @ w, rtol=1e-12, atol=1e-14)
    assert abs(r0["ws"][-1] - 1.0) < 1e-13


def test_shard_range():
    from mpopis_b200.sharding import shard_range
    assert [shard_range(1 << 20, r, 8) for r in (0, 7)] == [(0, 131072), (7 * 131072, 131072)]
    with pytest.raises(ValueError):
        shard_range(150, 0, 8)
    with pytest.raises(ValueError):
        shard_range(128, 2, 2)


def _reference_select(costs, m):
    """POL:455-461 literally: stable sortperm (Base.isless: −0.0 < 0.0, NaN last), elite slice, maximum(abs.(diff))."""
    from mpopis_b200.sharding import cost_keys
    order = np.lexsort((np.arange(len(costs)), cost_keys(costs)))
    el = costs[order[:m]]
    with np.errstate(invalid="ignore"):
        d = np.abs(np.diff(el))
    gap = np.nan if (d.size and np.isnan(d).any()) else (d.max() if d.size else np.nan)
    return np.sort(order[:m]), bool(gap < 10e-3) if m > 1 else False


def _cost_families(rng, K):
    base = rng.normal(0.0, 3.0, K)
    yield "normal", base
    yield "ties", np.round(base, 1)
    yield "constant", np.full(K, 2.5)
    yield "converged", 7.0 + rng.uniform(0, 4e-3, K)                      # every elite gap < 10e-3: stop
    yield "one_gap_at_threshold", np.concatenate([np.arange(K // 2) * 1e-3, [K * 1e-3 + 9.0e-3], np.full(K - K // 2 - 1, 50.0)])
    yield "one_gap_above", np.concatenate([np.arange(K // 2) * 9.99e-3, np.full(K - K // 2, 50.0)])
    yield "gap_exactly_10e-3", np.concatenate([np.arange(K // 2) * 10e-3, np.full(K - K // 2, 50.0)])
    c = base.copy(); c[::7] = np.nan
    yield "nan_sprinkled", c
    yield "all_nan", np.full(K, np.nan)
    c = np.abs(base) * 1e-3; c[::5] = -0.0; c[1::5] = 0.0
    yield "signed_zeros", c
    c = base.copy(); c[:3] = -np.inf; c[3:6] = np.inf
    yield "infinities", c
    yield "huge_spread", base * 1e200
    yield "denormals", base * 5e-324 * 1e3


@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_selection_algorithm_matches_a_global_stable_sort(G):
    """The sort-free :cemppi selection of csrc/select.cu (radix select on (cost key, sample id), early-stop test from the two
    smallest costs / the pigeonhole bound / value buckets, ownership windows), restated step by step in numpy
    (sharding.ce_select_emulation), picks exactly the m globally smallest samples — ties by global index — and takes the
    reference's stop decision maximum(abs.(diff(elite costs))) < 10e-3 (POL:455-461), for 1, 2, 3 and 8 shards."""
    from mpopis_b200.sharding import ce_select_emulation
    rng = np.random.default_rng(G)
    kloc = 264
    K = G * kloc
    for m in (2, int(round(0.2 * K)), K // 2):
        for name, costs in _cost_families(rng, K):
            costs = np.asarray(costs, dtype=np.float64)
            elite_ref, stop_ref = _reference_select(costs, m)
            got = []
            for r in range(G):
                ids, stop, _ = ce_select_emulation(costs, m, k0=r * kloc, kloc=kloc)
                assert stop == stop_ref, f"{name}, m={m}, shard {r}: stop {stop} vs reference {stop_ref}"
                assert np.all((ids >= r * kloc) & (ids < (r + 1) * kloc))
                got.append(ids)
            assert np.array_equal(np.concatenate(got), elite_ref), f"{name}, m={m}: elite set differs"
    ids, stop, _ = ce_select_emulation(np.array([3.0, 1.0, 2.0]), 1)   # m = 1: nothing to diff, no stop (POL:458 on one cost)
    assert ids.tolist() == [1] and stop is False


class _FakeEngine:
    """Stands in for a CUDA-backed Engine: records what sharding.connect asks of it (no GPU on this box)."""

    def __init__(self, rank, device=0):
        self.rank, self.device, self.calls, self.blobs = rank, device, [], None

    def comm_init(self, nccl_id):
        self.calls.append(("comm_init", len(nccl_id)))

    def comm_peer_export(self):
        self.calls.append(("export",))
        return bytes([self.rank]) * 128

    def comm_peer_attach(self, blobs):
        self.calls.append(("attach", len(blobs)))
        self.blobs = list(blobs)


def _connect_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from mpopis_b200 import _lib, sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _lib.comm_id = lambda: b"i" * 128  # NCCL is not involved on the CPU box
    res = {}
    for name, peer, env in (("auto", None, None), ("off_by_env", None, "0"), ("off_by_arg", False, None)):
        if env is None:
            os.environ.pop("MPOPIS_COMM_PEER", None)
        else:
            os.environ["MPOPIS_COMM_PEER"] = env
        e = _FakeEngine(rank)
        res[name] = (sharding.connect(e, dist, rank, world, peer=peer), e.calls, e.blobs)
    np.save(Path(out_dir) / f"c{rank}.npy", np.array([res], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_connect_negotiates_the_transport_identically_on_every_rank(tmp_path):
    """sharding.connect (N > 1 host plumbing): NCCL id broadcast, then — one host, peer access — every rank exports its
    128-byte blob, all-gathers them IN RANK ORDER and attaches; MPOPIS_COMM_PEER=0 / peer=False keep NCCL. With
    world_size = 1 nothing is touched."""
    from mpopis_b200 import sharding
    solo = _FakeEngine(0)
    assert sharding.connect(solo, None, 0, 1) == "none" and solo.calls == []
    world = 2
    mp.spawn(_connect_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        res = np.load(tmp_path / f"c{rank}.npy", allow_pickle=True)[0]
        kind, calls, blobs = res["auto"]
        assert kind == "peer" and calls == [("comm_init", 128), ("export",), ("attach", world)]
        assert blobs == [bytes([r]) * 128 for r in range(world)]   # rank order, identical on every rank
        for name in ("off_by_env", "off_by_arg"):
            kind, calls, blobs = res[name]
            assert kind == "nccl" and calls == [("comm_init", 128)] and blobs is None
