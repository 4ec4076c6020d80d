"""BASELINE config 5 (sample-sharded :cemppi) on real GPUs: a policy sharded over 2 GPUs (one process per
GPU; exchanges over the engine's own NCCL communicator, and over the peer-memory kernels of csrc/comm.cu through CUDA
IPC) must reproduce the single-GPU result. Skipped on a 1-GPU box."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, policy, K, out_dir, peer):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from conftest import configure, engine_kwargs, make_env
    from mpopis_b200 import _lib, sharding
    from mpopis_b200.engine import Engine
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = make_env("car")
    g = configure(Engine(_lib.product(), **engine_kwargs(policy, env, K, 30, 5, sigma_est="ss", device=rank, rank=rank,
                                                        world_size=world)), env, policy)
    assert sharding.connect(g, dist, rank, world, peer=peer) == ("peer" if peer else "nccl")
    g.seed(77)
    U = np.zeros(g.cs)
    st = env.state
    res = []
    for step in range(2):
        ctrl, U, its = g.plan(st, step, U)
        res.append((ctrl, U.copy(), its))
    f = g.fetch()
    np.savez(Path(out_dir) / f"r{rank}.npz", ctrl=res[-1][0], U=res[-1][1], its=res[-1][2], costs=f["costs"],
             weights=f["weights"])
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer"])
@pytest.mark.parametrize("policy", ["cemppi", "μΣaismppi", "pmcmppi", "cmamppi"])
def test_two_gpu_shards_match_single_gpu(tmp_path, gpu_bound, policy, peer):
    import torch.multiprocessing as mp
    from conftest import configure, engine_kwargs, make_env
    from mpopis_b200.engine import Engine
    K = 4096
    mp.spawn(_worker, args=(2, _free_port(), policy, K, str(tmp_path), peer), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, 30, 5, sigma_est="ss")), env, policy)
    g.seed(77)
    U = np.zeros(g.cs)
    for step in range(2):
        ctrl, U, its = g.plan(env.state, step, U)
    f = g.fetch()
    for r in (r0, r1):
        assert int(r["its"]) == its
        np.testing.assert_allclose(r["ctrl"], ctrl, rtol=1e-5, atol=1e-9)   # north-star tolerance
        np.testing.assert_allclose(r["U"], U, rtol=1e-5, atol=1e-9)
        rel = np.abs(r["costs"] - f["costs"]) / np.maximum(1, np.abs(f["costs"]))
        assert (rel > 1e-9).sum() <= max(1, K // 500)
    assert np.array_equal(r0["costs"], r1["costs"]) and np.array_equal(r0["ctrl"], r1["ctrl"])  # ranks agree bitwise
