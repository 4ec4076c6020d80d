"""BASELINE config 5's code path (a :cemppi policy sharded by sample) on ONE GPU: G virtual ranks share
the device through the loop-back communicator (csrc/comm.cu), each driven from its own host thread. The kernels are
the multi-GPU ones — all-gathered costs, redundant elite selection, ownership-compacted elite moments, fixed-order
all-reduces — only the transport differs from NCCL. Every rank must reproduce the unsharded engine AND the oracle:
identical AIS iteration counts, identical elite sets, control / U within the north-star tolerance (1e-5)."""
import os

import numpy as np
import pytest
from conftest import configure, engine_kwargs, make_env, synthetic_states

from mpopis_b200 import _lib, sharding
from mpopis_b200.engine import Engine

pytestmark = pytest.mark.gpu
CONTROL_RTOL = 1e-5


def sharded_engines(gpu_bound, policy, env, K, T, N, G, peer=False, **kw):
    """peer=False: host-barrier collectives (the reference transport of the loop-back group); peer=True: the
    peer-memory kernels of csrc/comm.cu — the production single-node transport — between the virtual ranks."""
    if peer and os.environ.get("CUDA_LAUNCH_BLOCKING") == "1":
        pytest.skip("peer-memory collectives between virtual ranks need concurrent kernels (CUDA_LAUNCH_BLOCKING=1 is set)")
    grp = _lib.LoopbackGroup(G)
    engs = []
    for r in range(G):
        e = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, T, N, rank=r, world_size=G, **kw)), env, policy)
        e.comm_init_loopback(grp)
        engs.append(e)
    if peer:
        sharding.run_virtual_ranks([e.comm_peer_loopback for e in engs])
    return grp, engs


def plan_all(engs, state, step, U, Z=None, u=None):
    return sharding.run_virtual_ranks([(lambda e=e: e.plan(state, step, U, Z=Z, resample_u=u)) for e in engs])


@pytest.mark.parametrize("peer", [False, True], ids=["hostbarrier", "peer"])
@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("sigma_est", ["ss", "mle"])
def test_sharded_cemppi_equals_unsharded_and_oracle(gpu_bound, orc, G, sigma_est, peer):
    env = make_env("car")
    K, T, N = 4104, 30, 6  # 4104 = 2^3 · 3^3 · 19: divisible by 2, 3 and 8; > 2048 so one GPU also takes select.cu
    kw = dict(sigma_est=sigma_est)
    grp, engs = sharded_engines(gpu_bound, "cemppi", env, K, T, N, G, peer=peer, **kw)
    one = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, **kw)), env, "cemppi")
    cpu = configure(orc.engine(nthreads=8, **engine_kwargs("cemppi", env, K, T, N, **kw)), env, "cemppi")
    rng = np.random.Generator(np.random.Philox(key=100 + G))
    U = rng.uniform(-0.2, 0.2, one.cs)
    st = synthetic_states()[3]
    for step in range(2):
        Z = rng.standard_normal((one.cs, K, N))
        res = plan_all(engs, st, step, U, Z)
        c1, u1, i1 = one.plan(st, step, U, Z=Z)
        cc, uc, ic = cpu.plan(st, step, U, Z=Z)
        f1, fc = one.fetch(), cpu.fetch()
        S1, Up1 = one.fetch_proposal()
        Sc, Upc = cpu.fetch_proposal()
        assert i1 == ic
        np.testing.assert_allclose(c1, cc, rtol=CONTROL_RTOL, atol=1e-8)
        np.testing.assert_allclose(S1, Sc, rtol=1e-6, atol=1e-12)
        m = int(np.rint(K * 0.2))
        for r, (cr, ur, ir) in enumerate(res):
            assert ir == i1 == ic, f"rank {r}: its {ir} vs unsharded {i1} / oracle {ic}"
            np.testing.assert_allclose(cr, c1, rtol=CONTROL_RTOL, atol=1e-9)
            np.testing.assert_allclose(ur, u1, rtol=CONTROL_RTOL, atol=1e-9)
            np.testing.assert_allclose(cr, cc, rtol=CONTROL_RTOL, atol=1e-8)
            np.testing.assert_allclose(ur, uc, rtol=CONTROL_RTOL, atol=1e-8)
            fr = engs[r].fetch()
            # the shards sum the elite moments in another chunk order than one GPU does: Σ′ (hence L, E and the costs of
            # the later iterations) agree to rounding, not bitwise
            relc = np.abs(fr["costs"] - f1["costs"]) / np.maximum(1.0, np.abs(f1["costs"]))
            assert (relc > 1e-9).sum() <= max(1, K // 500), f"{(relc > 1e-9).sum()} costs differ from the unsharded engine"
            assert np.array_equal(fr["costs"], engs[0].fetch()["costs"]), "all virtual ranks hold the same gathered costs"
            Sr, Upr = engs[r].fetch_proposal()
            np.testing.assert_allclose(Sr, S1, rtol=1e-9, atol=1e-14)   # the last adapted Σ′
            np.testing.assert_allclose(Upr, Up1, rtol=1e-9, atol=1e-12)
            # identical elite membership on the last iteration's costs, rank window by rank window
            ids, _, _ = engs[r].elite_select(fr["costs"], m, k0=r * (K // G), kloc=K // G, early_stop=False)
            lo, hi = r * (K // G), (r + 1) * (K // G)
            elite_ref = np.sort(np.argsort(fr["costs"], kind="stable")[:m])  # select.cu itself: test_gpu_select.py
            assert np.array_equal(ids, elite_ref[(elite_ref >= lo) & (elite_ref < hi)])
        assert all(np.array_equal(res[0][0], x[0]) and np.array_equal(res[0][1], x[1]) for x in res[1:]), \
            "virtual ranks agree bitwise"
        U = uc
        st, _, _, _ = cpu.env_step(st, cc, step)
    for e in engs:
        e.close()
    grp.close()


@pytest.mark.parametrize("peer", [False, True], ids=["hostbarrier", "peer"])
@pytest.mark.parametrize("policy", ["μΣaismppi", "pmcmppi", "cmamppi", "imppi"])
def test_other_policies_sharded(gpu_bound, orc, policy, peer):
    if peer and policy == "cmamppi":
        pytest.skip("cooperative kernels (Σ^-1/2, merge sort) cannot run beside a spinning peer on the SAME device")
    env = make_env("car")
    K, T, N, G = 1536, 20, 4, 3
    grp, engs = sharded_engines(gpu_bound, policy, env, K, T, N, G, peer=peer)
    cpu = configure(orc.engine(nthreads=8, **engine_kwargs(policy, env, K, T, N)), env, policy)
    rng = np.random.Generator(np.random.Philox(key=5))
    Z, u = rng.standard_normal((cpu.cs, K, cpu.N)), rng.uniform(size=(K, max(cpu.N - 1, 1)))
    U = rng.uniform(-0.2, 0.2, cpu.cs)
    res = plan_all(engs, env.state, 0, U, Z, u)
    cc, uc, ic = cpu.plan(env.state, 0, U, Z=Z, resample_u=u)
    for cr, ur, ir in res:
        assert ir == ic
        np.testing.assert_allclose(cr, cc, rtol=CONTROL_RTOL, atol=1e-8)
        np.testing.assert_allclose(ur, uc, rtol=CONTROL_RTOL, atol=1e-8)
    for e in engs:
        e.close()
    grp.close()


def test_sharded_early_stop(gpu_bound, orc):
    """POL:459-461 across shards: zero noise -> identical elite costs -> every rank breaks at n = 1."""
    env = make_env("car")
    K, T, N, G = 96, 10, 5, 4
    grp, engs = sharded_engines(gpu_bound, "cemppi", env, K, T, N, G, sigma_est="mle")
    cpu = configure(orc.engine(nthreads=4, **engine_kwargs("cemppi", env, K, T, N, sigma_est="mle")), env, "cemppi")
    Z = np.zeros((cpu.cs, K, N))
    U = np.full(cpu.cs, 0.05)
    res = plan_all(engs, env.state, 0, U, Z)
    cc, uc, ic = cpu.plan(env.state, 0, U, Z=Z)
    assert ic == 1
    for cr, ur, ir in res:
        assert ir == 1
        np.testing.assert_allclose(cr, cc, rtol=1e-12)
        np.testing.assert_allclose(ur, uc, rtol=1e-12)
    for e in engs:
        e.close()
    grp.close()


@pytest.mark.parametrize("peer", [False, True], ids=["hostbarrier", "peer"])
def test_device_rng_is_independent_of_the_sharding(gpu_bound, peer):
    """The Philox counter is the GLOBAL sample id: a sharded policy draws the same noise as the unsharded one. With
    peer=True the step runs as a captured CUDA graph containing the peer-memory collectives (epochs live on the
    device), replayed for the second control step."""
    env = make_env("car")
    K, T, N, G = 4096, 30, 5, 4
    grp, engs = sharded_engines(gpu_bound, "cemppi", env, K, T, N, G, peer=peer, sigma_est="ss")
    one = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="ss")), env, "cemppi")
    for e in engs + [one]:
        e.seed(77)
    U = np.zeros(one.cs)
    st = env.state
    for step in range(3):
        res = plan_all(engs, st, step, U)
        c1, u1, i1 = one.plan(st, step, U)
        for cr, ur, ir in res:
            assert ir == i1
            np.testing.assert_allclose(cr, c1, rtol=CONTROL_RTOL, atol=1e-9)
            np.testing.assert_allclose(ur, u1, rtol=CONTROL_RTOL, atol=1e-9)
        relc = np.abs(engs[1].fetch()["costs"] - one.fetch()["costs"]) / np.maximum(1.0, np.abs(one.fetch()["costs"]))
        assert (relc > 1e-9).sum() <= max(1, K // 500)
        U = u1
    for e in engs:
        e.close()
    grp.close()
