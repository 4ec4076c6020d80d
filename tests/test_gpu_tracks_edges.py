"""§8f rank 3 (other tracks, sample factors, variable lane width) and edge cases of the hot path:
the exact look-up-table pruned nearest-point search must return bit-identical indices to the reference's
full scan on every bundled track; tiny / degenerate sizes; the largest single-GPU configuration."""
import numpy as np
import pytest
from conftest import configure, engine_kwargs, make_env

from mpopis_b200.engine import Engine, EngineError
from mpopis_b200.envs import CarRacingEnv
from mpopis_b200.tracks import Track, bundled_track_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", bundled_track_names())
@pytest.mark.parametrize("sf", [20, 10, 3])
def test_track_lut_is_exact_on_every_track(gpu_bound, orc, name, sf):
    if sf == 3 and name not in ("curve", "cubic", "curve4"):
        pytest.skip("dense sampling checked on three tracks")
    trk = Track(name, sample_factor=sf)
    rng = np.random.default_rng(len(name) * 100 + sf)
    width = rng.uniform(6.0, 18.0, trk.x.size)  # variable lane width (Track(infile, width::Vector), TRK:14)
    trk = Track(name, width=width, sample_factor=sf)
    env = CarRacingEnv(track=trk)
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 32, 5)), env, "gmppi")
    c = configure(orc.engine(**engine_kwargs("gmppi", env, 32, 5)), env, "gmppi")
    lo = np.array([trk.xs.min(), trk.ys.min()]) - 70.0  # beyond the table's margin -> full-scan fallback too
    hi = np.array([trk.xs.max(), trk.ys.max()]) + 70.0
    pos = np.concatenate([
        rng.uniform(lo, hi, (6000, 2)),
        np.stack([trk.xs, trk.ys], 1) + rng.normal(0, 1e-9, (trk.xs.size, 2)),           # on the sampled points
        0.5 * (np.stack([trk.xs, trk.ys], 1) + np.roll(np.stack([trk.xs, trk.ys], 1), 1, 0)),  # equidistant mid-points
    ])
    for variant in (0, 3, 4, 1):  # with (0, 3) and without (1) the table
        g.set_option("rollout_variant", variant)
        gi, gj, gd, gw = g.track_query(pos)
        ci, cj, cd, cw = c.track_query(pos)
        assert np.array_equal(gi, ci), f"{name} sf={sf} variant={variant}: {np.sum(gi != ci)} nearest indices differ"
        assert np.array_equal(gj, cj) and np.array_equal(gw, cw)
        np.testing.assert_allclose(gd, cd, rtol=1e-12, atol=1e-12)


def test_rollouts_on_another_track(gpu_bound, orc):
    env = CarRacingEnv(track=Track("cubic", width=12.0, sample_factor=10))
    K, T = 512, 40
    g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, 4, sigma_est="lw")), env, "cemppi")
    c = configure(orc.engine(nthreads=8, **engine_kwargs("cemppi", env, K, T, 4, sigma_est="lw")), env, "cemppi")
    Z = np.random.default_rng(1).standard_normal((g.cs, K, 4))
    st = np.array([env.track.xs[3], env.track.ys[3], np.arctan2(env.track.ys[4] - env.track.ys[3], env.track.xs[4] - env.track.xs[3]),
                   12.0, 0.0, 0.0, 0.0, 0.0])
    (cg, ug, ig), (cc, uc, ic) = g.plan(st, 0, np.zeros(g.cs), Z=Z), c.plan(st, 0, np.zeros(g.cs), Z=Z)
    assert ig == ic
    np.testing.assert_allclose(cg, cc, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(ug, uc, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("policy,K,T,N", [("gmppi", 1, 1, 1), ("mppi", 3, 1, 1), ("cemppi", 10, 1, 3), ("imppi", 2, 2, 2),
                                           ("μΣaismppi", 33, 3, 2), ("pmcmppi", 31, 2, 3), ("cmamppi", 12, 2, 3)])
def test_tiny_sizes(gpu_bound, orc, policy, K, T, N):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, T, N, sigma_est="mle")), env, policy)
    c = configure(orc.engine(**engine_kwargs(policy, env, K, T, N, sigma_est="mle")), env, policy)
    rng = np.random.default_rng(K)
    Z, u = rng.standard_normal((g.cs, K, g.N)), rng.uniform(size=(K, max(g.N - 1, 1)))
    U = rng.uniform(-0.1, 0.1, g.cs)
    try:
        rc = c.plan(env.state, 0, U, Z=Z, resample_u=u)
    except EngineError as e:  # e.g. singular covariance from too few samples: the engine must fail the same way
        with pytest.raises(EngineError) as ei:
            g.plan(env.state, 0, U, Z=Z, resample_u=u)
        assert ei.value.code == e.code
        return
    rg = g.plan(env.state, 0, U, Z=Z, resample_u=u)
    assert rg[2] == rc[2]
    np.testing.assert_allclose(rg[0], rc[0], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(rg[1], rc[1], rtol=1e-5, atol=1e-8)


def test_largest_single_gpu_configuration(gpu_bound):
    """BASELINE config 5's K = 2^20 on ONE GPU (the N = 1 leg of the scaling run): properties only."""
    env = make_env("car")
    K, T, N = 1 << 20, 50, 3
    g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="ss")), env, "cemppi")
    g.seed(5)
    ctrl, U2, its = g.plan(env.state, 0, np.zeros(g.cs))
    f = g.fetch()
    assert its == N and np.all(np.isfinite(ctrl)) and np.all(np.abs(ctrl) <= 1)
    assert abs(f["weights"].sum() - 1) < 1e-9 and f["weights"].argmax() == f["costs"].argmin()
    perm = g.sortperm(f["costs"])
    assert np.array_equal(perm, np.lexsort((np.arange(K), f["costs"])))  # stable arg-sort at full size
