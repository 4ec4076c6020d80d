"""Parity tests proper: the CUDA engine, called through the C-ABI, against (1) the committed golden
fixtures, (2) the live CPU oracle on the same seeded inputs, (3) size-independent properties at the
BASELINE sizes. Tolerances (stated by BASELINE.json:north_star): control within 1e-5 relative of the
CPU path; integer work (track indexing, elite selection / sort order, iteration counts) bit-exact.
The assertions below are much tighter than 1e-5 wherever the arithmetic allows."""
from pathlib import Path

import numpy as np
import pytest
from conftest import configure, engine_kwargs, julia_sortperm, make_env, synthetic_states

from mpopis_b200.engine import Engine, EngineError

pytestmark = pytest.mark.gpu

CONTROL_RTOL = 1e-5  # north-star tolerance
TIGHT = 1e-9         # what the FP64 kernels actually achieve on smooth quantities
POLICIES = ["mppi", "gmppi", "imppi", "cemppi", "cmamppi", "μaismppi", "μΣaismppi", "pmcmppi"]
ASCII = {"μaismppi": "muaismppi", "μΣaismppi": "musigmaaismppi"}
G = np.load(Path(__file__).resolve().parent / "golden" / "golden_v1.npz")


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(1.0, np.abs(np.asarray(b)))


def assert_costs_close(got, ref, max_flip_frac=2e-3):
    """Costs agree to TIGHT except for isolated samples that sit on a penalty threshold (−1e6 / −5000 /
    −11000 terms flip on a 1-ulp state difference); those are counted, not hidden."""
    r = rel(got, ref)
    flips = int((r > TIGHT).sum())
    assert flips <= max(1, int(max_flip_frac * r.size)), f"{flips} of {r.size} costs differ by more than {TIGHT}"
    return flips


def pair(gpu_bound, orc, policy, env, K, T, N=10, variant=None, **kw):
    g = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, T, N, **kw)), env, policy)
    if variant is not None:  # otherwise the engine's default (or MPOPIS_ROLLOUT_VARIANT)
        g.set_option("rollout_variant", variant)
    c = configure(orc.engine(nthreads=8, **engine_kwargs(policy, env, K, T, N, **kw)), env, policy)
    return g, c


# ---------------------------------------------------------------------------------------------------
# 1. golden fixtures
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 3, 4])
def test_golden_rollout_costs(gpu_bound, variant):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 64, 50)), env, "gmppi")
    g.set_option("rollout_variant", variant)
    states = [env.state] + list(G["states"][:5])
    for s, ref in zip(states, G["roll1_costs"]):
        assert_costs_close(g.rollout_costs(s, 0, G["roll1_U"], G["roll1_U"], G["roll1_E"]), ref)
    env3 = make_env("car", 3)
    g3 = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env3, 32, 50)), env3, "gmppi")
    g3.set_option("rollout_variant", variant)
    assert_costs_close(g3.rollout_costs(env3.state, 0, np.zeros(300), np.zeros(300), G["roll3_E"]), G["roll3_costs"])
    mc = make_env("mc")
    gm = configure(Engine(gpu_bound, **engine_kwargs("gmppi", mc, 32, 15, lam=0.1)), mc, "gmppi")
    for t0, ref in zip((0, 192), G["rollmc_costs"]):
        np.testing.assert_allclose(gm.rollout_costs(mc.state, t0, np.zeros(15), np.zeros(15), G["rollmc_E"]), ref, rtol=1e-12)


def test_golden_track_indices_bit_exact(gpu_bound):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    idx, idx2, dist, within = g.track_query(G["trk_pos"])
    assert np.array_equal(idx, G["trk_idx"]) and np.array_equal(idx2, G["trk_idx2"])
    assert np.array_equal(within, G["trk_within"])
    np.testing.assert_allclose(dist, G["trk_dist"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("envname", ["car", "mc"])
def test_golden_control_step(gpu_bound, policy, envname):
    tag = f"plan_{ASCII.get(policy, policy)}_{envname}_"
    if envname == "car":
        env, K, T, N, kw, st = make_env("car"), 96, 20, 4, dict(lam=10.0, sigma_est="ss"), G["states"][2]
    else:
        env, K, T, N, kw = make_env("mc"), 20, 15, 5, dict(lam=0.1, lam_ais=0.1, sigma_est="mle")
        st = env.state
    g = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, T, N, **kw)), env, policy)
    ctrl, U2, its = g.plan(st, 0, np.zeros(g.cs), Z=G[tag + "Z"], resample_u=G[tag + "u"])
    f = g.fetch()
    assert its == int(G[tag + "its"])  # integer: AIS iterations executed (early stop)
    assert_costs_close(f["costs"], G[tag + "costs"])
    np.testing.assert_allclose(ctrl, G[tag + "control"], rtol=CONTROL_RTOL, atol=1e-7)
    np.testing.assert_allclose(U2, G[tag + "U"], rtol=CONTROL_RTOL, atol=1e-7)
    assert np.max(np.abs(f["weights"] - G[tag + "weights"])) < 1e-6


# ---------------------------------------------------------------------------------------------------
# 2. live oracle, same seeded inputs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 3, 4])
@pytest.mark.parametrize("n_cars", [1, 2, 3])
def test_rollout_costs_vs_oracle(gpu_bound, orc, variant, n_cars):
    env = make_env("car", n_cars)
    K, T = 2048 if n_cars == 1 else 384, 50
    g, c = pair(gpu_bound, orc, "gmppi", env, K, T, 1, variant)
    rng = np.random.default_rng(11 + n_cars)
    E = rng.standard_normal((g.cs, K)) * np.tile([0.25, np.sqrt(0.1)], g.cs // 2)[:, None] * 1.5
    U = rng.uniform(-0.4, 0.4, g.cs)
    states = [env.state] if n_cars > 1 else [env.state] + synthetic_states()[:6]
    flips = 0
    for s in states:
        flips += assert_costs_close(g.rollout_costs(s, 0, U, U, E), c.rollout_costs(s, 0, U, U, E))
    print(f"threshold flips: {flips} of {K * len(states)}")


@pytest.mark.parametrize("variant", [0, 3, 4])
def test_rollout_costs_reversing_car_and_wrap(gpu_bound, orc, variant):
    """Vx <= 0 takes the literal slip-angle path (variant 0) or the quadrant-aware ratio form (variant 3, with
    the v3 repair when Vx changes sign while braking); |Ψ| crosses π (heading wrap)."""
    env = make_env("car")
    K, T = 512, 30
    g, c = pair(gpu_bound, orc, "gmppi", env, K, T, 1, variant)
    rng = np.random.default_rng(5)
    E = rng.standard_normal((g.cs, K)) * 0.5
    U = np.tile([0.8, -1.0], T)  # hard braking while steering
    for s in (np.array([20.0, 0.0, 3.1, 2.0, 0.5, 0.4, 0.1, 0.0]), np.array([20.0, 0.0, -3.12, -3.0, 0.2, -0.5, 0.0, 0.0]),
              np.array([20.0, 0.0, 9.5, 0.0, 0.0, 0.0, 0.0, 0.0])):
        assert_costs_close(g.rollout_costs(s, 0, U, U, E), c.rollout_costs(s, 0, U, U, E), max_flip_frac=5e-3)


def test_control_cost_alpha_not_one(gpu_bound, orc):
    env = make_env("car")
    K, T = 256, 20
    g, c = pair(gpu_bound, orc, "gmppi", env, K, T, 1, alpha=0.5)
    rng = np.random.default_rng(6)
    A = rng.standard_normal((g.cs, g.cs))
    Sinv = np.linalg.inv(A @ A.T / g.cs + np.eye(g.cs))
    E, U, Uo = rng.standard_normal((g.cs, K)) * 0.3, rng.uniform(-0.5, 0.5, g.cs), rng.uniform(-0.5, 0.5, g.cs)
    assert_costs_close(g.rollout_costs(env.state, 0, U, Uo, E, Sinv), c.rollout_costs(env.state, 0, U, Uo, E, Sinv))
    # and through a full AIS plan (γ != 0 uses the engine's own triangular solves)
    for pol in ("mppi", "cemppi", "μΣaismppi"):
        g, c = pair(gpu_bound, orc, pol, env, K, T, 3, alpha=0.7, sigma_est="oas")
        Z = rng.standard_normal((g.cs, K, g.N))
        U0 = rng.uniform(-0.3, 0.3, g.cs)
        (cg, ug, ig), (cc, uc, ic) = g.plan(env.state, 0, U0, Z=Z), c.plan(env.state, 0, U0, Z=Z)
        assert ig == ic
        np.testing.assert_allclose(cg, cc, rtol=CONTROL_RTOL, atol=1e-8)
        np.testing.assert_allclose(ug, uc, rtol=CONTROL_RTOL, atol=1e-8)


CONFIGS = [  # BASELINE.json configs 1-4 (5 is the sharded one, see test_sharded.py) at oracle-friendly K
    ("mc", 1, "mppi", 20, 15, 1, dict(lam=0.1)),
    ("car", 1, "cemppi", 150, 50, 10, dict(lam=10.0, sigma_est="ss")),
    ("car", 1, "μΣaismppi", 4096, 50, 5, dict(lam=10.0, lam_ais=20.0)),
    ("car", 3, "cmamppi", 375, 50, 10, dict(lam=10.0)),
]


@pytest.mark.parametrize("envname,n_cars,policy,K,T,N,kw", CONFIGS)
def test_baseline_configs_vs_oracle(gpu_bound, orc, envname, n_cars, policy, K, T, N, kw):
    env = make_env(envname, n_cars)
    g, c = pair(gpu_bound, orc, policy, env, K, T, N, **kw)
    Z = np.random.Generator(np.random.Philox(key=K)).standard_normal((g.cs, K, g.N))
    U = np.zeros(g.cs)
    st = env.state
    for step in range(2):  # two consecutive control steps: also pins the roll of U
        (cg, ug, ig), (cc, uc, ic) = g.plan(st, step, U, Z=Z), c.plan(st, step, U, Z=Z)
        fg, fc = g.fetch(), c.fetch()
        assert ig == ic
        flips = assert_costs_close(fg["costs"], fc["costs"])
        np.testing.assert_allclose(cg, cc, rtol=CONTROL_RTOL, atol=1e-8)
        np.testing.assert_allclose(ug, uc, rtol=CONTROL_RTOL, atol=1e-8)
        Sg, Ug = g.fetch_proposal()
        Sc, Uc = c.fetch_proposal()
        np.testing.assert_allclose(Ug, Uc, rtol=1e-7, atol=1e-9)
        print(f"{policy} K={K} step {step}: |Δcontrol|={np.max(np.abs(cg - cc)):.2e} |ΔU|={np.max(np.abs(ug - uc)):.2e} flips={flips}")
        U = uc
        st, _, _, _ = c.env_step(st, cc, step)


@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("sigma_est", ["mle", "ss"])
def test_every_policy_vs_oracle(gpu_bound, orc, policy, sigma_est):
    if sigma_est == "ss" and policy != "cemppi":
        pytest.skip("Σ_est only applies to :cemppi")
    env = make_env("car")
    K, T, N = 512, 30, 6
    g, c = pair(gpu_bound, orc, policy, env, K, T, N, sigma_est=sigma_est)
    rng = np.random.Generator(np.random.Philox(key=77))
    Z, u = rng.standard_normal((g.cs, K, g.N)), rng.uniform(size=(K, max(g.N - 1, 1)))
    st = synthetic_states()[4]
    U = rng.uniform(-0.2, 0.2, g.cs)
    (cg, ug, ig), (cc, uc, ic) = g.plan(st, 0, U, Z=Z, resample_u=u), c.plan(st, 0, U, Z=Z, resample_u=u)
    assert ig == ic
    fg, fc = g.fetch(E=True), c.fetch(E=True)
    assert_costs_close(fg["costs"], fc["costs"])
    np.testing.assert_allclose(fg["E"], fc["E"], rtol=1e-6, atol=1e-9)  # E .+ (U − U_orig), POL:468
    np.testing.assert_allclose(cg, cc, rtol=CONTROL_RTOL, atol=1e-8)
    np.testing.assert_allclose(ug, uc, rtol=CONTROL_RTOL, atol=1e-8)
    if policy == "cemppi" and sigma_est == "ss":
        assert abs(g.last_shrinkage() - c.last_shrinkage()) < 1e-8


@pytest.mark.parametrize("method", ["mle", "lw", "ss", "rblw", "oas"])
def test_cov_estimators_vs_oracle(gpu_bound, orc, method):
    env = make_env("car")
    g, c = pair(gpu_bound, orc, "cemppi", env, 1024, 50, 2)
    rng = np.random.default_rng(3)
    for n in (30, 205, 1000):
        X = rng.standard_normal((100, n)) * rng.uniform(0.1, 2.0, (100, 1)) + rng.standard_normal((100, 1))
        (mg, Sg), (mc_, Sc) = g.cov_estimate(X, method), c.cov_estimate(X, method)
        np.testing.assert_allclose(mg, mc_, rtol=1e-12, atol=1e-13)
        # λ̂ of the common-variance targets divides by tr(S²) − tr²(S)/p (cancellation): summation-order
        # differences of ~1e-16 in S show up amplified there
        lam_tol = 1e-8 if method in ("rblw", "oas") else 1e-10
        np.testing.assert_allclose(Sg, Sc, rtol=100 * lam_tol, atol=1e-12)
        assert abs(g.last_shrinkage() - c.last_shrinkage()) < lam_tol
    w = rng.uniform(size=1000)
    np.testing.assert_allclose(g.cov_estimate(X, "mle", w=w)[1], c.cov_estimate(X, "mle", w=w)[1], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g.cov_estimate(X, "mle", corrected=True)[1], c.cov_estimate(X, "mle", corrected=True)[1],
                               rtol=1e-9, atol=1e-12)


def test_linear_algebra_vs_numpy(gpu_bound):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    rng = np.random.default_rng(1)
    # cs of BASELINE configs 1, 2, 4; the register-tile sizes (16, 64, 112, 160); the blocked kernel (161..430: full and
    # partial last panels, one row / one column below a panel) and the unblocked fallback beyond it
    for n in (1, 2, 15, 16, 17, 64, 100, 112, 113, 160, 161, 192, 193, 200, 257, 300, 430, 431):
        A = rng.standard_normal((n, 2 * n))
        S = A @ A.T / (2 * n) + 0.05 * np.eye(n)
        np.testing.assert_allclose(g.cholesky(S), np.linalg.cholesky(S), rtol=1e-10, atol=1e-12, err_msg=f"n={n}")
        if n in (15, 100, 300):
            C = g.inv_sqrt(S)
            np.testing.assert_allclose(C @ C @ S, np.eye(n), atol=1e-9)
    for n in (200, 300):  # a failing pivot in the first and in a later panel of the blocked kernel
        for bad in (3, n - 5):
            S = np.eye(n)
            S[bad, bad] = -1.0
            with pytest.raises(EngineError) as ei:
                g.cholesky(S)
            assert ei.value.code == -4
    with pytest.raises(EngineError) as ei:
        g.cholesky(np.diag([1.0, -1.0, 2.0]))
    assert ei.value.code == -4  # PosDefException
    with pytest.raises(EngineError) as ei:
        g.inv_sqrt(np.diag([1.0, -1.0, 2.0]))
    assert ei.value.code == -4


def test_sortperm_bit_exact(gpu_bound, orc):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    rng = np.random.default_rng(0)
    for n in (1, 7, 150, 2049, 65536, 300001):
        x = np.round(rng.normal(0, 1000, n), 1)  # many exact ties
        x[::13] = 0.0
        x[::17] = -0.0
        x[::19] = 1e6
        ref = orc.sortperm(x)
        assert np.array_equal(g.sortperm(x), ref)
        assert np.array_equal(ref, julia_sortperm(x))


def test_weights_vs_oracle(gpu_bound, orc):
    env = make_env("car")
    g, c = pair(gpu_bound, orc, "gmppi", env, 32, 50, 1)
    rng = np.random.default_rng(2)
    for K in (20, 150, 4097, 65536):
        costs = rng.normal(-800, 300, K)
        costs[::11] += 1e6
        for lam in (0.1, 10.0, 20.0):
            wg, wc = g.weights(costs, lam), c.weights(costs, lam)
            np.testing.assert_allclose(wg, wc, rtol=1e-12, atol=1e-300)
            assert abs(wg.sum() - 1) < 1e-12


def test_philox_normals_vs_oracle(gpu_bound, orc):
    env = make_env("car")
    for K, T in ((150, 50), (1000, 7)):
        g, c = pair(gpu_bound, orc, "gmppi", env, K, T, 1)
        for e in (g, c):
            e.seed(0xDEADBEEFCAFE)
        for step, it in ((0, 0), (3, 2), (70000, 9)):
            np.testing.assert_allclose(g.sample_normals(step, it), c.sample_normals(step, it), rtol=1e-12, atol=1e-14)
    mc = make_env("mc")
    g, c = pair(gpu_bound, orc, "gmppi", mc, 20, 15, 1, lam=0.1)  # odd cs: last pair half-used
    np.testing.assert_allclose(g.sample_normals(1, 0), c.sample_normals(1, 0), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("policy", ["cemppi", "pmcmppi", "cmamppi"])
def test_device_rng_plan_vs_oracle(gpu_bound, orc, policy):
    """Throughput mode: both sides draw from their own restatement of the same Philox stream."""
    env = make_env("car")
    K, T, N = 256, 20, 4
    g, c = pair(gpu_bound, orc, policy, env, K, T, N, sigma_est="ss")
    U = np.zeros(g.cs)
    for e in (g, c):
        e.seed(42)
    st = env.state
    for step in range(3):
        (cg, ug, ig), (cc, uc, ic) = g.plan(st, step, U), c.plan(st, step, U)
        assert ig == ic
        np.testing.assert_allclose(cg, cc, rtol=CONTROL_RTOL, atol=1e-8)
        np.testing.assert_allclose(ug, uc, rtol=CONTROL_RTOL, atol=1e-8)
        U = uc
        st, _, _, _ = c.env_step(st, cc, step)


def test_env_step_and_reward_vs_oracle(gpu_bound, orc):
    for envname, n_cars in (("car", 1), ("car", 3), ("mc", 1)):
        env = make_env(envname, n_cars)
        g, c = pair(gpu_bound, orc, "gmppi", env, 32, 5, 1, lam=1.0)
        rng = np.random.default_rng(4)
        sg = sc = env.state.copy()
        tg = tc = 0
        for i in range(40):
            a = rng.uniform(-1, 1, g.as_)
            sg, tg, rg, dg = g.env_step(sg, a, tg)
            sc, tc, rc, dc = c.env_step(sc, a, tc)
            np.testing.assert_allclose(sg, sc, rtol=1e-9, atol=1e-10)
            assert (tg, dg) == (tc, dc) and abs(rg - rc) <= 1e-9 * max(1, abs(rc))
            assert abs(g.env_reward(sg, dg) - rc) <= 1e-9 * max(1, abs(rc))


def test_early_stop_is_reproduced(gpu_bound, orc):
    """POL:459-461: with zero noise every elite cost is identical -> break at n = 1, before any update."""
    env = make_env("car")
    K, T, N = 64, 10, 5
    for pol in ("cemppi", "cmamppi"):
        g, c = pair(gpu_bound, orc, pol, env, K, T, N, sigma_est="mle")
        Z = np.zeros((g.cs, K, N))
        U = np.full(g.cs, 0.05)
        (cg, ug, ig), (cc, uc, ic) = g.plan(env.state, 0, U, Z=Z), c.plan(env.state, 0, U, Z=Z)
        assert ig == ic == 1
        np.testing.assert_allclose(cg, cc, rtol=1e-12)
        np.testing.assert_allclose(ug, uc, rtol=1e-12)
    # with the break disabled the CE loop continues on Σ′ = 1e-8 I and runs all N iterations
    g2 = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="mle", early_stop=False)), env, "cemppi")
    assert g2.plan(env.state, 0, U, Z=np.zeros((g2.cs, K, N)))[2] == N


def test_errors_mirror_the_reference(gpu_bound):
    env = make_env("car")
    g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, 64, 10, 3)), env, "cemppi")
    with pytest.raises(EngineError) as ei:
        g.set_sigma(np.eye(3))
    assert ei.value.code == -1 and "Covariance matrix size problem" in str(ei.value)
    S = np.eye(20)
    S[0, 0] = -1.0
    g.set_sigma(S)
    with pytest.raises(EngineError) as ei:
        g.plan(env.state, 0, np.zeros(20))
    assert ei.value.code == -4 and "PosDefException" in str(ei.value)
    with pytest.raises(EngineError):
        Engine(gpu_bound, **engine_kwargs("cemppi", env, 64, 10, 3, device=99))
    bare = Engine(gpu_bound, **engine_kwargs("gmppi", env, 64, 10))
    with pytest.raises(EngineError, match="environment not set"):
        bare.plan(env.state, 0, np.zeros(20))


def test_trajectory_logger(gpu_bound, orc):
    env = make_env("car")
    K, T = 48, 12
    g, c = pair(gpu_bound, orc, "gmppi", env, K, T, 1, log_trajectories=True)
    Z = np.random.default_rng(0).standard_normal((g.cs, K, 1))
    g.plan(env.state, 0, np.zeros(g.cs), Z=Z)
    c.plan(env.state, 0, np.zeros(g.cs), Z=Z)
    tg, tc = g.fetch(traj=True)["traj"], c.fetch(traj=True)["traj"]
    assert tg.shape == (K, T, 8)
    np.testing.assert_allclose(tg, tc, rtol=1e-9, atol=1e-10)


def test_resident_loop_equals_host_loop(gpu_bound):
    env = make_env("car")
    K, T, N = 1024, 50, 4
    a = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N)), env, "cemppi")
    b = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N)), env, "cemppi")
    for e in (a, b):
        e.seed(9)
    st, U, t = env.state.copy(), np.zeros(a.cs), 0
    a.resident_reset(st, 0, U)
    for step in range(5):
        a.resident_plan(True)
        ctrl, U, its = b.plan(st, t, U)
        st, t, _, _ = b.env_step(st, ctrl, t)
    sa, Ua, ca, _ = a.resident_read()
    assert a.resident_total_its() == 5 * N
    np.testing.assert_allclose(ca, ctrl, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(Ua, U, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sa, st, rtol=1e-12, atol=1e-14)


# ---------------------------------------------------------------------------------------------------
# 3. properties at BASELINE sizes (no oracle: K = 65 536)
# ---------------------------------------------------------------------------------------------------
def test_properties_at_full_size(gpu_bound):
    env = make_env("car")
    K, T, N = 65536, 50, 10
    g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="ss")), env, "cemppi")
    g.seed(1)
    U = np.zeros(g.cs)
    ctrl, U2, its = g.plan(env.state, 0, U)
    f = g.fetch(E=True)
    w, E, costs = f["weights"], f["E"], f["costs"]
    assert its == N and abs(w.sum() - 1) < 1e-10 and np.all(w >= 0) and w.argmax() == costs.argmin()
    wc = U + E @ w  # shift identity POL:468 + POL:226-231
    np.testing.assert_allclose(ctrl, np.clip(wc[:2], -1, 1), rtol=1e-9)
    np.testing.assert_allclose(U2[:-2], wc[2:], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(U2[-2:], U[-2:])
    # elite selection at full size: the device permutation is the stable arg-sort of the fetched costs
    perm = g.sortperm(costs)
    assert np.array_equal(perm, np.argsort(costs, kind="stable"))
    # rollouts are a pure function of (state, U + E[:,k]): re-evaluating a permuted subset reproduces the costs bitwise
    Ssig, Ulast = g.fetch_proposal()
    idx = np.random.default_rng(0).permutation(K)[:4096]
    h = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, 4096, T)), env, "gmppi")
    c2 = h.rollout_costs(env.state, 0, U, U, E[:, idx])  # E is already shifted by (U_last − U)
    np.testing.assert_allclose(c2, costs[idx], rtol=1e-12)
    # determinism: same seed, same step -> identical control
    g.seed(1)
    ctrl_b, U2_b, _ = g.plan(env.state, 0, U)
    assert np.array_equal(ctrl, ctrl_b) and np.array_equal(U2, U2_b)
    # the device generator is N(0, I)
    z = g.sample_normals(5, 3)
    assert abs(z.mean()) < 2e-3 and abs(z.var() - 1) < 2e-3


def test_closed_loop_lap_engine_vs_oracle(gpu_bound, orc):
    """Closed loop (EXC:203-281) side by side: same injected noise each step, oracle env stepping."""
    env = make_env("car")
    K, T, N = 150, 50, 10
    g, c = pair(gpu_bound, orc, "cemppi", env, K, T, N, sigma_est="ss")
    st, Ug, Uc = env.state.copy(), np.zeros(g.cs), np.zeros(g.cs)
    worst = 0.0
    for step in range(25):
        Z = np.random.Generator(np.random.Philox(key=1000 + step)).standard_normal((g.cs, K, N))
        (cg, Ug, ig), (cc, Uc, ic) = g.plan(st, step, Uc, Z=Z), c.plan(st, step, Uc, Z=Z)
        assert ig == ic
        worst = max(worst, float(np.max(np.abs(cg - cc))), float(np.max(np.abs(Ug - Uc))))
        st, _, rew, _ = c.env_step(st, cc, step)
        assert rew > -4000  # stays on the track
    print(f"closed loop, 25 steps: worst |Δ| = {worst:.2e}")
    assert worst < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("policy,sigma_est", [("cemppi", "ss"), ("cemppi", "lw"), ("cemppi", "oas"), ("cemppi", "mle"),
                                              ("μΣaismppi", "mle"), ("pmcmppi", "mle"), ("imppi", "mle")])
def test_single_cta_moment_chain_equals_the_kernel_chain(gpu_bound, policy, sigma_est):
    """n <= 512 columns take ONE kernel for count / mean / scatter / shrinkage / Σ′ (stats.cu: moments_small_kernel);
    the multi-kernel chain stays selectable ("moments_small" = 0): same proposal, same control."""
    env = make_env("car")
    K, T, N = 256, 20, 4
    rng = np.random.Generator(np.random.Philox(key=5))
    Z, u = rng.standard_normal((2 * T, K, N)), rng.uniform(size=(K, N - 1))
    out = []
    for flag in (1, 0):
        g = configure(Engine(gpu_bound, **engine_kwargs(policy, env, K, T, N, sigma_est=sigma_est)), env, policy)
        g.set_option("moments_small", flag)
        ctrl, U2, its = g.plan(synthetic_states()[1], 0, np.zeros(g.cs), Z=Z, resample_u=u)
        S, Ul = g.fetch_proposal()
        out.append((ctrl, U2, its, S, Ul, g.last_shrinkage(), g.launch_count()))
    a, b = out
    assert a[2] == b[2] and a[6] < b[6]  # same iterations, fewer launches
    np.testing.assert_allclose(a[3], b[3], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a[5], b[5], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("envname,K,T", [("car", 150, 50), ("car", 500, 50), ("car", 64, 7), ("mc", 20, 15), ("mc", 300, 15)])
@pytest.mark.parametrize("sigma_est", ["ss", "lw", "rblw", "oas", "mle"])
def test_fused_small_adaptation_equals_the_kernel_chain_and_the_oracle(gpu_bound, orc, envname, K, T, sigma_est):
    """The reference's own sizes (K = 150, cs = 100; MountainCar K = 20, cs = 15) run sort + early stop + elite moments +
    shrinkage + Cholesky of one :cemppi iteration as ONE single-CTA launch (csrc/small_adapt.cu); "ce_small_fused" = 0
    restores the kernel chain. Same iterations, proposal, λ̂ and control as the chain — and as the oracle."""
    env = make_env(envname)
    N = 5
    rng = np.random.Generator(np.random.Philox(key=K + T))
    cpu = configure(orc.engine(nthreads=4, **engine_kwargs("cemppi", env, K, T, N, sigma_est=sigma_est)), env, "cemppi")
    Z = rng.standard_normal((cpu.cs, K, N))
    U0 = rng.uniform(-0.1, 0.1, cpu.cs)
    st = synthetic_states()[2] if envname == "car" else env.state
    out = []
    for flag in (1, 0):
        g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est=sigma_est)), env, "cemppi")
        g.set_option("ce_small_fused", flag)
        assert int(g.get_option("ce_small_fused")) == flag
        ctrl, U2, its = g.plan(st, 0, U0, Z=Z)
        S, Ul = g.fetch_proposal()
        out.append((ctrl, U2, its, S, Ul, g.last_shrinkage(), g.launch_count()))
        g.close()
    a, b = out
    assert a[2] == b[2] and a[6] < b[6]  # same iterations, fewer launches
    np.testing.assert_allclose(a[3], b[3], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a[5], b[5], rtol=1e-9, atol=1e-12)
    cc, uc, ic = cpu.plan(st, 0, U0, Z=Z)
    Sc, Uc = cpu.fetch_proposal()
    assert a[2] == ic
    np.testing.assert_allclose(a[0], cc, rtol=1e-5, atol=1e-8)   # north-star tolerance
    np.testing.assert_allclose(a[1], uc, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(a[3], Sc, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(a[5], cpu.last_shrinkage(), rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
def test_fused_small_adaptation_early_stop_and_failure(gpu_bound):
    """Zero noise -> identical elite costs -> the fused kernel raises the stop flag and leaves pol.U / Σ′ untouched (POL:458-461);
    a Σ′ that is not positive definite surfaces as the reference's PosDefException with the AIS iteration in the message."""
    from mpopis_b200.engine import EngineError
    env = make_env("car")
    K, T, N = 150, 10, 5
    res = []
    for flag in (1, 0):
        g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="mle")), env, "cemppi")
        g.set_option("ce_small_fused", flag)
        res.append(g.plan(env.state, 0, np.full(g.cs, 0.05), Z=np.zeros((g.cs, K, N))))
        g.close()
    assert res[0][2] == res[1][2] == 1
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
    # rank-deficient scatter matrix (every elite identical in all but one coordinate) with :mle and NaN noise -> NaN pivot
    g = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, K, T, N, sigma_est="mle")), env, "cemppi")
    Z = np.random.default_rng(3).standard_normal((g.cs, K, N))
    Z[0, :, 0] = np.nan
    with pytest.raises(EngineError, match="PosDefException.*AIS iteration 2"):
        g.plan(env.state, 0, np.zeros(g.cs), Z=Z)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_cars,K", [(1, 2050), (1, 70), (2, 333), (3, 96), (4, 64)])
def test_tma_staged_noise_equals_register_prefetch(gpu_bound, n_cars, K):
    """"rollout_stage" = 1 brings the noise tile of every warp into shared memory with TMA bulk copies (per-warp ring,
    mbarrier completion) instead of prefetching it into registers: same arithmetic (the two template instances may
    contract FMAs differently, hence rounding-level agreement rather than bit equality) — including warps that are
    only partly filled (K not a multiple of 32)."""
    env = make_env("car", n_cars)
    T = 50
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, K, T, 1)), env, "gmppi")
    rng = np.random.default_rng(100 + n_cars)
    E = rng.standard_normal((g.cs, K)) * 0.3
    U = rng.uniform(-0.3, 0.3, g.cs)
    st = env.state
    g.set_option("rollout_stage", 0)
    c0 = g.rollout_costs(st, 0, U, U, E)
    g.set_option("rollout_stage", 1)
    c1 = g.rollout_costs(st, 0, U, U, E)
    assert_costs_close(c1, c0)


def test_target_config_one_step_vs_oracle(gpu_bound, orc):
    """The north-star configuration itself (K = 65 536, T = 50, :cemppi, Σ_est = :ss) against the oracle on injected
    noise — N = 3 AIS iterations keeps the CPU side to a few seconds. Exercises select.cu (13 107 elites), the gathered
    elite moments, the shrinkage estimate and the DMMA E = L·Z at the size the bench measures."""
    env = make_env("car")
    K, T, N = 65536, 50, 3
    g, c = pair(gpu_bound, orc, "cemppi", env, K, T, N, sigma_est="ss")
    c.b.set_threads(c.h, 16)
    rng = np.random.Generator(np.random.Philox(key=65536))
    Z = rng.standard_normal((g.cs, K, N))
    U = rng.uniform(-0.1, 0.1, g.cs)
    st = synthetic_states()[2]
    (cg, ug, ig), (cc, uc, ic) = g.plan(st, 0, U, Z=Z), c.plan(st, 0, U, Z=Z)
    assert ig == ic == N
    flips = assert_costs_close(g.fetch()["costs"], c.fetch()["costs"])
    np.testing.assert_allclose(cg, cc, rtol=CONTROL_RTOL, atol=1e-8)
    np.testing.assert_allclose(ug, uc, rtol=CONTROL_RTOL, atol=1e-8)
    Sg, Ug = g.fetch_proposal()
    Sc, Uc = c.fetch_proposal()
    np.testing.assert_allclose(Ug, Uc, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(Sg, Sc, rtol=1e-6, atol=1e-12)
    assert abs(g.last_shrinkage() - c.last_shrinkage()) < 1e-8
    print(f"K=65536: |Δcontrol|={np.max(np.abs(cg - cc)):.2e} |ΔU|={np.max(np.abs(ug - uc)):.2e} cost flips={flips}")


@pytest.mark.parametrize("n_cars,K,T", [(1, 4099, 50), (1, 70, 23), (1, 33, 7), (2, 333, 30), (3, 375, 50), (3, 64, 11)])
def test_warp_specialised_rollouts_equal_the_thread_per_rollout_kernel(gpu_bound, n_cars, K, T):
    """rollout_variant 4 (rollout_split.cu: velocity warps + pose/reward warps through a shared-memory ring) against
    variant 3 (one thread per rollout): the same mathematics with the velocity recurrence re-associated for a shorter
    dependent chain, so costs and the logged per-step states agree to accumulated rounding (chaotic dynamics over 500
    Euler sub-steps: 1e-9 with a handful of penalty-threshold flips), not bitwise. Ragged K (partly filled last warp /
    single velocity warp in the last CTA) and short horizons included."""
    env = make_env("car", n_cars)
    g = configure(Engine(gpu_bound, **engine_kwargs("gmppi", env, K, T, log_trajectories=True)), env, "gmppi")
    rng = np.random.Generator(np.random.Philox(key=K + T))
    E = rng.standard_normal((g.cs, K)) * np.tile(np.array([0.25, 0.32] * n_cars), T)[:, None]
    U = rng.uniform(-0.3, 0.3, g.cs)
    out = {}
    for variant, spin in ((3, 0), (4, 0), (5, 0), (5.5, 1)):  # 5.5: variant 5 handing over through spin counters
        g.set_option("rollout_variant", int(variant))
        g.set_option("rollout_spin", spin)
        out[variant] = (g.rollout_costs(env.state, 0, U, U, E), g.fetch(costs=False, weights=False, traj=True)["traj"])
    for variant in (4, 5, 5.5):  # 5 = two pose warps per CTA at 160 registers per thread
        r = rel(out[variant][0], out[3][0])
        assert (r > TIGHT).sum() <= max(1, K // 500), f"{(r > TIGHT).sum()} of {K} costs differ (max rel {r.max():.2e})"
        print(f"variant {variant} vs 3: median rel {np.median(r):.2e}, 99.9 % {np.quantile(r, 0.999):.2e}")
        tr = np.abs(out[variant][1] - out[3][1]) / np.maximum(1.0, np.abs(out[3][1]))
        assert np.quantile(tr, 0.999) < 1e-9, "trajectory logs differ"


def test_trial_replicas_equal_sequential_trials(gpu_bound):
    """§8f-2: `num_trials` independent trials (car_example.jl:170) as concurrent device-resident replicas — each
    trial its own handle / stream / Philox key — reproduce the same trials run one after the other, bit for bit."""
    from mpopis_b200.trials import run_trial_replicas

    def make(dev):
        env = make_env("car")
        e = configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, 150, 50, 10, sigma_est="ss", device=dev)), env, "cemppi")
        return env, e

    seeds = [11, 12, 13, 14, 15, 16]
    conc, _ = run_trial_replicas(make, 6, 8, seeds=seeds)
    seq, _ = run_trial_replicas(make, 6, 8, seeds=seeds, concurrency=1)
    for a, b in zip(conc, seq):
        assert np.array_equal(a["state"], b["state"]) and np.array_equal(a["U"], b["U"]) and a["its"] == b["its"]
        assert a["reward_sum"] == b["reward_sum"] and np.isfinite(a["reward_sum"])
    assert len({tuple(r["state"]) for r in conc}) == len(seeds), "different seeds give different trials"
