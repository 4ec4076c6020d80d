"""Pins the C oracle against (a) known-answer vectors that exist independently of the reference
(Random123's Philox KATs), (b) numpy/scipy for the linear algebra, and (c) an independent numpy
restatement of the reference code (oracle/np_mirror.py). The reference itself ships no tests or
golden vectors and Julia is not installed, so this is as pinned as the oracle can get here:
PARITY UNPINNED (see oracle/mpopis_oracle.h)."""
import numpy as np
import pytest
from conftest import configure, engine_kwargs, julia_sortperm, make_env, synthetic_states

from mpopis_b200 import _abi
from oracle import np_mirror as npm


# Random123 kat_vectors, philox4x32 with 10 rounds
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,expect", PHILOX_KAT)
def test_philox_known_answers(orc, ctr, key, expect):
    assert tuple(orc.philox4x32_10(ctr, key)) == expect


def test_philox_normals_are_standard_normal(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("gmppi", env, 4096, 50)), env, "gmppi")
    e.seed(3)
    z = e.sample_normals(0, 0)
    assert z.shape == (100, 4096)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z ** 4) - 3) < 0.1
    z2 = e.sample_normals(0, 1)
    assert abs(np.corrcoef(z.ravel(), z2.ravel())[0, 1]) < 0.01
    assert np.array_equal(z, e.sample_normals(0, 0))  # counter-based: reproducible


def test_sortperm_is_stable(orc):
    rng = np.random.default_rng(0)
    x = rng.integers(0, 50, 5000).astype(float)  # many ties
    x[::97] = -0.0
    x[5::97] = 0.0
    assert np.array_equal(orc.sortperm(x), julia_sortperm(x))


def test_cholesky_and_inv_sqrt(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    rng = np.random.default_rng(1)
    A = rng.standard_normal((40, 60))
    S = A @ A.T / 60 + 0.1 * np.eye(40)
    L = e.cholesky(S)
    np.testing.assert_allclose(L, np.linalg.cholesky(S), rtol=1e-12, atol=1e-13)
    C = e.inv_sqrt(S)
    np.testing.assert_allclose(C @ C @ S, np.eye(40), atol=1e-10)
    np.testing.assert_allclose(C, C.T, atol=1e-12)
    with pytest.raises(Exception):
        e.cholesky(np.diag([1.0, -1.0]))


@pytest.mark.parametrize("method", ["mle", "lw", "ss", "rblw", "oas"])
@pytest.mark.parametrize("n", [30, 400])
def test_cov_estimators_match_numpy_mirror(orc, method, n):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("cemppi", env, 512, 50)), env, "cemppi")
    rng = np.random.default_rng(n)
    X = rng.standard_normal((100, n)) * rng.uniform(0.1, 2.0, (100, 1)) + rng.standard_normal((100, 1))
    mu, S = e.cov_estimate(X, method)
    mu_r, S_r, lam_r = npm.cov_estimate(X, method)
    np.testing.assert_allclose(mu, mu_r, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(S, S_r, rtol=1e-10, atol=1e-12)
    assert abs(e.last_shrinkage() - lam_r) < 1e-10
    if method == "mle":
        np.testing.assert_allclose(S, np.cov(X, bias=True), rtol=1e-10, atol=1e-12)
    else:
        assert 0.0 <= lam_r <= 1.0


def test_weighted_and_corrected_moments(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("cemppi", env, 512, 50)), env, "cemppi")
    rng = np.random.default_rng(5)
    X = rng.standard_normal((100, 300))
    w = rng.uniform(size=300)
    mu, S = e.cov_estimate(X, "mle", w=w)
    np.testing.assert_allclose(mu, X @ w / w.sum(), rtol=1e-12)
    np.testing.assert_allclose(S, np.cov(X, aweights=w, bias=True), rtol=1e-10, atol=1e-13)  # StatsBase corrected=false
    mu, S = e.cov_estimate(X, "mle", corrected=True)
    np.testing.assert_allclose(S, np.cov(X), rtol=1e-10, atol=1e-13)  # mean_and_cov(E′, 2): divisor n−1


def test_weights_properties(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    rng = np.random.default_rng(2)
    c = rng.normal(0, 300, 1000)
    w = e.weights(c, 10.0)
    np.testing.assert_allclose(w, npm.weights(c, 10.0), rtol=1e-13)
    assert abs(w.sum() - 1) < 1e-12 and w.argmax() == c.argmin()
    np.testing.assert_allclose(e.weights(c + 1234.5, 10.0), w, rtol=1e-9)  # ρ-shift invariance


def test_car_step_matches_numpy_mirror(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    P = env.params.as_array()
    rng = np.random.default_rng(3)
    states = synthetic_states() + [env.state.copy(), np.array([0, 0, 3.1, -5.0, 0.5, 0.1, 0.1, 0.0]),
                                   np.array([10, 5, -3.0, 0.0, 0.0, 0.0, 0.0, 0.0])]
    for s in states:
        for a in [(-1, -1), (1, 1), (0.3, -0.7), (-0.2, 0.9), (0, 0)] + [tuple(rng.uniform(-1, 1, 2))]:
            s1, t1, rew, _ = e.env_step(s, np.array(a, float), 0)
            ref = npm.car_step(P, env.dt, env.δt, s, a)
            np.testing.assert_allclose(s1, ref, rtol=1e-12, atol=1e-12)
            assert t1 == 1
            assert abs(rew - npm.car_reward(P, env.track.xs, env.track.ys, env.track.ws, ref)) <= 1e-9 * max(1, abs(rew))


def test_within_track_matches_numpy_mirror_exactly(orc):
    env = make_env("car")
    e = configure(orc.engine(**engine_kwargs("gmppi", env, 32, 50)), env, "gmppi")
    tx, ty, tw = env.track.xs, env.track.ys, env.track.ws
    assert len(tx) == 48  # 946-point curve.csv sub-sampled 1:20:end
    gx, gy = np.meshgrid(np.linspace(1, 255, 60), np.linspace(-156, 141, 60))  # SURVEY App. G-2 grid
    pos = np.stack([gx.ravel(), gy.ravel()], axis=1)
    idx, idx2, dist, within = e.track_query(pos)
    for q in range(0, len(pos), 7):
        i, j, d, w = npm.within_track(tx, ty, tw, pos[q])
        assert (idx[q], idx2[q], bool(within[q])) == (i, j, bool(w))
        assert abs(dist[q] - d) <= 1e-12 * max(1, d)
    # on a sampled point: distance 0, inside
    i0, _, d0, w0 = e.track_query(np.array([[tx[5], ty[5]]]))
    assert i0[0] == 5 and d0[0] < 1e-12 and w0[0]


@pytest.mark.parametrize("n_cars", [1, 3])
def test_rollout_costs_match_numpy_mirror(orc, n_cars):
    env = make_env("car", n_cars)
    K, T = 6, 12
    e = configure(orc.engine(**engine_kwargs("gmppi", env, K, T)), env, "gmppi")
    rng = np.random.default_rng(4)
    E = rng.standard_normal((e.cs, K)) * 0.4
    U = rng.uniform(-0.5, 0.5, e.cs)
    costs = e.rollout_costs(env.state, 0, U, U, E)
    Ps = [p.as_array() for p in env.car_params] if n_cars > 1 else [env.params.as_array()]
    for k in range(K):
        ref = npm.rollout_cost(Ps, env.dt, env.δt, env.track.xs, env.track.ys, env.track.ws, env.state, U + E[:, k])
        assert abs(costs[k] - ref) <= 1e-9 * max(1, abs(ref))


def test_control_cost_term(orc):
    """γ U_orig' Σ⁻¹ (V − U_orig) with α != 1 (POL:272)."""
    env = make_env("car")
    K, T = 8, 5
    kw = engine_kwargs("gmppi", env, K, T, alpha=0.6)
    e = configure(orc.engine(**kw), env, "gmppi")
    e0 = configure(orc.engine(**engine_kwargs("gmppi", env, K, T)), env, "gmppi")
    rng = np.random.default_rng(6)
    A = rng.standard_normal((e.cs, e.cs))
    Sinv = np.linalg.inv(A @ A.T + np.eye(e.cs))
    E, U, Uo = rng.standard_normal((e.cs, K)) * 0.3, rng.uniform(-0.5, 0.5, e.cs), rng.uniform(-0.5, 0.5, e.cs)
    c = e.rollout_costs(env.state, 0, U, Uo, E, Sinv)
    c0 = e0.rollout_costs(env.state, 0, U, Uo, E)
    gamma = 10.0 * (1 - 0.6)
    np.testing.assert_allclose(c - c0, gamma * Uo @ Sinv @ (U[:, None] + E - Uo[:, None]), rtol=1e-9, atol=1e-9)


def test_mountaincar_rollout_matches_numpy_mirror(orc):
    env = make_env("mc")
    K, T = 10, 15
    e = configure(orc.engine(**engine_kwargs("gmppi", env, K, T, lam=0.1)), env, "gmppi")
    rng = np.random.default_rng(7)
    E = rng.standard_normal((T, K)) * 1.2
    U = np.zeros(T)
    for t0 in (0, 190):  # env.t is inherited by the copies: max_steps reached inside the horizon
        costs = e.rollout_costs(env.state, t0, U, U, E)
        for k in range(K):
            ref = npm.mountaincar_rollout(env.params.as_array(), env.params.max_steps, env.state, t0, E[:, k])
            assert abs(costs[k] - ref) <= 1e-12 * max(1, abs(ref))


def test_cemppi_plan_matches_numpy_mirror(orc):
    env = make_env("car")
    K, T, N = 24, 8, 4
    for method in ("mle", "ss"):
        e = configure(orc.engine(**engine_kwargs("cemppi", env, K, T, N, sigma_est=method)), env, "cemppi")
        rng = np.random.default_rng(8)
        Z = rng.standard_normal((e.cs, K, N))
        U = rng.uniform(-0.2, 0.2, e.cs)
        ctrl, U2, its = e.plan(env.state, 0, U, Z=Z)
        out = e.fetch()
        rc, rU, rits, rcosts, rw = npm.ce_plan([env.params.as_array()], env.dt, env.δt, env.track.xs, env.track.ys,
                                               env.track.ws, env.state, U, block_diagm_cs(e.cs), Z, 10.0, N, 0.8, method)
        assert its == rits
        np.testing.assert_allclose(out["costs"], rcosts, rtol=1e-8)
        np.testing.assert_allclose(ctrl, rc, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(U2, rU, rtol=1e-7, atol=1e-9)


def block_diagm_cs(cs):
    return np.diag(np.tile([0.0625, 0.1], cs // 2))


def test_mppi_equals_gmppi_with_block_diagonal_sigma(orc):
    """:mppi samples T independent N(0, Σ_as) — identical to :gmppi with block_diagm(Σ_as, T)."""
    env = make_env("car")
    K, T = 40, 10
    rng = np.random.default_rng(9)
    Z = rng.standard_normal((2 * T, K, 1))
    cov = np.array([[0.0625, 0.02], [0.02, 0.1]])
    res = []
    for pol in ("mppi", "gmppi"):
        e = configure(orc.engine(**engine_kwargs(pol, env, K, T, alpha=0.8)), env, pol, cov=cov)
        res.append(e.plan(env.state, 0, np.full(2 * T, 0.1), Z=Z))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-10)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-10, atol=1e-12)


def test_roll_keeps_last_action_block(orc):
    """get_controls_roll_U! with pol.U aliasing pol.params.U₀: shift left by `as`,
    the last `as` entries keep their values."""
    env = make_env("car")
    K, T = 16, 6
    e = configure(orc.engine(**engine_kwargs("gmppi", env, K, T)), env, "gmppi")
    U = np.arange(12, dtype=float) / 100
    Z = np.zeros((12, K, 1))  # no noise -> weighted controls == U
    ctrl, U2, _ = e.plan(env.state, 0, U, Z=Z)
    np.testing.assert_allclose(ctrl, U[:2])
    np.testing.assert_allclose(U2[:-2], U[2:])
    np.testing.assert_allclose(U2[-2:], U[-2:])


def test_shift_identity_and_its(orc):
    """E .+ (pol.U − U_orig) (POL:468): control == clamp(U_orig + Σ w (E + ΔU))[1:as]."""
    env = make_env("car")
    K, T, N = 64, 10, 5
    for pol in ("imppi", "μaismppi", "μΣaismppi", "pmcmppi", "cemppi", "cmamppi"):
        e = configure(orc.engine(**engine_kwargs(pol, env, K, T, N, sigma_est="mle")), env, pol)
        rng = np.random.default_rng(10)
        Z = rng.standard_normal((e.cs, K, N))
        u = rng.uniform(size=(K, N - 1))
        U = np.zeros(e.cs)
        ctrl, U2, its = e.plan(env.state, 0, U, Z=Z, resample_u=u)
        out = e.fetch(E=True)
        assert its == N
        wc = U + out["E"] @ out["weights"]
        np.testing.assert_allclose(ctrl, np.clip(wc[:2], -1, 1), rtol=1e-10)
        np.testing.assert_allclose(U2[:-2], wc[2:], rtol=1e-10, atol=1e-14)
        assert abs(out["weights"].sum() - 1) < 1e-12


def test_oas_and_rblw_against_their_defining_iterations(orc):
    """Independent pin for two of the shrinkage estimators the reference takes from CovarianceEstimation.jl (POL:421-423):
    Chen, Wiesel, Eldar & Hero (2010) DEFINE the OAS intensity as the limit of the iteration
        ρ_{j+1} = [(1 − 2/p) tr(Σ_j S) + tr²(Σ_j)] / [(n + 1 − 2/p) tr(Σ_j S) + (1 − n/p) tr²(Σ_j)],  Σ_j = (1 − ρ_j) S + ρ_j F
    (its closed form is what the oracle and the engine implement) and the RBLW intensity by their eq. (17). Running the
    iteration here checks the closed form without sharing its code. sklearn's OAS is NOT usable as a cross-check: it
    documents a deliberately different formula (0.3085 vs 0.3060 on this matrix). :lw / :ss (shrinkage towards diag(S)) are checked against the
    element-by-element definition of Schäfer & Strimmer in the next test; the reference's own numbers for all four still
    need tests/golden/julia_v1.json (julia/make_fixtures.jl)."""
    rng = np.random.default_rng(0)
    for p, n in ((20, 60), (100, 30), (100, 819)):
        A = rng.normal(size=(p, p))
        X = A @ rng.normal(size=(p, n))
        e = orc.engine(policy="cemppi", env=_abi.ENV_MOUNTAIN_CAR, num_samples=max(n, 2), horizon=p, opt_its=2, lam=1.0)
        Xc = X - X.mean(axis=1, keepdims=True)
        S = Xc @ Xc.T / n
        F = np.trace(S) / p * np.eye(p)
        Sig = S.copy()
        for _ in range(500):
            num = (1 - 2 / p) * np.trace(Sig @ S) + np.trace(Sig) ** 2
            den = (n + 1 - 2 / p) * np.trace(Sig @ S) + (1 - n / p) * np.trace(Sig) ** 2
            rho = min(num / den, 1.0)
            Sig = (1 - rho) * S + rho * F
        _, S_oas = e.cov_estimate(X, "oas")
        assert abs(e.last_shrinkage() - rho) < 1e-10
        np.testing.assert_allclose(S_oas, Sig, rtol=1e-9, atol=1e-12)
        tr, tr2 = np.trace(S), np.sum(S * S)
        rho_rblw = min(((n - 2) / n * tr2 + tr * tr) / ((n + 2) * (tr2 - tr * tr / p)), 1.0)  # eq. (17)
        _, S_rblw = e.cov_estimate(X, "rblw")
        assert abs(e.last_shrinkage() - rho_rblw) < 1e-12
        np.testing.assert_allclose(S_rblw, (1 - rho_rblw) * S + rho_rblw * F, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("method", ["lw", "ss"])
def test_diagonal_target_shrinkage_against_schaefer_strimmer_directly(orc, method):
    """Independent check of the O(n·p) shrinkage statistic behind :lw / :ss (the reference's default
    estimator for the car is :ss). Schäfer & Strimmer (2005), target D "diagonal, unequal variances", define
        λ* = Σ_{i≠j} Var^(s_ij) / Σ_{i≠j} s_ij²,   Var^(s_ij) = n/(n−1)³ Σ_k (w_kij − w̄_ij)²,   s_ij = n/(n−1) w̄_ij,
        w_kij = (x_ki − x̄_i)(x_kj − x̄_j)
    (:ss — their own variant — applies it to the correlations, i.e. to data standardised by the sample standard
    deviations; both shrink S_mle towards diag(S_mle)). This test evaluates the definition element by element, O(n·p²),
    with no algebra shared with the oracle's closed form Σ_k[(Σ_i z²)² − Σ_i z⁴]. It pins the formula as published; the
    reference's own CovarianceEstimation.jl numbers still need julia/make_fixtures.jl."""
    rng = np.random.default_rng(11)
    for p, n in ((6, 40), (30, 30), (100, 30), (40, 819)):
        A = rng.normal(size=(p, p))
        X = A @ rng.normal(size=(p, n)) + rng.normal(size=(p, 1))
        Xc = X - X.mean(axis=1, keepdims=True)
        S = Xc @ Xc.T / n                                   # the MLE the reference shrinks (corrected = false)
        Z = Xc / np.sqrt(np.diag(S))[:, None] if method == "ss" else Xc
        W = Z[:, None, :] * Z[None, :, :]                   # w_kij, shape p x p x n
        wbar = W.mean(axis=2)
        var_s = n / (n - 1) ** 3 * ((W - wbar[:, :, None]) ** 2).sum(axis=2)
        s = n / (n - 1) * wbar
        off = ~np.eye(p, dtype=bool)
        lam = float(np.clip(var_s[off].sum() / (s[off] ** 2).sum(), 0.0, 1.0))
        e = orc.engine(policy="cemppi", env=_abi.ENV_MOUNTAIN_CAR, num_samples=max(n, 2), horizon=p, opt_its=2, lam=1.0)
        _, S_hat = e.cov_estimate(X, method)
        assert abs(e.last_shrinkage() - lam) < 1e-10 * max(1.0, lam), (p, n, e.last_shrinkage(), lam)
        target = (1 - lam) * S + lam * np.diag(np.diag(S))
        np.testing.assert_allclose(S_hat, target, rtol=1e-9, atol=1e-12)
