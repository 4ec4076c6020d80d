"""Reference-pinned golden vectors (closes "parity unpinned" when present).

tests/golden/julia_v1.json is written by julia/make_fixtures.jl from a STOCK MPOPIS checkout (it calls the reference's
own functions; nothing of this repo). Julia is not installed in the build image, so the file cannot be produced here:
until a maintainer commits it, every test in this module SKIPS LOUDLY and the oracle remains pinned only to published
formulas. With the file present the CPU oracle (and, under -m gpu, the CUDA engine) must reproduce the reference's
numbers: integer columns bit-exact, floating point to 1e-9 (third-party estimators) / 1e-5 (control, north star)."""
import json
from pathlib import Path

import numpy as np
import pytest
from conftest import configure, engine_kwargs, make_env

from mpopis_b200 import _abi

FIX = Path(__file__).resolve().parent / "golden" / "julia_v1.json"
pytestmark = pytest.mark.skipif(
    not FIX.exists(),
    reason="PARITY UNPINNED: tests/golden/julia_v1.json is absent — run `julia julia/make_fixtures.jl "
           "tests/golden/julia_v1.json` from a MPOPIS checkout and commit the result")


def arr(o):
    """Decode make_fixtures.jl's JSON: {"dims", "data"} is a column-major array; "NaN"/"Inf" strings are floats."""
    if isinstance(o, dict) and set(o) == {"dims", "data"}:
        return np.array([float(v) for v in o["data"]], dtype=np.float64).reshape(o["dims"], order="F")
    if isinstance(o, list):
        return np.array([float(v) if not isinstance(v, bool) else v for v in o])
    return o


@pytest.fixture(scope="module")
def fx():
    return json.loads(FIX.read_text())


def backends(orc):
    out = [("oracle", lambda **kw: orc.engine(nthreads=4, **kw))]
    try:
        import torch
        if torch.cuda.is_available():
            from mpopis_b200 import _lib
            from mpopis_b200.engine import Engine
            out.append(("cuda", lambda **kw: Engine(_lib.product(), **kw)))
    except Exception:
        pass
    return out


def test_car_step_known_answers(fx, orc):
    env = make_env("car")
    for name, mk in backends(orc):
        e = configure(mk(**engine_kwargs("gmppi", env, 32, 5)), env, "gmppi")
        for r in fx["car_step"]:
            s, t = arr(r["state0"]), 0
            s1, t, rew, _ = e.env_step(s, arr(r["action"]), t)
            np.testing.assert_allclose(s1, arr(r["state1"]), rtol=1e-9, atol=1e-10, err_msg=name)
            assert abs(rew - r["reward1"]) <= 1e-9 * max(1.0, abs(r["reward1"]))
            for _ in range(49):
                s1, t, rew, _ = e.env_step(s1, arr(r["action"]), t)
            np.testing.assert_allclose(s1, arr(r["state50"]), rtol=1e-6, atol=1e-7, err_msg=name)


def test_within_track_indices_bit_exact(fx, orc):
    w = fx["within_track"]
    env = make_env("car")
    for name, mk in backends(orc):
        e = configure(mk(**engine_kwargs("gmppi", env, 32, 5)), env, "gmppi")
        idx, _, dist, within = e.track_query(arr(w["pos"]).T)
        assert np.array_equal(idx, np.asarray(w["min_idx0"], dtype=idx.dtype)), name
        assert np.array_equal(within, np.asarray(w["within"], dtype=bool)), name
        np.testing.assert_allclose(dist, arr(w["dist"]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name,n_cars", [("car1", 1), ("car3", 3)])
def test_simulate_model_and_weights(fx, orc, name, n_cars):
    r = fx["simulate_model_" + name]
    env = make_env("car", n_cars)
    E = arr(r["E"])
    for bname, mk in backends(orc):
        e = configure(mk(**engine_kwargs("gmppi", env, E.shape[1], 50)), env, "gmppi")
        costs = e.rollout_costs(arr(r["state"]), 0, arr(r["U"]), arr(r["U"]), E)
        rel = np.abs(costs - arr(r["costs"])) / np.maximum(1.0, np.abs(arr(r["costs"])))
        assert (rel > 1e-9).sum() <= 1, f"{bname}: {(rel > 1e-9).sum()} costs differ from the reference"
        for lam, w in r["weights"].items():
            np.testing.assert_allclose(e.weights(arr(r["costs"]), float(lam)), arr(w), rtol=1e-12, atol=1e-300)
    assert np.array_equal(orc.sortperm(arr(r["costs"])), np.asarray(r["sortperm0"], dtype=np.int64))


@pytest.mark.parametrize("n", [30, 819])
@pytest.mark.parametrize("method", ["mle", "lw", "ss", "rblw", "oas"])
def test_covariance_estimators(fx, orc, n, method):
    r = fx[f"cov_n{n}"]
    X = arr(r["X"])
    e = orc.engine(policy="cemppi", env=_abi.ENV_MOUNTAIN_CAR, num_samples=n, horizon=X.shape[0], opt_its=2, lam=1.0)
    mu, S = e.cov_estimate(X, method)
    np.testing.assert_allclose(S, arr(r["cov_" + method]), rtol=1e-9, atol=1e-12,
                               err_msg=f"CovarianceEstimation {method} (POL:414-426,464)")


@pytest.mark.parametrize("n", [30, 819])
def test_weighted_and_unweighted_moments(fx, orc, n):
    r = fx[f"cov_n{n}"]
    X = arr(r["X"])
    e = orc.engine(policy="cemppi", env=_abi.ENV_MOUNTAIN_CAR, num_samples=n, horizon=X.shape[0], opt_its=2, lam=1.0)
    mu, S = e.cov_estimate(X, "mle", w=arr(r["w"]))
    np.testing.assert_allclose(mu, arr(r["mean_w"]), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(S, arr(r["cov_w"]), rtol=1e-9, atol=1e-12)
    mu, S = e.cov_estimate(X, "mle", corrected=True)
    np.testing.assert_allclose(mu, arr(r["mean_u"]), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(S, arr(r["cov_u"]), rtol=1e-9, atol=1e-12)


def test_inverse_square_root_and_cholesky(fx, orc):
    r = fx["inv_sqrt"]
    A = arr(r["A"])
    e = orc.engine(policy="cemppi", env=_abi.ENV_MOUNTAIN_CAR, num_samples=8, horizon=A.shape[0], opt_its=2, lam=1.0)
    np.testing.assert_allclose(e.inv_sqrt(A), arr(r["C"]), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(e.cholesky(A), arr(r["L"]), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("policy", ["gmppi", "imppi", "cemppi", "cmamppi", "μaismppi", "μΣaismppi"])
def test_full_control_steps_with_the_reference_noise(fx, orc, policy):
    rec = fx["policy_" + policy]
    K, T, N = rec["K"], rec["T"], rec["N"]
    env = make_env("car")
    for bname, mk in backends(orc):
        e = configure(mk(**engine_kwargs(policy, env, K, T, N, sigma_est="ss")), env, policy)
        for st in rec["steps"]:
            # Z_n = L_n⁻¹ E_n, L_n = chol(Σ′_n) with Σ′_n = inv(Σ⁻¹_n) as the reference held it
            Z = np.zeros((e.cs, K, e.N))
            for n, (E, Sinv) in enumerate(zip(st["E"], st["Sigma_inv"])):
                L = np.linalg.cholesky(np.linalg.inv(arr(Sinv)))
                Z[:, :, n] = np.linalg.solve(L, arr(E))
            ctrl, U2, its = e.plan(arr(st["state"]), 0, arr(st["U_before"]), Z=Z)
            assert its == st["its"], f"{bname} {policy}: AIS iterations {its} vs reference {st['its']}"
            np.testing.assert_allclose(ctrl, arr(st["control"]), rtol=1e-5, atol=1e-8)
            np.testing.assert_allclose(U2, arr(st["U_after"]), rtol=1e-5, atol=1e-8)
            rel = np.abs(e.fetch()["costs"] - arr(st["costs"][-1])) / np.maximum(1.0, np.abs(arr(st["costs"][-1])))
            assert (rel > 1e-7).sum() <= 1


def test_mountaincar_trajectory(fx, orc):
    r = fx["mountaincar"]
    env = make_env("mc")
    e = configure(orc.engine(**engine_kwargs("mppi", env, 8, 5)), env, "mppi")
    s, t = np.array([-0.5, 0.0]), 0
    for a, x, v, rew in zip(r["actions"], r["x"], r["v"], r["reward"]):
        s, t, rr, _ = e.env_step(s, [a], t)
        assert abs(s[0] - x) <= 1e-12 and abs(s[1] - v) <= 1e-12 and abs(rr - rew) <= 1e-9
