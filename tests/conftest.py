import os
import sys
from pathlib import Path

# loop-back groups run up to 8 virtual ranks x 2 streams on one device and their peer-memory collectives spin on
# flags: with the default 8 hardware queues a spinning kernel could sit in front of the very kernel it waits for
os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
# ... and they share ONE CUDA context: a lazily loaded kernel would wait for the spinning peer (csrc/comm.cu)
os.environ["CUDA_MODULE_LOADING"] = "EAGER"  # set, not setdefault: a LAZY inherited from the caller would fail the peer tests

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


# ---- shared builders ------------------------------------------------------------------------
from mpopis_b200 import _abi  # noqa: E402
from mpopis_b200.envs import CarRacingEnv, MountainCarEnv, MultiCarRacingEnv  # noqa: E402
from mpopis_b200.policies import block_diagm, cma_constants  # noqa: E402


def make_env(kind, n_cars=1):
    if kind == "mc":
        e = MountainCarEnv()
        e.reset(np.array([-0.5, 0.0]))  # fixed start (the reference's is unseeded-random)
        return e
    return CarRacingEnv() if n_cars == 1 else MultiCarRacingEnv(n_cars)


def configure(eng, env, policy, cov=None):
    """Wire an Engine (oracle- or CUDA-backed) like get_policy would."""
    env.configure_engine(eng)
    if cov is None:
        cov = [1.5] if isinstance(env, MountainCarEnv) else block_diagm([0.0625, 0.1], getattr(env, "N", 1))
    eng.set_sigma(np.asarray(cov, dtype=float))
    if policy == "cmamppi":
        c = cma_constants(eng.K, eng.cs, 0.8)
        eng.set_cma(sigma=0.75, m_elite=c["m_elite"], mu_eff=c["μ_eff"], c_sigma=c["cσ"], d_sigma=c["dσ"],
                    c_Sigma=c["cΣ"], c1=c["c1"], c_mu=c["cμ"], E_norm=c["E"], ws=c["ws"])
    return eng


def engine_kwargs(policy, env, K, T, N=10, lam=10.0, alpha=1.0, lam_ais=20.0, sigma_est="ss", **kw):
    kind = _abi.ENV_MOUNTAIN_CAR if isinstance(env, MountainCarEnv) else _abi.ENV_CAR_RACING
    return dict(policy=policy, env=kind, n_cars=getattr(env, "N", 1), num_samples=K, horizon=T, opt_its=N, lam=lam,
                alpha=alpha, lambda_ais=lam_ais, sigma_est=sigma_est, **kw)


def synthetic_states(n=15, seed=20260917):
    """on-track states with random speed / slip / steering."""
    env = CarRacingEnv()
    tx, ty = env.track.xs, env.track.ys
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for j in range(n):
        i = (3 * j) % len(tx)
        i2 = (i + 1) % len(tx)
        tang = np.arctan2(ty[i2] - ty[i], tx[i2] - tx[i])
        off = rng.uniform(-10, 10)
        x = tx[i] - off * np.sin(tang)
        y = ty[i] + off * np.cos(tang)
        out.append(np.array([x, y, tang, rng.uniform(5, 30), rng.normal(0, 1), rng.normal(0, 0.2),
                             rng.uniform(-0.31, 0.31), 0.0]))
    return out


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_bound():
    from mpopis_b200 import _lib
    return _lib.product()  # raises (fails the test) when the CUDA library is missing: no fallback


def julia_sortperm(x):
    """sortperm(x) with Base.isless: ascending, -0.0 before +0.0, ties by index."""
    x = np.asarray(x, dtype=np.float64)
    tie = np.where(x == 0, np.where(np.signbit(x), 0, 1), 0)
    return np.lexsort((np.arange(x.size), tie, x))
