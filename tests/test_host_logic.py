"""Host-side mirror of the reference's policy / entry-point API (no GPU): parameter derivation,
error behaviour, CMA constants, track sub-sampling, and the closed-loop drivers run end to end
with the ORACLE injected as backend (test-only seam; the product default is the CUDA library)."""
import io
import math

import numpy as np
import pytest
from conftest import make_env

import harness_examples as H
import mpopis_b200 as M
from mpopis_b200 import policies as P
from mpopis_b200.envs import CarRacingEnvParams
from mpopis_b200.tracks import Track, bundled_track_names


def test_block_diagm_and_round():
    B = M.block_diagm([0.0625, 0.1], 3)
    assert B.shape == (6, 6) and np.allclose(np.diag(B), [0.0625, 0.1] * 3)
    A = np.array([[1.0, 2.0], [3.0, 4.0]])
    B = M.block_diagm(A, 2)
    assert np.array_equal(B[2:, 2:], A) and np.all(B[:2, 2:] == 0)
    assert [P.julia_round(x) for x in (0.5, 1.5, 2.5, 29.999999999999996, 30.000000000000004)] == [0, 2, 2, 30, 30]


def test_policy_params_sizes_and_errors():
    env = make_env("car")
    params, U0, Σ, _, _ = P.make_policy_params(env, "gmppi", num_samples=150, horizon=50, U0=[0.0, 0.0],
                                               cov_mat=M.block_diagm([0.0625, 0.1], 1))
    assert (params.ss, params.as_, params.cs) == (8, 2, 100) and U0.shape == (100,) and Σ.shape == (100, 100)
    params, _, Σ, _, _ = P.make_policy_params(env, "mppi", horizon=50, U0=[0.0, 0.0], cov_mat=M.block_diagm([0.0625, 0.1], 1))
    assert Σ.shape == (2, 2)  # :mppi keeps the as x as covariance (POL:66-68)
    with pytest.raises(ValueError, match="U₀ must be length"):
        P.make_policy_params(env, "gmppi", U0=[0.0, 0.0, 0.0], cov_mat=[1.0, 1.0])
    with pytest.raises(ValueError, match="Covariance matrix size problem"):
        P.make_policy_params(env, "gmppi", U0=[0.0, 0.0], cov_mat=np.eye(3))
    with pytest.raises(ValueError, match="No cost method"):
        P.make_policy_params(env, "gmppi", U0=[0.0, 0.0], cov_mat=[1.0, 1.0], weight_method="CE")
    with pytest.raises(ValueError, match="Incorrect type"):
        P.make_policy_params(env, "xmppi", U0=[0.0, 0.0], cov_mat=[1.0, 1.0])
    env3 = make_env("car", 3)
    params, _, Σ, _, _ = P.make_policy_params(env3, "gmppi", horizon=50, U0=np.zeros(6), cov_mat=M.block_diagm([0.0625, 0.1], 3))
    assert (params.ss, params.as_, params.cs) == (24, 6, 300) and Σ.shape == (300, 300)
    mc = make_env("mc")
    params, _, Σ, _, _ = P.make_policy_params(mc, "gmppi", num_samples=20, horizon=15, U0=[0.0], cov_mat=[1.5])
    assert (params.ss, params.as_, params.cs) == (2, 1, 15) and np.allclose(np.diag(Σ), 1.5)


def test_cma_constants_match_survey_values():
    # SURVEY App. A-6 (evaluated from POL:513-525)
    c = M.cma_constants(150, 100, 0.8)
    assert c["m_elite"] == 30
    for k, v in dict(μ_eff=24.8445, cσ=0.206744, dσ=1.206744, cΣ=0.0406562, c1=1.94429e-4, cμ=4.38874e-3, E=9.97505).items():
        assert abs(c[k] - v) < 2e-6 * max(1, abs(v)) * 5, (k, c[k], v)
    c = M.cma_constants(375, 300, 0.8)
    assert c["m_elite"] == 75
    for k, v in dict(μ_eff=60.8005, cσ=0.171680, dσ=1.171680, cΣ=0.0138062, c1=2.20161e-5, cμ=1.28893e-3, E=17.3061).items():
        assert abs(c[k] - v) < 1e-5 * max(1, abs(v)), (k, c[k], v)
    assert M.cma_constants(20, 15, 0.8)["m_elite"] == 4
    ws = c["ws"]
    assert abs(ws[:75].sum() - 1) < 1e-12 and np.all(ws[:75] > 0) and ws[187] == 0.0 and np.all(ws[188:] < 0)


def test_tracks():
    assert "curve" in bundled_track_names() and len(bundled_track_names()) == 12
    t = Track("curve")
    assert (t.x.size, t.xs.size, t.sample_factor) == (946, 48, 20) and np.all(t.ws == 15.0)
    assert Track("curve", sample_factor=10).xs.size == 95
    with pytest.raises(ValueError):
        Track("curve", width=np.ones(3))
    env = M.CarRacingEnv()
    assert env.track.xs.size == 48 and np.allclose(env.state, [0, 0, math.pi / 2, 10, 0, 0, 0, 0])
    assert M.CarRacingEnv(CarRacingEnvParams()).track.xs.size == 95  # CarRacingEnv(params) uses sample_factor 10 (CAR:133)
    env3 = M.MultiCarRacingEnv(3)
    assert np.allclose(env3.state.reshape(3, 8)[:, 0], [0.0, 5.0, -5.0])  # MCR:166-174
    assert np.allclose(CarRacingEnvParams().as_array()[[11, 12, 17]], np.radians([18, 90, 45]))


def test_get_policy_symbols_and_defaults():
    env = make_env("car")
    args = (env, 150, 50, 10.0, 1.0, np.zeros(2), M.block_diagm([0.0625, 0.1], 1), False, 10, 20.0, 0.8, "ss", 0.75, 0.8)
    names = {"mppi": "MPPI_Policy", "gmppi": "GMPPI_Policy", "imppi": "IMPPI_Policy", "cemppi": "CEMPPI_Policy",
             "cmamppi": "CMAMPPI_Policy", "μΣaismppi": "μΣAISMPPI_Policy", "μaismppi": "μAISMPPI_Policy", "pmcmppi": "PMCMPPI_Policy"}
    for sym, cls in names.items():
        pol = M.get_policy(sym, *args)
        assert type(pol).__name__ == cls and pol.params.num_samples == 150
        assert M.get_policy(":" + sym, *args).symbol == pol.symbol
    with pytest.raises(ValueError, match="No policy_type"):
        M.get_policy("nesmppi", *args)
    with pytest.raises(ValueError, match="Not a valid Σ estimation"):
        M.CEMPPI_Policy(env, Σ_est="bogus", U0=[0.0, 0.0], cov_mat=[1.0, 1.0])
    pol = M.get_policy("cmamppi", *args)
    assert pol.m_elite == 30 and abs(pol.σ - 0.75) < 1e-15


def test_product_backend_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a B200 the product path must raise, never compute."""
    import ctypes
    try:
        n = ctypes.CDLL("libcuda.so.1").cuInit(0)
    except OSError:
        n = -1
    if n == 0:
        pytest.skip("a CUDA driver is present on this box")
    env = make_env("mc")
    pol = M.GMPPI_Policy(env, num_samples=20, horizon=15, U0=[0.0], cov_mat=[1.5], λ=0.1)
    with pytest.raises(Exception, match="no CUDA device|error -5|error -2"):
        pol(env)


def test_simulate_mountaincar_with_oracle_backend(orc):
    """BASELINE config 1 (MountainCar :mppi K=20 H=15, reference plumbing on CPU)."""
    from mpopis_b200.envs import _DeviceEnvMixin
    _DeviceEnvMixin._backend = orc.bound()
    try:
        buf = io.StringIO()
        res = H.simulate_mountaincar(num_trials=2, num_steps=60, policy_type="mppi", num_samples=20, horizon=15, λ=0.1,
                                     seed=11, x0=-0.5, out=buf, backend=orc.bound())
        txt = buf.getvalue()
        assert "MountainCar" in txt and "Trials AVE" in txt and "Trial    1" in txt
        assert np.all(res["trials"]["steps"] >= 1) and np.all(np.isfinite(res["trials"]["rews"]))
        res2 = H.simulate_mountaincar(num_trials=2, num_steps=60, policy_type="mppi", num_samples=20, horizon=15, λ=0.1,
                                      seed=11, x0=-0.5, out=io.StringIO(), backend=orc.bound())
        assert np.array_equal(res["trials"]["rews"], res2["trials"]["rews"])  # same seed -> same run
        res3 = H.simulate_mountaincar(num_trials=1, num_steps=200, policy_type="cemppi", seed=3, x0=-0.5,
                                      out=io.StringIO(), backend=orc.bound())
        assert res3["trials"]["rews"][0] > 50000  # the default CE-MPPI controller reaches the goal
    finally:
        _DeviceEnvMixin._backend = None


def test_simulate_car_racing_with_oracle_backend(orc):
    from mpopis_b200.envs import _DeviceEnvMixin
    _DeviceEnvMixin._backend = orc.bound()
    try:
        buf = io.StringIO()
        res = H.simulate_car_racing(num_trials=1, num_steps=12, num_samples=64, horizon=20, ais_its=3, seed=5, out=buf,
                                    backend=orc.bound())
        assert res["trials"]["steps"][0] == 12 and res["trials"]["T_viols"][0] == 0
        assert res["trials"]["mean_vs"][0] > 9.0 and "CE Σ Est Method:" in buf.getvalue()
        res = H.simulate_car_racing(num_trials=1, num_steps=5, num_cars=2, policy_type="μaismppi", num_samples=32, horizon=10,
                                    ais_its=2, seed=5, out=io.StringIO(), backend=orc.bound())
        assert res["trials"]["steps"][0] == 5
    finally:
        _DeviceEnvMixin._backend = None


def test_policy_logger_with_oracle_backend(orc):
    env = make_env("car")
    env._backend = orc.bound()
    pol = M.CEMPPI_Policy(env, num_samples=16, horizon=6, λ=10.0, U0=[0.0, 0.0], cov_mat=M.block_diagm([0.0625, 0.1], 1),
                          opt_its=2, log=True, backend=orc.bound())
    pol.seed(1)
    a = pol(env)
    assert a.shape == (2, 1) and len(pol.logger.trajectories) == 16 and pol.logger.trajectories[0].shape == (6, 8)
    assert pol.logger.traj_costs.shape == (16,) and abs(pol.logger.traj_weights.sum() - 1) < 1e-12
    # the logged trajectory of sample k ends where re-simulating its (clamped) controls ends
    mc = make_env("mc")
    pol = M.MPPI_Policy(mc, num_samples=8, horizon=5, λ=0.1, U0=[0.0], cov_mat=[1.5], backend=orc.bound())
    assert pol(mc).shape == (1,)  # Vector for as == 1 (UTL:63-66)


def test_trial_replicas_schedule_breadth_first_in_waves():
    """run_trial_replicas (car_example.jl:170 as concurrent replicas) — the host-side schedule, with a
    recording stand-in for the engine: trials are dealt round-robin over the devices, every trial gets its own seed,
    step s of every resident trial is enqueued before step s + 1 of any, `concurrency` bounds the resident trials per
    wave, nothing is read back before a wave's last step, and every handle is closed."""
    log = []

    class FakeEnv:
        state, t = np.arange(8.0), 3

    class FakeEngine:
        cs = 4

        def __init__(self, dev, idx):
            self.dev, self.idx, self.steps = dev, idx, 0

        def seed(self, s):
            log.append(("seed", self.idx, s))

        def resident_reset(self, state, t, U):
            assert t == 3 and U.shape == (4,) and not U.any()
            log.append(("reset", self.idx))

        def resident_plan(self, advance):
            assert advance is True
            self.steps += 1
            log.append(("plan", self.idx, self.steps))

        def resident_read(self):
            log.append(("read", self.idx))
            return np.full(8, self.idx), np.zeros(4), np.zeros(2), None

        def resident_reward_sum(self):
            return float(self.idx)

        def resident_total_its(self):
            return 10 * self.steps

        def close(self):
            log.append(("close", self.idx))

    made = []

    def make_engine(dev):
        made.append(dev)
        return FakeEnv(), FakeEngine(dev, len(made) - 1)

    out, wall = M.run_trial_replicas(make_engine, num_trials=5, num_steps=3, devices=(0, 1), seeds=[11, 12, 13, 14, 15],
                                     concurrency=2)
    assert made == [0, 1, 0, 1, 0] and wall >= 0.0
    assert [o["device"] for o in out] == made and [o["seed"] for o in out] == [11, 12, 13, 14, 15]
    assert [o["reward_sum"] for o in out] == [0.0, 1.0, 2.0, 3.0, 4.0] and all(o["its"] == 30 for o in out)
    waves = [[0, 1], [2, 3], [4]]
    pos = 0
    for wave in waves:
        expect = [x for i in wave for x in (("seed", i, 11 + i), ("reset", i))]
        expect += [("plan", i, s) for s in (1, 2, 3) for i in wave]            # breadth first
        expect += [("read", i) for i in wave] + [("close", i) for i in wave]   # read back the whole wave, then release it
        assert log[pos:pos + len(expect)] == expect
        pos += len(expect)
    assert pos == len(log)
    out, _ = M.run_trial_replicas(make_engine, num_trials=2, num_steps=1)      # defaults: seeds 1.., one device, one wave
    assert [o["seed"] for o in out] == [1, 2] and [o["device"] for o in out] == [0, 0]
