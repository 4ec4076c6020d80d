"""The EnvpoolEnv seam (MPOPIS_ENV_EXTERNAL; POL:148-184, 240-259; UTL:42-53, 103-121): the engine samples,
adds the control cost, adapts, weighs and updates the control, while a caller-side batched simulator rolls the K
control sequences out. Here the "external simulator" is the oracle's own MountainCar / CarRacing rollout, so the
external plan must reproduce the built-in plan of the same policy on the same injected noise.

CPU part: oracle(external) == oracle(built-in) for every policy symbol.  GPU part: CUDA engine(external) == oracle
(built-in), through the C-ABI callback."""
import numpy as np
import pytest

from conftest import configure, engine_kwargs, make_env
from mpopis_b200 import _abi
from mpopis_b200.engine import Engine, EngineError
from mpopis_b200.policies import block_diagm

POLICIES = ["mppi", "gmppi", "imppi", "cemppi", "cmamppi", "μaismppi", "μΣaismppi", "pmcmppi"]


def simulator(orc, envname, T, K, alpha_free_kw):
    """A batched simulator living with the caller: rolls K clamped control sequences out with the oracle's env."""
    env = make_env(envname)
    sim = configure(orc.engine(nthreads=4, **engine_kwargs("gmppi", env, K, T, 1, **alpha_free_kw)), env, "gmppi")
    calls = []

    def rollout(controls):  # [K, as, T] -> costs[K]
        K_, as_, T_ = controls.shape
        E = controls.transpose(2, 1, 0).reshape(T_ * as_, K_)  # r = a + as*t
        z = np.zeros(T_ * as_)
        calls.append(controls.copy())
        return sim.rollout_costs(env.state, env.t, z, z, E)

    return env, rollout, calls


def external_engine(bound_or_orc, policy, as_, K, T, N, lo, hi, is_oracle, **kw):
    args = dict(policy=policy, env=_abi.ENV_EXTERNAL, num_samples=K, horizon=T, opt_its=N, ext_action_size=as_, **kw)
    e = bound_or_orc.engine(nthreads=1, **args) if is_oracle else Engine(bound_or_orc, **args)
    e.set_external_env(lo, hi)
    return e


def run_pair(ext, ref, env, rollout, policy, cov, seed=3):
    from mpopis_b200.policies import cma_constants
    for e in (ext, ref):
        e.set_sigma(np.asarray(cov, dtype=float))
        if policy == "cmamppi":
            c = cma_constants(e.K, e.cs, 0.8)
            e.set_cma(sigma=0.75, m_elite=c["m_elite"], mu_eff=c["μ_eff"], c_sigma=c["cσ"], d_sigma=c["dσ"],
                      c_Sigma=c["cΣ"], c1=c["c1"], c_mu=c["cμ"], E_norm=c["E"], ws=c["ws"])
    rng = np.random.Generator(np.random.Philox(key=seed))
    Z = rng.standard_normal((ext.cs, ext.K, ext.N))
    u = rng.random((ext.K, max(ext.N - 1, 1)))
    U0 = rng.uniform(-0.3, 0.3, ext.cs)
    c1, U1, its1 = ext.plan_external(U0, rollout, Z=Z, resample_u=u)
    c2, U2, its2 = ref.plan(env.state, env.t, U0, Z=Z, resample_u=u)
    return (c1, U1, its1, ext.fetch()), (c2, U2, its2, ref.fetch())


@pytest.mark.parametrize("policy", POLICIES)
def test_oracle_external_equals_builtin_mountaincar(orc, policy):
    K, T, N = 24, 15, 4
    kw = dict(lam=0.1, alpha=0.7, lambda_ais=0.1, sigma_est="mle")
    env, rollout, calls = simulator(orc, "mc", T, K, dict(lam=0.1))
    ext = external_engine(orc, policy, 1, K, T, N, [-1.0], [1.0], True, **kw)
    ref = configure(orc.engine(nthreads=1, **engine_kwargs(policy, env, K, T, N, lam=0.1, alpha=0.7, lam_ais=0.1,
                                                           sigma_est="mle")), env, policy, cov=[1.5])
    (c1, U1, i1, f1), (c2, U2, i2, f2) = run_pair(ext, ref, env, rollout, policy, [1.5])
    assert i1 == i2 and len(calls) == i1
    assert all(np.all(np.abs(c) <= 1.0) for c in calls)  # clamped to the action space (UTL:42-53)
    np.testing.assert_allclose(f1["costs"], f2["costs"], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(c1, c2, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(U1, U2, rtol=1e-10, atol=1e-12)


def test_oracle_external_callback_error_is_reported(orc):
    ext = external_engine(orc, "gmppi", 1, 8, 5, 1, [-1.0], [1.0], True, lam=1.0)
    ext.set_sigma(np.array([1.0]))
    with pytest.raises(RuntimeError, match="simulator down"):
        ext.plan_external(np.zeros(5), lambda c: (_ for _ in ()).throw(RuntimeError("simulator down")))


@pytest.mark.gpu
@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("envname", ["mc", "car"])
def test_gpu_external_equals_oracle_builtin(gpu_bound, orc, policy, envname):
    if envname == "mc":
        K, T, N, as_, cov = 160, 15, 4, 1, [1.5]  # m_elite = 32 > cs = 15: Σ′ stays well conditioned
        kw = dict(lam=0.1, alpha=0.7, lam_ais=0.1, sigma_est="mle")
    else:
        K, T, N, as_, cov = 192, 12, 4, 2, block_diagm([0.0625, 0.1], 1)
        kw = dict(lam=10.0, alpha=0.9, lam_ais=20.0, sigma_est="ss")
    env, rollout, calls = simulator(orc, envname, T, K, dict(lam=kw["lam"]))
    ekw = dict(lam=kw["lam"], alpha=kw["alpha"], lambda_ais=kw["lam_ais"], sigma_est=kw["sigma_est"])
    ext = external_engine(gpu_bound, policy, as_, K, T, N, [-1.0] * as_, [1.0] * as_, False, **ekw)
    ref = configure(orc.engine(nthreads=4, **engine_kwargs(policy, env, K, T, N, **kw)), env, policy, cov=cov)
    (c1, U1, i1, f1), (c2, U2, i2, f2) = run_pair(ext, ref, env, rollout, policy, cov)
    assert i1 == i2 and len(calls) == i1  # one callback per executed AIS iteration, early stop included
    np.testing.assert_allclose(f1["costs"], f2["costs"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(c1, c2, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(U1, U2, rtol=1e-5, atol=1e-7)
    assert ext.launch_count() > 0


@pytest.mark.gpu
def test_gpu_external_misuse_errors(gpu_bound):
    ext = external_engine(gpu_bound, "cemppi", 2, 32, 5, 3, [-1, -1], [1, 1], False, lam=1.0)
    with pytest.raises(EngineError, match="plan_external"):
        ext.plan(np.zeros(0), 0, np.zeros(10))  # the built-in entry point has no simulator to call
    with pytest.raises(EngineError, match="world_size|ext_action_size"):
        Engine(gpu_bound, policy="cemppi", env=_abi.ENV_EXTERNAL, num_samples=32, horizon=5, ext_action_size=0)


def test_policy_functor_with_an_external_env(orc):
    """pol(env::EnvpoolEnv) through the Python mirror: get_policy + ExternalEnv drive plan_external, and a closed loop
    on a caller-side simulator (a point mass with quadratic cost, in numpy) makes progress. Oracle backend, no GPU."""
    from mpopis_b200 import ExternalEnv, get_policy

    class PointMass:  # the caller's batched simulator: x'' = a, cost = Σ_t (x − 1)² + 0.1 v²
        def __init__(self):
            self.x, self.v, self.dt = 0.0, 0.0, 0.1

        def rollout(self, controls):  # [K, as=1, T] -> costs[K]; the simulator itself is not advanced (restore=true)
            K, _, T = controls.shape
            x, v, cost = np.full(K, self.x), np.full(K, self.v), np.zeros(K)
            for t in range(T):
                v = v + self.dt * controls[:, 0, t]
                x = x + self.dt * v
                cost += (x - 1.0) ** 2 + 0.1 * v ** 2
            return cost

        def step(self, a):
            self.v += self.dt * float(a)
            self.x += self.dt * self.v

    sim = PointMass()
    env = ExternalEnv([-2.0], [2.0], sim.rollout)
    pol = get_policy("cemppi", env, 64, 12, 1.0, 1.0, [0.0], [1.0], False, 4, 20.0, 0.8, "mle", 0.75, 0.8,
                     backend=orc.bound())
    pol.seed(5)
    assert pol.params.as_ == 1 and pol.params.ss == 0 and pol.params.cs == 12
    for _ in range(40):
        a = pol(env)
        assert a.shape == (1,) and -2.0 <= a[0] <= 2.0  # clamped first action (UTL:88-101)
        sim.step(a[0])
    assert abs(sim.x - 1.0) < 0.25  # the controller drove the point mass to the target
