"""Policy / operator API — host mirror of src/mppi_mpopi_policies.jl (POL) and
src/examples/example_utils.jl (get_policy).

Same constructor names, keyword arguments, defaults and error behaviour as the reference; the
functor `pol(env)` returns the control and rolls `pol.U`. All arithmetic of the functor
(sampling, rollouts, weights, AIS updates, control) happens in the CUDA engine behind the C-ABI
(`mpopis_b200_plan`); this file only derives sizes/constants and marshals.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from .engine import Engine


# --------------------------------------------------------------------------------------------
# helpers exported by the reference (UTL:2-21)
# --------------------------------------------------------------------------------------------
def action_space_size(act_space) -> int:
    lo, _ = act_space
    return len(lo)


def block_diagm(A, rep_number: int) -> np.ndarray:
    """block_diagm(A::Vector, n) = diagm(repeat(A, n)) (UTL:9-11); Matrix version UTL:13-21."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:
        return np.diag(np.tile(A, rep_number))
    r = A.shape[0]
    B = np.zeros((r * rep_number, r * rep_number))
    for ii in range(0, r * rep_number, r):
        B[ii:ii + r, ii:ii + r] = A
    return B


def julia_round(x: float) -> int:
    """round(Int, x): ties to even (POL:437, 515)."""
    return int(np.rint(x))


@dataclass
class Information_Theoretic:
    λ: float


@dataclass
class MPPI_Logger:
    """POL:2-6."""
    trajectories: list
    traj_costs: np.ndarray
    traj_weights: np.ndarray


@dataclass
class MPPI_Policy_Params:
    """POL:8-19."""
    num_samples: int
    horizon: int
    λ: float
    α: float
    U0: np.ndarray
    ss: int
    as_: int
    cs: int
    weight_method: Information_Theoretic
    log: bool


def make_policy_params(env, type_: str, *, num_samples=50, horizon=50, λ=1.0, α=1.0, U0=(0.0,), cov_mat=(1.0,),
                       weight_method="IT", elite_threshold=0.8, rng=None, log=False):
    """MPPI_Policy_Params(env, type; kwargs...) POL:36-102 -> (params, U0, Σ, rng, logger)."""
    st = np.asarray(env.state)
    if st.ndim == 1:
        ss = st.shape[0]
    elif st.ndim == 2:
        ss = st.shape[1]
    else:
        raise ValueError("State must be Vector or Matrix")  # POL:55
    as_ = action_space_size(env.action_space())
    cs = as_ * horizon
    U0 = np.asarray(U0, dtype=np.float64).reshape(-1)
    if U0.size == as_:
        U0 = np.tile(U0, horizon)  # POL:61-63
    if U0.size != cs:
        raise ValueError("U₀ must be length of action space or control space")  # POL:64
    if type_ == "mppi":
        repeat_num, check_size = 1, as_
    elif type_ == "gmppi":
        repeat_num, check_size = horizon, cs
    else:
        raise ValueError("Incorrect type for MPPPI")  # POL:73
    cov = np.asarray(cov_mat, dtype=np.float64)
    if cov.shape[0] == as_:
        cov = block_diagm(cov, repeat_num)  # POL:76-78
    if cov.ndim != 2 or cov.shape[0] != check_size:
        raise ValueError("Covariance matrix size problem")  # POL:79
    if cov.shape[0] != cov.shape[1]:
        raise ValueError("Covriance must be square")  # POL:80
    if weight_method == "IT":
        weight_m = Information_Theoretic(λ)
    else:
        # POL:85 references an undefined variable: any value but :IT throws in the reference
        raise ValueError(f"No cost method implemented for {weight_method}")
    logger = MPPI_Logger([np.empty((horizon, ss)) for _ in range(num_samples)] if log else [],
                         np.empty(num_samples), np.empty(num_samples))
    params = MPPI_Policy_Params(num_samples, horizon, float(λ), float(α), U0, ss, as_, cs, weight_m, bool(log))
    return params, U0, cov, rng, logger


def cma_constants(num_samples: int, cs: int, elite_perc_threshold: float):
    """CMAMPPI_Policy constructor arithmetic, POL:513-525."""
    m, n = num_samples, cs
    m_elite = julia_round((1.0 - elite_perc_threshold) * m)
    ws = math.log((m + 1) / 2) - np.log(np.arange(1, m + 1, dtype=np.float64))
    ws[:m_elite] /= np.sum(ws[:m_elite])
    μ_eff = 1 / np.sum(ws[:m_elite] ** 2)
    cσ = (μ_eff + 2) / (n + μ_eff + 5)
    dσ = 1 + 2 * max(0, math.sqrt((μ_eff - 1) / (n + 1)) - 1) + cσ
    cΣ = (4 + μ_eff / n) / (n + 4 + 2 * μ_eff / n)
    c1 = 2 / ((n + 1.3) ** 2 + μ_eff)
    cμ = min(1 - c1, 2 * (μ_eff - 2 + 1 / μ_eff) / ((n + 2) ** 2 + μ_eff))
    ws[m_elite:] *= -(1 + c1 / cμ) / np.sum(ws[m_elite:])
    E = n ** 0.5 * (1 - 1 / (4 * n) + 1 / (21 * n ** 2))
    return dict(m_elite=m_elite, ws=ws, μ_eff=float(μ_eff), cσ=float(cσ), dσ=float(dσ), cΣ=float(cΣ),
                c1=float(c1), cμ=float(cμ), E=float(E))


# --------------------------------------------------------------------------------------------
# policies
# --------------------------------------------------------------------------------------------
class AbstractPathIntegralPolicy:
    symbol = "gmppi"
    _family = "gmppi"

    def __init__(self, env, *, opt_its=1, backend=None, device=0, rank=0, world_size=1, nccl_id=None,
                 early_stop=True, **kwargs):
        self.params, self.U, self.Σ, self.rng, self.logger = make_policy_params(env, self._family, **kwargs)
        self.env = env
        self.opt_its = int(opt_its)
        self._engine_args = dict(device=device, rank=rank, world_size=world_size, early_stop=early_stop)
        self._nccl_id = nccl_id
        self._backend = backend
        self._eng = None
        self._seed = None

    # engine creation is deferred to the first use so that parameter errors surface exactly like
    # the reference's constructors (before any device is touched)
    def _extra_cfg(self) -> dict:
        return {}

    def _after_create(self, eng: Engine):
        pass

    def engine(self) -> Engine:
        if self._eng is None:
            if self._backend is None:
                from . import _lib
                bound = _lib.product()  # raises if the CUDA library is missing: no CPU fallback
            else:
                bound = self._backend
            p = self.params
            env_kind = self.env._env_kind()
            if env_kind == _abi.ENV_EXTERNAL:
                self._engine_args["ext_action_size"] = self.env.action_space_size()
            self._eng = Engine(bound, policy=self.symbol, env=env_kind, n_cars=getattr(self.env, "N", 1),
                               num_samples=p.num_samples, horizon=p.horizon, opt_its=self.opt_its, lam=p.λ,
                               alpha=p.α, log_trajectories=p.log, **self._engine_args, **self._extra_cfg())
            self.env.configure_engine(self._eng)
            self._eng.set_sigma(self.Σ)
            self._after_create(self._eng)
            if self._nccl_id is not None:
                self._eng.comm_init(self._nccl_id)
            if self._seed is not None:
                self._eng.seed(self._seed)
        return self._eng

    def connect(self, dist) -> str:
        """Join the shards of this policy (constructed with rank / world_size on every process) over an initialised
        torch.distributed group: NCCL communicator plus, on one NVLink node, the peer-memory collectives
        (sharding.connect). Returns the transport of the per-iteration exchanges: "peer" or "nccl"."""
        from . import sharding
        if self._nccl_id is not None:
            raise ValueError("this policy was given an nccl_id: its communicator is already set up")
        return sharding.connect(self.engine(), dist, self._engine_args["rank"], self._engine_args["world_size"])

    def seed(self, seed: int):
        """Random.seed!(pol, seed) (MPOPIS.jl:54) — keys the engine's Philox stream."""
        self._seed = int(seed)
        if self._eng is not None:
            self._eng.seed(self._seed)

    def __call__(self, env):
        """(pol::AbstractGMPPI_Policy)(env) POL:221-238 / (pol::MPPI_Policy)(env) POL:121-146."""
        eng = self.engine()
        if env._env_kind() == _abi.ENV_EXTERNAL:  # pol(env::EnvpoolEnv): POL:148-184, 240-259
            control, U_rolled, self.last_its = eng.plan_external(self.U, env.rollout)
        else:
            control, U_rolled, self.last_its = eng.plan(env.state, env.t, self.U)
        self.U[:] = U_rolled  # in place: pol.U aliases pol.params.U₀
        if self.params.log:  # POL:140-143, 233-236
            out = eng.fetch(costs=True, weights=True, traj=True)
            self.logger.traj_costs = out["costs"]
            self.logger.traj_weights = out["weights"]
            self.logger.trajectories = list(out["traj"])
        # get_model_controls returns a Vector for as == 1 and an as x 1 Matrix otherwise (UTL:63-66)
        return control if self.params.as_ == 1 else control.reshape(-1, 1)


class AbstractGMPPI_Policy(AbstractPathIntegralPolicy):
    pass


class MPPI_Policy(AbstractPathIntegralPolicy):
    """POL:107-119."""
    symbol, _family = "mppi", "mppi"


class GMPPI_Policy(AbstractGMPPI_Policy):
    """POL:284-301."""
    symbol = "gmppi"


class IMPPI_Policy(AbstractGMPPI_Policy):
    """POL:321-344."""
    symbol = "imppi"

    def __init__(self, env, *, opt_its=10, **kwargs):
        super().__init__(env, opt_its=opt_its, **kwargs)


class CEMPPI_Policy(AbstractGMPPI_Policy):
    """POL:379-432."""
    symbol = "cemppi"

    def __init__(self, env, *, opt_its=10, ce_elite_threshold=0.8, Σ_est="mle", **kwargs):
        if Σ_est not in _abi.SIGMA_EST:
            raise ValueError("CEMPPI_Policy - Not a valid Σ estimation method")  # POL:425
        super().__init__(env, opt_its=opt_its, **kwargs)
        self.ce_elite_threshold = float(ce_elite_threshold)
        self.Σ_estimation_method = Σ_est

    def _extra_cfg(self):
        return dict(ce_elite_threshold=self.ce_elite_threshold, sigma_est=self.Σ_estimation_method)


class CMAMPPI_Policy(AbstractGMPPI_Policy):
    """POL:478-530."""
    symbol = "cmamppi"

    def __init__(self, env, *, opt_its=10, σ=1.0, elite_perc_threshold=0.8, **kwargs):
        super().__init__(env, opt_its=opt_its, **kwargs)
        c = cma_constants(self.params.num_samples, self.params.cs, elite_perc_threshold)
        self.σ = float(σ)
        self.m_elite, self.ws = c["m_elite"], c["ws"]
        self.μ_eff, self.cσ, self.dσ, self.cΣ, self.c1, self.cμ, self.E = (
            c["μ_eff"], c["cσ"], c["dσ"], c["cΣ"], c["c1"], c["cμ"], c["E"])

    def _after_create(self, eng):
        eng.set_cma(sigma=self.σ, m_elite=self.m_elite, mu_eff=self.μ_eff, c_sigma=self.cσ, d_sigma=self.dσ,
                    c_Sigma=self.cΣ, c1=self.c1, c_mu=self.cμ, E_norm=self.E, ws=self.ws)


class _AIS(AbstractGMPPI_Policy):
    def __init__(self, env, *, opt_its=10, λ_ais=20.0, **kwargs):
        super().__init__(env, opt_its=opt_its, **kwargs)
        self.λ_ais = float(λ_ais)

    def _extra_cfg(self):
        return dict(lambda_ais=self.λ_ais)


class μAISMPPI_Policy(_AIS):
    """POL:612-637."""
    symbol = "μaismppi"


class μΣAISMPPI_Policy(_AIS):
    """POL:677-702."""
    symbol = "μΣaismppi"


class PMCMPPI_Policy(_AIS):
    """POL:748-773."""
    symbol = "pmcmppi"


def seed_b(pol, seed):
    """seed!(pol, seed)."""
    pol.seed(seed)


def get_policy(policy_type, env, num_samples, horizon, λ, α, U0, cov_mat, pol_log, ais_its, λ_ais,
               ce_elite_threshold, ce_Σ_est, cma_σ, cma_elite_threshold, **engine_kwargs):
    """get_policy (example_utils.jl:12-130)."""
    pt = str(policy_type).lstrip(":")
    common = dict(num_samples=num_samples, horizon=horizon, λ=λ, α=α, U0=U0, cov_mat=cov_mat, log=pol_log,
                  **engine_kwargs)
    if pt == "mppi":
        return MPPI_Policy(env, **common)
    if pt == "gmppi":
        return GMPPI_Policy(env, **common)
    if pt == "imppi":
        return IMPPI_Policy(env, opt_its=ais_its, **common)
    if pt == "cemppi":
        return CEMPPI_Policy(env, opt_its=ais_its, ce_elite_threshold=ce_elite_threshold, Σ_est=str(ce_Σ_est).lstrip(":"),
                             **common)
    if pt == "cmamppi":
        return CMAMPPI_Policy(env, opt_its=ais_its, σ=cma_σ, elite_perc_threshold=cma_elite_threshold, **common)
    if pt in ("μΣaismppi", "musigmaaismppi"):
        return μΣAISMPPI_Policy(env, opt_its=ais_its, λ_ais=λ_ais, **common)
    if pt in ("μaismppi", "muaismppi"):
        return μAISMPPI_Policy(env, opt_its=ais_its, λ_ais=λ_ais, **common)
    if pt == "pmcmppi":
        return PMCMPPI_Policy(env, opt_its=ais_its, λ_ais=λ_ais, **common)
    raise ValueError(f"No policy_type of {policy_type}")  # example_utils.jl:126
