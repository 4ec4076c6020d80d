"""Host-side environment objects — mirror of the reference's env types as seen by the policies
and the entry points: CarRacingEnv (CAR), MultiCarRacingEnv (MCR) and RLEnvs' continuous
MountainCarEnv with the overrides of mountaincar_example.jl:4-22 (EXM).

These hold parameters, state and the step counter (what the Julia shim reads to fill the C
structs). The dynamics, `reward` and `within_track` are NOT re-implemented on the
host: `env(a)` and `reward(env)` call the engine's device kernels (mpopis_b200_env_step /
mpopis_b200_env_reward), the same code the rollouts use.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, fields

import numpy as np

from . import _abi
from .tracks import Track


@dataclass
class CarRacingEnvParams:
    """CAR:2-21 (declaration order == C-ABI order); defaults CAR:68-93."""
    m: float = 2000.0
    Izz: float = 3764.0
    h_cm: float = 0.3
    l_f: float = 1.53
    l_r: float = 1.23
    C_D0: float = 241.0
    C_D1: float = 25.1
    C_αf: float = 150000.0
    C_αr: float = 280000.0
    μ_f: float = 0.9
    μ_r: float = 0.9
    δ_max: float = math.radians(18)
    δ_dot_max: float = math.radians(90)
    Fx_max: float = 7200.0
    Fx_min: float = 22500.0
    λ_brake: float = 0.6
    λ_drive: float = 0.0
    β_limit: float = math.radians(45)

    def as_array(self) -> np.ndarray:
        return np.array([getattr(self, f.name) for f in fields(self)], dtype=np.float64)


class _DeviceEnvMixin:
    """Lazily creates a tiny engine handle used only for env(a) / reward(env)."""
    _eng = None
    _backend = None  # tests may inject the oracle's bound library here

    def _engine(self):
        if self._eng is None:
            from . import _lib
            from .engine import Engine
            bound = self._backend if self._backend is not None else _lib.product()
            self._eng = Engine(bound, policy="gmppi", env=self._env_kind(), n_cars=getattr(self, "N", 1),
                               num_samples=32, horizon=1)
            self._configure(self._eng)
        return self._eng


class CarRacingEnv(_DeviceEnvMixin):
    """mutable struct CarRacingEnv (CAR:28-37). state = [x, y, Ψ, Vx, Vy, Ψ̇, δ, pedal] (CAR:161-172)."""
    N = 1

    def __init__(self, params: CarRacingEnvParams | None = None, *, dt=0.1, δt=0.01, track="curve",
                 track_sample_factor=None, rng=None, **param_kwargs):
        # CarRacingEnv(; kwargs...) uses sample_factor 20 (CAR:91); CarRacingEnv(params; ...) uses 10 (CAR:133)
        if track_sample_factor is None:
            track_sample_factor = 20 if params is None else 10
        self.params = params if params is not None else CarRacingEnvParams(**param_kwargs)
        self.state = np.zeros(8)
        self.done = False
        self.t = 0
        self.dt, self.δt = float(dt), float(δt)
        self.track = track if isinstance(track, Track) else Track(track, sample_factor=track_sample_factor)
        self.rng = rng if rng is not None else np.random.default_rng()
        self.last_reward = 0.0
        self.reset()

    def _env_kind(self):
        return _abi.ENV_CAR_RACING

    def _configure(self, eng):
        eng.set_car_env(self.params.as_array(), self.dt, self.δt, self.track.xs, self.track.ys, self.track.ws)

    def configure_engine(self, eng):
        self._configure(eng)

    def action_space(self):
        return np.array([-1.0, -1.0]), np.array([1.0, 1.0])  # CAR:156-159

    def action_space_size(self) -> int:
        return 2

    def reset(self, state=None):
        """reset!(env) CAR:215-223 / reset!(env, state) CAR:225-230."""
        if state is None:
            self.state = np.zeros(8)
            self.state[2] = math.radians(90)
            self.state[3] = 10.0
        else:
            self.state = np.asarray(state, dtype=np.float64).copy()
        self.t = 0
        self.done = False

    def seed(self, seed):
        self.rng = np.random.default_rng(seed)  # Random.seed!(env.rng, seed), CAR:154

    def __call__(self, a):
        """env(a) CAR:238-250."""
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2:
            if a.shape[1] != 1:
                raise ValueError("Only implented for one step")  # CAR:248
            a = a[:, 0]
        lo, hi = self.action_space()
        if a.shape != (2,) or np.any(a < lo) or np.any(a > hi) or np.any(np.isnan(a)):
            raise ValueError("Action is not in action space")  # CAR:239
        self.state, self.t, self.last_reward, _ = self._engine().env_step(self.state, a, self.t)
        return self


def reward(env) -> float:
    """RLBase.reward(env) — CAR:201-213, MCR:145-158, EXM:10-22 (evaluated on the device)."""
    return env._engine().env_reward(env.state, getattr(env, "done", False))


def within_track(env):
    """within_track(env) -> (within, dist) for CarRacingEnv (CAR:178-180); bool for MultiCarRacingEnv (MCR:122-128)."""
    eng = env._engine()
    if isinstance(env, MultiCarRacingEnv):
        pos = env.state.reshape(env.N, 8)[:, :2]
        return bool(np.all(eng.track_query(pos)[3]))
    _, _, dist, within = eng.track_query(env.state[:2].reshape(1, 2))
    return bool(within[0]), float(dist[0])


def calculate_β(env) -> float:
    return math.atan2(env.state[4], env.state[3])  # CAR:181-183


def exceed_β(env) -> bool:
    if isinstance(env, MultiCarRacingEnv):  # MCR:130-136
        return any(abs(math.atan2(s[4], s[3])) > p.β_limit
                   for s, p in zip(env.state.reshape(env.N, 8), env.car_params))
    return abs(calculate_β(env)) > env.params.β_limit  # CAR:184-189


class MultiCarRacingEnv(_DeviceEnvMixin):
    """mutable struct MultiCarRacingEnv (MCR:2-12): N cars, joint state 8N, joint action 2N."""

    def __init__(self, N=2, *, dt=0.1, δt=0.01, track="curve", car_params=(), rng=None):
        if len(car_params) > N:
            raise ValueError("# Car parameters must be ≤ # cars")  # MCR:35
        if N > _abi.MAX_CARS:
            raise ValueError(f"at most {_abi.MAX_CARS} cars are supported by the engine")
        self.N = int(N)
        self.car_params = [car_params[i] if i < len(car_params) else CarRacingEnvParams() for i in range(self.N)]
        # sub-envs load their Track with sample_factor 20 (MCR:42 -> CAR:91), 10 if params were passed (MCR:40)
        sf = 10 if len(car_params) > 0 else 20
        self.track = track if isinstance(track, Track) else Track(track, sample_factor=sf)
        self.dt, self.δt = float(dt), float(δt)
        self.state = np.zeros(8 * self.N)
        self.done = False
        self.t = 0
        self.rng = rng if rng is not None else np.random.default_rng()
        self.last_reward = 0.0
        self.reset()

    def _env_kind(self):
        return _abi.ENV_CAR_RACING

    def _configure(self, eng):
        P = np.concatenate([p.as_array() for p in self.car_params])
        eng.set_car_env(P, self.dt, self.δt, self.track.xs, self.track.ys, self.track.ws)

    def configure_engine(self, eng):
        self._configure(eng)

    def action_space(self):
        return -np.ones(2 * self.N), np.ones(2 * self.N)  # MCR:75-84

    def action_space_size(self) -> int:
        return 2 * self.N

    def reset(self, state=None):
        """reset!(env) MCR:160-180: car 1 at the origin, others offset ±5 m in x per pair."""
        if state is None:
            s = np.zeros((self.N, 8))
            for ii in range(1, self.N + 1):
                if ii >= 2:
                    s[ii - 1, 0] = ii / 2 * 5.0 if ii % 2 == 0 else (1 - ii) / 2 * 5.0
                s[ii - 1, 2] = math.radians(90)
                s[ii - 1, 3] = 10.0
            self.state = s.reshape(-1)
        else:
            self.state = np.asarray(state, dtype=np.float64).copy()
        self.t = 0
        self.done = False

    def seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, a):
        """env(a) MCR:200-216 (no action-space check, as in the reference)."""
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2:
            if a.shape[1] != 1:
                raise ValueError("Only implented for one step")
            a = a[:, 0]
        if a.size != 2 * self.N:
            raise ValueError("Action space of each car is of size 2")  # MCR:201
        self.state, _, self.last_reward, _ = self._engine().env_step(self.state, a, self.t)
        return self


@dataclass
class MountainCarEnvParams:
    """RLEnvs MountainCarEnvParams for continuous=true."""
    min_pos: float = -1.2
    max_pos: float = 0.6
    max_speed: float = 0.07
    goal_pos: float = 0.45
    goal_velocity: float = 0.0
    power: float = 0.0015
    gravity: float = 0.0025
    max_steps: int = 200

    def as_array(self) -> np.ndarray:
        return np.array([self.min_pos, self.max_pos, self.max_speed, self.goal_pos, self.goal_velocity,
                         self.power, self.gravity], dtype=np.float64)


class MountainCarEnv(_DeviceEnvMixin):
    """RLEnvs MountainCarEnv(continuous=true) as driven by EXM:4-22, 126."""

    def __init__(self, *, continuous=True, rng=None, **param_kwargs):
        if not continuous:
            raise ValueError("only MountainCarEnv(continuous=true) is on the MPOPIS path (EXM:126)")
        self.params = MountainCarEnvParams(**param_kwargs)
        self.rng = rng if rng is not None else np.random.default_rng()
        self.state = np.zeros(2)
        self.done = False
        self.t = 0
        self.last_reward = 0.0
        self.reset()

    def _env_kind(self):
        return _abi.ENV_MOUNTAIN_CAR

    def _configure(self, eng):
        eng.set_mountaincar_env(self.params.as_array(), self.params.max_steps)

    def configure_engine(self, eng):
        self._configure(eng)

    def action_space(self):
        return np.array([-1.0]), np.array([1.0])

    def action_space_size(self) -> int:
        return 1

    def reset(self, state=None):
        if state is None:  # reset!: x = 0.2·rand() − 0.6, v = 0
            self.state = np.array([0.2 * self.rng.random() - 0.6, 0.0])
        else:
            self.state = np.asarray(state, dtype=np.float64).copy()
        self.t = 0
        self.done = False

    def seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, a):
        a = np.atleast_1d(np.asarray(a, dtype=np.float64)).reshape(-1)
        if a.size != 1:
            raise ValueError("Only implented for 1 step")  # EXM:5
        self.state, self.t, self.last_reward, self.done = self._engine().env_step(self.state, a, self.t)
        return self


class ExternalEnv:
    """The EnvpoolEnv seam (src/envs/envpool_env.jl; POL:148-184, 240-259; UTL:42-53, 103-121): a batched
    simulator that lives with the caller. `rollout(controls[K, as, T]) -> costs[K]` must step the simulator's K
    model environments from the current real state with the given (already clamped) controls, return
    −Σ_t reward per sample and restore the simulator (reset!(env; restore=true)). Sampling, control cost,
    adaptation, weights and the control update run on the device."""

    def __init__(self, action_low, action_high, rollout):
        self.lo = np.atleast_1d(np.asarray(action_low, dtype=np.float64))
        self.hi = np.atleast_1d(np.asarray(action_high, dtype=np.float64))
        if self.lo.shape != self.hi.shape or self.lo.ndim != 1:
            raise ValueError("action bounds must be vectors of the same length")
        self.rollout = rollout
        self.state = np.zeros(0)
        self.t = 0
        self.done = False

    def _env_kind(self):
        return _abi.ENV_EXTERNAL

    def configure_engine(self, eng):
        eng.set_external_env(self.lo, self.hi)

    def action_space(self):
        return self.lo, self.hi

    def action_space_size(self) -> int:
        return int(self.lo.size)


def state(env) -> np.ndarray:
    return env.state
