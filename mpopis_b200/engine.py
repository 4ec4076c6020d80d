"""Thin object wrapper over one C-ABI handle (include/mpopis_b200.h).

`Engine` is backend-agnostic on purpose: given the product library it drives the CUDA engine;
tests hand it the CPU oracle's bound library (same signatures, prefix `orc_`) to get the checker.
The product code path never constructs an Engine on anything but libmpopis_b200.so
(see `_lib.product()`), and that loader fails loudly when the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mpopis_b200 error {code}: {msg}")
        self.code = code


def _d(a):
    """double* of a C-contiguous float64 array (or NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _colmajor(a) -> np.ndarray:
    """Flat buffer of a 2-D/3-D array in Julia (column-major) order."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


class Engine:
    """One policy handle. Sizes follow MPPI_Policy_Params (POL:8-19): K, T, as, cs, ss."""

    def __init__(self, bound: _abi.Bound, *, policy: str, env: int, n_cars: int = 1, num_samples: int,
                 horizon: int, opt_its: int = 1, lam: float = 1.0, alpha: float = 1.0, lambda_ais: float = 20.0,
                 ce_elite_threshold: float = 0.8, sigma_est: str = "mle", early_stop: bool = True,
                 log_trajectories: bool = False, device: int = 0, rank: int = 0, world_size: int = 1,
                 ext_action_size: int = 0):
        if policy not in _abi.POLICY:
            raise ValueError(f"No policy_type of {policy}")  # example_utils.jl:126
        if sigma_est not in _abi.SIGMA_EST:
            raise ValueError("CEMPPI_Policy - Not a valid Σ estimation method")  # POL:425
        self.b = bound
        self.policy = policy
        self.cfg = _abi.Cfg(
            abi_version=_abi.ABI_VERSION, policy=_abi.POLICY[policy], env=env, n_cars=n_cars,
            num_samples=num_samples, horizon=horizon, opt_its=opt_its, lambda_=lam, alpha=alpha,
            lambda_ais=lambda_ais, ce_elite_threshold=ce_elite_threshold, sigma_est=_abi.SIGMA_EST[sigma_est],
            early_stop=int(early_stop), log_trajectories=int(log_trajectories), device=device, rank=rank,
            world_size=world_size, ext_action_size=int(ext_action_size))
        self.K, self.T = num_samples, horizon
        self.N = 1 if policy in ("mppi", "gmppi") else opt_its
        if env == _abi.ENV_EXTERNAL:
            self.as_, self.ss = int(ext_action_size), 0
        else:
            self.as_ = 2 * n_cars if env == _abi.ENV_CAR_RACING else 1
            self.ss = 8 * n_cars if env == _abi.ENV_CAR_RACING else 2
        self.cs = self.as_ * horizon
        self.world_size, self.rank, self.device = world_size, rank, device
        self.Kloc = num_samples // max(world_size, 1)
        self.h = C.c_void_p()
        self._chk(bound.create(C.byref(self.cfg), C.byref(self.h)))

    # -- plumbing ---------------------------------------------------------------------------
    def _chk(self, rc: int):
        if rc != 0:
            raise EngineError(rc, self.b.error())

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.b.destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ------------------------------------------------------------------------
    def set_car_env(self, params, dt, ddt, trk_x, trk_y, trk_w):
        p = _f64(params).reshape(-1)
        x, y, w = _f64(trk_x), _f64(trk_y), _f64(trk_w)
        n_cars = p.size // _abi.CAR_NPARAMS
        self._chk(self.b.set_car_env(self.h, n_cars, _d(p), dt, ddt, _d(x), _d(y), _d(w), x.size))

    def set_mountaincar_env(self, params7, max_steps):
        p = _f64(params7)
        self._chk(self.b.set_mountaincar_env(self.h, _d(p), int(max_steps)))

    def set_external_env(self, action_lo, action_hi):
        lo, hi = _f64(action_lo).reshape(-1), _f64(action_hi).reshape(-1)
        if lo.size != self.as_ or hi.size != self.as_:
            raise ValueError("action bounds must have one entry per action component")
        self._chk(self.b.set_external_env(self.h, _d(lo), _d(hi)))

    def plan_external(self, U, rollout, Z=None, resample_u=None):
        """pol(env::EnvpoolEnv): `rollout(controls)` gets the clamped model controls as a [K, as, T] array
        (get_model_controls UTL:42-53) and returns the K trajectory costs −Σ_t reward (UTL:103-121)."""
        Uio = _f64(U).copy()
        ctrl = np.zeros(self.as_)
        its = C.c_int32(0)
        err = []

        def trampoline(_user, controls, K, as_, T, out):
            try:
                ctl = np.ctypeslib.as_array(controls, shape=(T, as_, K)).transpose(2, 1, 0)  # [k, a, t] view
                costs = np.asarray(rollout(ctl), dtype=np.float64).reshape(-1)
                if costs.size != K:
                    raise ValueError(f"rollout returned {costs.size} costs for {K} samples")
                np.ctypeslib.as_array(out, shape=(K,))[:] = costs
                return 0
            except Exception as e:  # never unwind through the C frames
                err.append(e)
                return 1

        cb = _abi.ROLLOUT_FN(trampoline)
        Zc = None if Z is None else _colmajor(Z)
        uc = None if resample_u is None else _colmajor(resample_u)
        rc = self.b.plan_external(self.h, _d(Uio), cb, None, _d(Zc), _d(uc), _d(ctrl), C.byref(its))
        if err:
            raise err[0]
        self._chk(rc)
        return ctrl, Uio, its.value

    def set_sigma(self, Sigma):
        S = np.asarray(Sigma, dtype=np.float64)
        if S.ndim == 1:  # block_diagm(::Vector) = diagm, UTL:9-11
            S = np.diag(S)
        flat = _colmajor(S)
        self._chk(self.b.set_sigma(self.h, _d(flat), S.shape[0]))

    def set_cma(self, *, sigma, m_elite, mu_eff, c_sigma, d_sigma, c_Sigma, c1, c_mu, E_norm, ws):
        cma = _abi.Cma(sigma, m_elite, mu_eff, c_sigma, d_sigma, c_Sigma, c1, c_mu, E_norm)
        w = _f64(ws)
        self._chk(self.b.set_cma(self.h, C.byref(cma), _d(w), w.size))

    def seed(self, seed: int):
        self._chk(self.b.seed(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_option(self, key: str, value: float):
        self._chk(self.b.set_option(self.h, key.encode(), float(value)))

    def get_option(self, key: str) -> float:
        v = C.c_double(0.0)
        self._chk(self.b.get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    # -- the hot path ---------------------------------------------------------------------------
    def plan(self, state, env_t, U, Z=None, resample_u=None):
        """(pol)(env): returns (control[as], U_rolled[cs], its_run). Z: (N, cs, K)-indexable noise
        given as array of shape (cs, K, N) in Julia order, or None for the engine's Philox stream."""
        s, Uio = _f64(state), _f64(U).copy()
        ctrl = np.zeros(self.as_)
        its = C.c_int32(0)
        if Z is None:
            self._chk(self.b.plan(self.h, _d(s), int(env_t), _d(Uio), _d(ctrl), C.byref(its)))
        else:
            Zf = _colmajor(Z)
            assert Zf.size == self.cs * self.K * self.N, "Z must be cs x K x N"
            uf = None if resample_u is None else _colmajor(resample_u)
            self._chk(self.b.plan_with_noise(self.h, _d(s), int(env_t), _d(Uio), _d(Zf), _d(uf), _d(ctrl),
                                             C.byref(its)))
        return ctrl, Uio, its.value

    def fetch(self, costs=True, weights=True, E=False, traj=False):
        out = {}
        c = np.zeros(self.K) if costs else None
        w = np.zeros(self.K) if weights else None
        e = np.zeros(self.cs * self.Kloc) if E else None
        t = np.zeros(self.Kloc * self.T * self.ss) if traj else None
        self._chk(self.b.fetch(self.h, _d(c), _d(w), _d(e), _d(t)))
        if costs:
            out["costs"] = c
        if weights:
            out["weights"] = w
        if E:
            out["E"] = e.reshape((self.cs, self.Kloc), order="F")
        if traj:  # K matrices T x ss (column-major) -> [K, T, ss]
            out["traj"] = t.reshape((self.Kloc, self.ss, self.T)).transpose(0, 2, 1)
        return out

    def fetch_proposal(self):
        S, U = np.zeros(self.cs * self.cs), np.zeros(self.cs)
        self._chk(self.b.fetch_proposal(self.h, _d(S), _d(U)))
        return S.reshape((self.cs, self.cs), order="F"), U

    def rollout_costs(self, state, env_t, U, U_orig, E, Sigma_inv=None):
        """simulate_model(pol, env, E, Σ_inv, U_orig) POL:261-278; E is cs x K."""
        s, u, uo, e = _f64(state), _f64(U), _f64(U_orig), _colmajor(E)
        si = None if Sigma_inv is None else _colmajor(Sigma_inv)
        out = np.zeros(self.K)
        self._chk(self.b.rollout_costs(self.h, _d(s), int(env_t), _d(u), _d(uo), _d(e), _d(si), _d(out)))
        return out

    def weights(self, costs, lam):
        c = _f64(costs)
        w = np.zeros(c.size)
        self._chk(self.b.weights(self.h, _d(c), c.size, float(lam), _d(w)))
        return w

    def track_query(self, pos):
        """pos: (n, 2). Returns (idx, idx2, dist, within) — within_track, TRK:68-92 (0-based)."""
        p = _f64(pos).reshape(-1, 2)
        n = p.shape[0]
        idx, idx2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
        dist, within = np.zeros(n), np.zeros(n, np.uint8)
        self._chk(self.b.track_query(self.h, _d(p.reshape(-1)), n, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                     idx2.ctypes.data_as(C.POINTER(C.c_int32)), _d(dist),
                                     within.ctypes.data_as(C.POINTER(C.c_uint8))))
        return idx, idx2, dist, within.astype(bool)

    def env_step(self, state, action, env_t):
        s, a = _f64(state).copy(), _f64(action).reshape(-1)
        t, rew, done = C.c_int64(int(env_t)), C.c_double(0.0), C.c_uint8(0)
        self._chk(self.b.env_step(self.h, _d(s), _d(a), C.byref(t), C.byref(rew), C.byref(done)))
        return s, t.value, rew.value, bool(done.value)

    def env_reward(self, state, done=False):
        s, rew = _f64(state), C.c_double(0.0)
        self._chk(self.b.env_reward(self.h, _d(s), int(bool(done)), C.byref(rew)))
        return rew.value

    def sample_normals(self, step, iteration):
        z = np.zeros(self.cs * self.Kloc)
        self._chk(self.b.sample_normals(self.h, int(step), int(iteration), _d(z)))
        return z.reshape((self.cs, self.Kloc), order="F")

    def cov_estimate(self, X, sigma_est="mle", w=None, corrected=False):
        """X: p x n (columns = observations). Returns (mean[p], cov[p,p])."""
        Xa = np.asarray(X, dtype=np.float64)
        p, n = Xa.shape
        xf = _colmajor(Xa)
        wf = None if w is None else _f64(w)
        mu, S = np.zeros(p), np.zeros(p * p)
        self._chk(self.b.cov_estimate(self.h, _abi.SIGMA_EST[sigma_est], _d(xf), p, n, _d(wf), int(corrected),
                                      _d(mu), _d(S)))
        return mu, S.reshape((p, p), order="F")

    def last_shrinkage(self) -> float:
        v = C.c_double(0.0)
        self._chk(self.b.last_shrinkage(self.h, C.byref(v)))
        return v.value

    def cholesky(self, A):
        Aa = np.asarray(A, dtype=np.float64)
        n = Aa.shape[0]
        L = np.zeros(n * n)
        self._chk(self.b.cholesky(self.h, _d(_colmajor(Aa)), n, _d(L)))
        return L.reshape((n, n), order="F")

    def inv_sqrt(self, A):
        Aa = np.asarray(A, dtype=np.float64)
        n = Aa.shape[0]
        Cm = np.zeros(n * n)
        self._chk(self.b.inv_sqrt(self.h, _d(_colmajor(Aa)), n, _d(Cm)))
        return Cm.reshape((n, n), order="F")

    # -- product-only: device-resident loop, comm, introspection -------------------------------------
    def resident_reset(self, state, env_t, U):
        s, u = _f64(state), _f64(U)
        self._chk(self.b.resident_reset(self.h, _d(s), int(env_t), _d(u)))

    def resident_plan(self, advance_env=True):
        self._chk(self.b.resident_plan(self.h, int(advance_env)))

    def resident_read(self):
        s, u, c = np.zeros(self.ss), np.zeros(self.cs), np.zeros(self.as_)
        its = C.c_int32(0)
        self._chk(self.b.resident_read(self.h, _d(s), _d(u), _d(c), C.byref(its)))
        return s, u, c, its.value

    def resident_total_its(self) -> int:
        v = C.c_int64(0)
        self._chk(self.b.resident_total_its(self.h, C.byref(v)))
        return v.value

    def resident_reward_sum(self) -> float:
        v = C.c_double(0.0)
        self._chk(self.b.resident_reward_sum(self.h, C.byref(v)))
        return v.value

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        self._chk(self.b.measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def bench_rowsum(self, reps=20):
        ms, nbytes = C.c_double(0.0), C.c_double(0.0)
        self._chk(self.b.bench_rowsum(self.h, int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def sortperm(self, costs):
        c = _f64(costs)
        out = np.zeros(c.size, dtype=np.int64)
        self._chk(self.b.sortperm(self.h, _d(c), c.size, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def elite_select(self, costs, m, k0=0, kloc=None, early_stop=True):
        """(ascending elite ids inside [k0, k0+kloc), stop, tau4) — select.cu's sort-free order[1:m] + POL:458-461."""
        c = _f64(costs)
        kloc = c.size - k0 if kloc is None else kloc
        ids = np.zeros(min(m, kloc) + 1, dtype=np.int64)
        n, stop, tau = C.c_int64(0), C.c_int32(0), np.zeros(4)
        self._chk(self.b.elite_select(self.h, _d(c), c.size, int(m), int(k0), int(kloc), int(bool(early_stop)),
                                      ids.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n), C.byref(stop), _d(tau)))
        return ids[:n.value].copy(), bool(stop.value), tau

    def comm_init(self, nccl_id: bytes):
        buf = C.create_string_buffer(nccl_id, 128)
        self._chk(self.b.comm_init(self.h, buf))

    def comm_init_loopback(self, group):
        """Attach this handle (virtual rank) to a loop-back group (`_lib.LoopbackGroup`)."""
        self._chk(self.b.comm_init_loopback(self.h, group.ptr))

    def comm_peer_export(self) -> bytes:
        """128-byte blob (CUDA IPC handles of this rank's control region and cost vector) for comm_peer_attach."""
        buf = C.create_string_buffer(128)
        self._chk(self.b.comm_peer_export(self.h, buf))
        return buf.raw

    def comm_peer_attach(self, blobs):
        """Switch the per-iteration exchanges to peer-memory kernels; `blobs`: every rank's export, in rank order."""
        raw = b"".join(blobs)
        self._chk(self.b.comm_peer_attach(self.h, C.create_string_buffer(raw, len(raw)), len(raw)))

    def comm_peer_loopback(self):
        """The same between the virtual ranks of a loop-back group (collective: every rank's thread calls it; needs
        CUDA_MODULE_LOADING=EAGER in the environment before CUDA initialises)."""
        self._chk(self.b.comm_peer_loopback(self.h))

    def warp_cycles(self) -> np.ndarray:
        """Per-warp clock64() cycles of the most recent rollout launch (needs set_option("rollout_profile", 1))."""
        n = 8 * (self.Kloc // 32 + 2)  # variant 4/5: [total, waiting] per warp (3 per 64 rollouts) + phase clocks
        out = np.zeros(n, dtype=np.int64)
        self._chk(self.b.warp_cycles(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)), n))
        return out

    def launch_count(self) -> int:
        return int(self.b.launch_count(self.h))

    def last_timing(self):
        r, t, n = C.c_double(0), C.c_double(0), C.c_int32(0)
        self._chk(self.b.last_timing(self.h, C.byref(r), C.byref(t), C.byref(n)))
        return {"rollout_ms": r.value, "total_ms": t.value, "rollout_launches": n.value}
