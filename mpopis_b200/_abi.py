"""ctypes description of include/mpopis_b200.h.

`bind(lib, prefix)` attaches argtypes/restypes for every entry point. The CPU oracle
(oracle/mpopis_oracle.h, test infrastructure) mirrors the same signatures with the prefix
`orc_`, so tests drive both through the same `Engine` wrapper.
"""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 1
MAX_CARS = 8
CAR_NPARAMS = 18
MC_NPARAMS = 7

OK, ERR_BAD_ARG, ERR_CUDA, ERR_NCCL, ERR_NOT_PD, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5

POLICY = {
    "mppi": 0, "gmppi": 1, "imppi": 2, "cemppi": 3, "cmamppi": 4,
    "μaismppi": 5, "μΣaismppi": 6, "pmcmppi": 7,
    # ASCII aliases for the Unicode symbols of example_utils.jl:87,100
    "muaismppi": 5, "musigmaaismppi": 6,
}
ENV_CAR_RACING, ENV_MOUNTAIN_CAR, ENV_EXTERNAL = 0, 1, 2
SIGMA_EST = {"mle": 0, "lw": 1, "ss": 2, "rblw": 3, "oas": 4}


class Cfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("policy", C.c_int32), ("env", C.c_int32), ("n_cars", C.c_int32),
        ("num_samples", C.c_int64), ("horizon", C.c_int64), ("opt_its", C.c_int64),
        ("lambda_", C.c_double), ("alpha", C.c_double), ("lambda_ais", C.c_double),
        ("ce_elite_threshold", C.c_double),
        ("sigma_est", C.c_int32), ("early_stop", C.c_int32), ("log_trajectories", C.c_int32),
        ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
        ("ext_action_size", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class Cma(C.Structure):
    _fields_ = [
        ("sigma", C.c_double), ("m_elite", C.c_int64), ("mu_eff", C.c_double), ("c_sigma", C.c_double),
        ("d_sigma", C.c_double), ("c_Sigma", C.c_double), ("c1", C.c_double), ("c_mu", C.c_double),
        ("E_norm", C.c_double),
    ]


_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)
_U8 = C.POINTER(C.c_uint8)
_H = C.c_void_p
# mpopis_rollout_fn: int (*)(void *user, const double *controls, int64 K, int64 as, int64 T, double *traj_cost_out)
ROLLOUT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _D, C.c_int64, C.c_int64, C.c_int64, _D)

# name -> (restype, argtypes). `h` = opaque handle.
SIGNATURES = {
    "last_error": (C.c_char_p, []),
    "create": (C.c_int, [C.POINTER(Cfg), C.POINTER(_H)]),
    "destroy": (C.c_int, [_H]),
    "set_car_env": (C.c_int, [_H, C.c_int32, _D, C.c_double, C.c_double, _D, _D, _D, C.c_int64]),
    "set_mountaincar_env": (C.c_int, [_H, _D, C.c_int64]),
    "set_external_env": (C.c_int, [_H, _D, _D]),
    "plan_external": (C.c_int, [_H, _D, ROLLOUT_FN, C.c_void_p, _D, _D, _D, _I32]),
    "set_sigma": (C.c_int, [_H, _D, C.c_int64]),
    "set_cma": (C.c_int, [_H, C.POINTER(Cma), _D, C.c_int64]),
    "seed": (C.c_int, [_H, C.c_uint64]),
    "plan": (C.c_int, [_H, _D, C.c_int64, _D, _D, _I32]),
    "plan_with_noise": (C.c_int, [_H, _D, C.c_int64, _D, _D, _D, _D, _I32]),
    "fetch": (C.c_int, [_H, _D, _D, _D, _D]),
    "fetch_proposal": (C.c_int, [_H, _D, _D]),
    "rollout_costs": (C.c_int, [_H, _D, C.c_int64, _D, _D, _D, _D, _D]),
    "weights": (C.c_int, [_H, _D, C.c_int64, C.c_double, _D]),
    "track_query": (C.c_int, [_H, _D, C.c_int64, _I32, _I32, _D, _U8]),
    "env_step": (C.c_int, [_H, _D, _D, _I64, _D, _U8]),
    "env_reward": (C.c_int, [_H, _D, C.c_uint8, _D]),
    "sample_normals": (C.c_int, [_H, C.c_int64, C.c_int64, _D]),
    "cov_estimate": (C.c_int, [_H, C.c_int32, _D, C.c_int64, C.c_int64, _D, C.c_int32, _D, _D]),
    "cholesky": (C.c_int, [_H, _D, C.c_int64, _D]),
    "inv_sqrt": (C.c_int, [_H, _D, C.c_int64, _D]),
    "last_shrinkage": (C.c_int, [_H, _D]),
}
# entry points only the product library has
PRODUCT_ONLY = {
    "abi_version": (C.c_int, []),
    "comm_id": (C.c_int, [C.c_void_p]),
    "comm_init": (C.c_int, [_H, C.c_void_p]),
    "loopback_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "loopback_destroy": (C.c_int, [C.c_void_p]),
    "comm_init_loopback": (C.c_int, [_H, C.c_void_p]),
    "comm_peer_export": (C.c_int, [_H, C.c_void_p]),
    "comm_peer_attach": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "comm_peer_loopback": (C.c_int, [_H]),
    "set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "get_option": (C.c_int, [_H, C.c_char_p, _D]),
    "resident_reset": (C.c_int, [_H, _D, C.c_int64, _D]),
    "resident_plan": (C.c_int, [_H, C.c_int32]),
    "resident_read": (C.c_int, [_H, _D, _D, _D, _I32]),
    "resident_total_its": (C.c_int, [_H, _I64]),
    "resident_reward_sum": (C.c_int, [_H, _D]),
    "measure_fp64_peak": (C.c_int, [_H, _D]),
    "sortperm": (C.c_int, [_H, _D, C.c_int64, _I64]),
    "elite_select": (C.c_int, [_H, _D, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _I64, _I64, _I32, _D]),
    "bench_rowsum": (C.c_int, [_H, C.c_int32, _D, _D]),
    "launch_count": (C.c_int64, [_H]),
    "last_timing": (C.c_int, [_H, _D, _D, _I32]),
    "stream": (C.c_void_p, [_H]),
    "warp_cycles": (C.c_int, [_H, _I64, C.c_int64]),
}
ORACLE_ONLY = {
    "set_threads": (C.c_int, [_H, C.c_int]),
}


class Bound:
    """Namespace of bound functions: b.create(...), b.plan(...), ..."""

    def __init__(self, lib, prefix: str, extra: dict):
        self.lib, self.prefix = lib, prefix
        for name, (res, args) in {**SIGNATURES, **extra}.items():
            fn = getattr(lib, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def error(self) -> str:
        msg = self.last_error()
        return msg.decode("utf-8", "replace") if msg else ""


def bind(lib, prefix: str) -> Bound:
    return Bound(lib, prefix, PRODUCT_ONLY if prefix == "mpopis_b200_" else ORACLE_ONLY)


def exported_symbols() -> list[str]:
    """Every symbol include/mpopis_b200.h declares (checked by the CPU test-suite)."""
    return ["mpopis_b200_" + n for n in {**SIGNATURES, **PRODUCT_ONLY}]
