"""mpopis_b200 — a B200-native (sm_100a) MPPI/MPOPI sampling engine behind the policy API of
sisl/MPOPIS. The product is the C-ABI shared library `libmpopis_b200.so` (include/mpopis_b200.h,
hand-written CUDA kernels under csrc/); this package is the Python mirror of the reference's host
side — the same constructors, symbols and entry points a Julia user of MPOPIS.jl knows:

    from mpopis_b200 import CEMPPI_Policy, CarRacingEnv, get_policy, run_trial_replicas

There is no CPU path: constructing/calling a policy without the built library and a B200 raises.
"""
from ._abi import ABI_VERSION
from .envs import (CarRacingEnv, CarRacingEnvParams, ExternalEnv, MountainCarEnv, MountainCarEnvParams, MultiCarRacingEnv,
                   calculate_β, exceed_β, reward, state, within_track)
from .policies import (CEMPPI_Policy, CMAMPPI_Policy, GMPPI_Policy, IMPPI_Policy, MPPI_Policy, PMCMPPI_Policy,
                       action_space_size, block_diagm, cma_constants, get_policy, seed_b, μAISMPPI_Policy,
                       μΣAISMPPI_Policy)
from .tracks import Track
from .trials import run_trial_replicas

__all__ = [
    "ABI_VERSION", "MPPI_Policy", "GMPPI_Policy", "IMPPI_Policy", "CEMPPI_Policy", "CMAMPPI_Policy",
    "μAISMPPI_Policy", "μΣAISMPPI_Policy", "PMCMPPI_Policy", "Track", "CarRacingEnv", "CarRacingEnvParams",
    "MultiCarRacingEnv", "ExternalEnv", "MountainCarEnv", "MountainCarEnvParams", "within_track", "calculate_β", "exceed_β",
    "block_diagm", "action_space_size", "reward", "state", "get_policy", "seed_b", "cma_constants",
    "run_trial_replicas",
]
