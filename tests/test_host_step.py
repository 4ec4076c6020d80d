"""CPU check of the device arithmetic: csrc/car_model.cuh is __host__ __device__, so nvcc can build a host
program (tools/host_step_check.cu) that runs the SAME step functions the rollout kernels run and compares the
fast formulations (MODE 0 "v3", MODE 3 "v4" speculative straight-line + repair) with the literal restatement of
CAR:282-344 (MODE 1, itself parity-tested against the oracle on the GPU) from identical random states — including
cars sliding backwards (Vx < 0), saturated tyres, steering and pedal at their limits."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path_factory.mktemp("hsc") / "host_step_check"
    subprocess.run([nvcc, "-Wno-deprecated-gpu-targets", "-O2", "-std=c++17", "-Xcompiler", "-ffp-contract=off",
                    "-I", str(ROOT / "mpopis_b200" / "csrc"), str(ROOT / "tools" / "host_step_check.cu"),
                    "-o", str(exe)], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("seed", [1, 2])
def test_fast_steps_match_the_literal_step(checker, seed):
    out = subprocess.run([str(checker), "300000", str(seed)], check=True, capture_output=True, text=True).stdout
    r = json.loads(out.splitlines()[0])
    assert r["nan_v4"] == 0
    # one control step (10 Euler sub-steps) from the same state: rounding-level agreement, FP64
    assert r["max_rel_err_v3_vs_literal"] < 5e-12
    assert r["max_rel_err_v4_vs_literal"] < 5e-12
    assert r["max_rel_err_v4_vs_v3"] < 5e-12  # v4 re-associates the Euler updates (a few ulp per sub-step)
    assert r["reversed_frac"] > 0.2          # the sample really exercises Vx < 0
    assert 0.0 < r["repaired_frac"] < 0.2    # ... and both the speculative and the repair path
