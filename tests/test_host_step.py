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
    # 25 consecutive control steps with sin/cos of δ and Ψ carried across steps (re-evaluated every 5th)
    assert r["max_rel_err_seq25_v4_vs_literal"] < 1e-10
    assert r["reversed_frac"] > 0.2          # the sample really exercises Vx < 0
    assert 0.0 < r["repaired_frac"] < 0.2    # ... and both the speculative and the repair path


def test_device_codegen_keeps_the_track_scan(tmp_path):
    """Guards the __host__ __device__ intrinsic stand-ins of car_model.cuh: the DEVICE pass must map mul_rn/add_rn to
    the un-fused intrinsics and keep the arg-min scan of within_track (a self-recursive stand-in once made the
    compiler delete the whole loop — only visible on the GPU)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    src = tmp_path / "wt.cu"
    src.write_text(f"""
#include "{ROOT / 'mpopis_b200' / 'csrc' / 'car_model.cuh'}"
using namespace mpopis;
__global__ void k(const double *x, const double *y, const double *w, int n, double px, double py, int *out, double *d) {{
  TrackView tr{{x, y, w, n, nullptr, 0, 0, 0, 0, 0}};
  int a, b;
  double dist;
  const bool wi = within_track<false>(tr, px, py, &a, &b, &dist);
  out[0] = a, out[1] = b, out[2] = wi, d[0] = dist;
}}
""")
    ptx = tmp_path / "wt.ptx"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-ptx", str(src), "-o",
                    str(ptx)], check=True, capture_output=True)
    text = ptx.read_text()
    assert text.count("mul.rn.f64") >= 4 and text.count("setp.lt.f64") >= 2, "the nearest-point scan was optimised away"
    assert "bra" in text
