"""Spread of per-warp run time inside ONE rollout launch (K = 65536 :cemppi, last AIS iteration of a control step):
how much of the kernel's duration is tail — warps that repair steps or scan the whole track — rather than the mean."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env, eng = make_engine(_lib.product(), K, 0, 1, 0, early_stop=False)
eng.set_option("rollout_profile", 1)
if len(sys.argv) > 2:
    eng.set_option("rollout_queue", int(sys.argv[2]))
U = np.zeros(eng.cs)
st = env.state.copy()
for i in range(12):
    ctrl, U, its = eng.plan(st, i, U)
    st, _, _, _ = eng.env_step(st, ctrl, i)
    if i in (0, 5, 11):
        c = eng.warp_cycles().astype(float)
        tm = eng.last_timing()
        q = np.percentile(c, [0, 5, 50, 95, 99, 100])
        print(f"step {i}: rollout launch {tm['rollout_ms'] / tm['rollout_launches'] * 1e3:.1f} us = "
              f"{tm['rollout_ms'] / tm['rollout_launches'] * 1e-3 * 1.965e9:.0f} cycles at 1965 MHz; per-warp cycles "
              f"min {q[0]:.0f} p5 {q[1]:.0f} median {q[2]:.0f} p95 {q[3]:.0f} p99 {q[4]:.0f} max {q[5]:.0f}; "
              f"warps > 1.1 x median: {np.mean(c > 1.1 * q[2]) * 100:.1f} %, > 1.25 x: {np.mean(c > 1.25 * q[2]) * 100:.2f} %",
              flush=True)
