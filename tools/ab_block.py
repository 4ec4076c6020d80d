"""A/B: threads per CTA of the rollout kernel (K=65536 :cemppi control step, CUDA-event timings)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib
for K in (65536, 150):
    for blk in (32, 64, 128):
        env, eng = make_engine(_lib.product(), K, 0, 1, 0, early_stop=False)
        eng.set_option("rollout_block", blk)
        U = np.zeros(eng.cs)
        for i in range(4):
            ctrl, U2, its = eng.plan(env.state, i, U)
        tm = eng.last_timing()
        print(f"K={K} block={blk}: total {tm['total_ms']:.3f} ms rollout {tm['rollout_ms']:.3f} ms")
        eng.close()
