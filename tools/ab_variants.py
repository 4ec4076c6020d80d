"""A/B on one GPU, same process: rollout variant (3 = v4 thread-per-rollout, 4 = v5 warp-specialised) at several K.
CUDA events of the engine itself (last_timing): ms per control step and µs per rollout launch."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib

bound = _lib.product()
for K in [int(x) for x in (sys.argv[1:] or ["65536", "150", "4096", "262144"])]:
    for variant, spin in ((3, 0), (5, 0), (5, 1)):
        env, eng = make_engine(bound, K, 0, 1, 0)
        eng.set_option("rollout_variant", variant)
        eng.set_option("rollout_spin", spin)
        U, st = np.zeros(eng.cs), env.state.copy()
        tot, roll = [], []
        for i in range(8):
            ctrl, U, its = eng.plan(st, i, U)
            tm = eng.last_timing()
            if i >= 3:
                tot.append(tm["total_ms"]), roll.append(tm["rollout_ms"] / tm["rollout_launches"] * 1e3)
        print(f"K={K} variant={variant} spin={spin}: step {np.median(tot):.3f} ms, rollout launch {np.median(roll):.1f} µs "
              f"(min {np.min(roll):.1f}), its={its}, control={ctrl}", flush=True)
        eng.close()
