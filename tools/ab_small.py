"""A/B of the fused small-size adaptation (small_adapt.cu, option "ce_small_fused") at the reference's own sizes.
CUDA events of the engine itself (last_timing): ms per control step with the CUDA graph on (the default)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib

bound = _lib.product()
for K in [int(x) for x in (sys.argv[1:] or ["150", "300", "500"])]:
    for fused in (0, 1):
        env, eng = make_engine(bound, K, 0, 1, 0)
        eng.set_option("ce_small_fused", fused)
        U, st = np.zeros(eng.cs), env.state.copy()
        tot = []
        for i in range(12):
            ctrl, U, its = eng.plan(st, i, U)
            if i >= 4:
                tot.append(eng.last_timing()["total_ms"])
        print(f"K={K} ce_small_fused={fused}: step {np.median(tot):.3f} ms (min {np.min(tot):.3f}), its={its}, control={ctrl}", flush=True)
        eng.close()
