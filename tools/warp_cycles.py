"""Per-warp clock64() of one rollout launch (set_option("rollout_profile", 1)).
variant 3: one entry per warp. variants 4 / 5 (rollout_split.cu): per CTA 2 velocity warps + 1 (variant 4) or 2
(variant 5) pose warps, each [total cycles, cycles spent spinning on the ring's mbarriers]."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib

for K in [int(x) for x in (sys.argv[1:] or ["150", "4096", "65536"])]:
    for variant in (3, 4, 5):
        env, eng = make_engine(_lib.product(), K, 0, 1, 0)
        eng.set_option("rollout_variant", variant)
        eng.set_option("rollout_profile", 1)
        U, st = np.zeros(eng.cs), env.state.copy()
        for i in range(3):
            ctrl, U, its = eng.plan(st, i, U)
        c = eng.warp_cycles()
        if variant == 3:
            c = c[: (K + 31) // 32]
            print(f"K={K} variant 3: warps {c.size} cycles median {np.median(c):.0f} max {c.max()} min {c.min()}")
        else:
            n, wpc = (K + 63) // 64, (3 if variant == 4 else 4)
            c = c[: n * wpc * 2].reshape(n, wpc, 2)
            for role, sl in (("velocity", c[:, :2, :].reshape(-1, 2)), ("pose", c[:, 2:, :].reshape(-1, 2))):
                sl = sl[sl[:, 0] > 0]
                print(f"K={K} variant {variant} {role:8s}: warps {len(sl)} total median {np.median(sl[:, 0]):.0f} max {sl[:, 0].max()} "
                      f"({np.median(sl[:, 0]) / 50:.0f} per control step), spinning median {np.median(sl[:, 1]):.0f}")
        tm = eng.last_timing()
        print(f"    rollout launch {tm['rollout_ms'] / tm['rollout_launches'] * 1e3:.1f} us, step {tm['total_ms']:.3f} ms", flush=True)
        eng.close()
