"""A/B on one GPU: rollout variant (0 = v3 branchy, 3 = v4 straight-line) x E = L·Z kernel (1 = DMMA row blocks,
2 = DMMA column tiles, 3 = DMMA column split) x threads per rollout CTA x noise staging (0 = register prefetch,
1 = TMA bulk copies into a per-warp shared-memory ring), K = 65536 :cemppi control steps without early stop
(CUDA-event timings from the engine: whole step and the rollout launches inside it)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
configs = [(0, 1, 64, 0, 0), (3, 2, 64, 0, 0), (3, 2, 64, 1, 0), (3, 2, 64, 0, 12), (3, 2, 64, 0, 14), (3, 2, 64, 0, 16),
           (3, 2, 64, 0, 10), (3, 2, 128, 0, 12)]
ref_ctrl = None
for variant, apl, blk, stage, queue in configs:
    env, eng = make_engine(_lib.product(), K, 0, 1, 0, early_stop=False)
    eng.set_option("rollout_variant", variant)
    eng.set_option("apply_l", apl)
    eng.set_option("rollout_block", blk)
    eng.set_option("rollout_stage", stage)
    eng.set_option("rollout_queue", queue)
    U = np.zeros(eng.cs)
    tot, roll = [], []
    for i in range(6):
        ctrl, U2, its = eng.plan(env.state, i, U)
        tm = eng.last_timing()
        if i >= 2:
            tot.append(tm["total_ms"]), roll.append(tm["rollout_ms"] / max(tm["rollout_launches"], 1))
    if ref_ctrl is None:
        ref_ctrl = ctrl
    print(f"K={K} variant={variant} apply_l={apl} block={blk} stage={stage} queue={queue}: step {np.median(tot):.3f} ms, rollout launch "
          f"{np.median(roll) * 1e3:.1f} us, control {ctrl} (|Δ vs first config| {np.max(np.abs(ctrl - ref_ctrl)):.2e})",
          flush=True)
    eng.close()
