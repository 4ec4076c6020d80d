"""Short workload for ncu / MPOPIS_TRACE: BASELINE config C4 (3-car :cmamppi, K = 375, cs = 300) or C3."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_sweep_engine
from mpopis_b200 import _lib
label = sys.argv[1] if len(sys.argv) > 1 else "C4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else (375 if label == "C4" else 4096)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
env, eng = make_sweep_engine(_lib.product(), label, K)
U, st = np.zeros(eng.cs), env.state.copy()
for i in range(steps):
    ctrl, U, its = eng.plan(st, i, U)
    print("step", i, eng.last_timing()["total_ms"], "ms", its)
