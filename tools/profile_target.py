"""Short workload for ncu captures: a few control steps of the headline config."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from bench import make_engine
from mpopis_b200 import _lib
K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
env, eng = make_engine(_lib.product(), K, 0, 1, 0)
U = np.zeros(eng.cs)
st = env.state.copy()
for i in range(steps):
    ctrl, U, its = eng.plan(st, i, U)
print("ok", ctrl, its)
