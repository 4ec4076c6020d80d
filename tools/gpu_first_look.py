"""First GPU look: smoke parity, per-variant rollout parity, and rough timings."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import __graft_entry__ as ge
from mpopis_b200 import _abi, _lib
from mpopis_b200.engine import Engine
from mpopis_b200.envs import CarRacingEnv
from mpopis_b200.policies import block_diagm
from oracle import oracle

ge.smoke()
env = CarRacingEnv()
rng = np.random.default_rng(0)
K, T = 4096, 50
kw = dict(policy="gmppi", env=_abi.ENV_CAR_RACING, num_samples=K, horizon=T, lam=10.0)
gpu = Engine(_lib.product(), **kw); cpu = oracle.engine(nthreads=8, **kw)
for e in (gpu, cpu):
    env.configure_engine(e); e.set_sigma(block_diagm([0.0625, 0.1], 1))
E = rng.standard_normal((gpu.cs, K)) * np.tile([0.25, 0.316], T)[:, None]
U = rng.uniform(-0.3, 0.3, gpu.cs)
t0 = time.time(); cc = cpu.rollout_costs(env.state, 0, U, U, E); tc = time.time() - t0
for var in (0, 1, 2):
    gpu.set_option("rollout_variant", var)
    cg = gpu.rollout_costs(env.state, 0, U, U, E)
    rel = np.abs(cg - cc) / np.maximum(1, np.abs(cc))
    print(f"variant {var}: max rel {rel.max():.3e} median {np.median(rel):.3e} n>1e-9: {(rel>1e-9).sum()}  cpu 8thr {tc*1e3:.1f} ms = {K*T/tc:.3e} rs/s")
gpu.set_option("rollout_variant", 0)
# timing
for K in (150, 65536, 262144):
    for var in (0, 1, 2):
        g = Engine(_lib.product(), policy="cemppi", env=_abi.ENV_CAR_RACING, num_samples=K, horizon=T, opt_its=10, lam=10.0, sigma_est="ss", early_stop=False)
        env.configure_engine(g); g.set_sigma(block_diagm([0.0625, 0.1], 1)); g.seed(1)
        g.set_option("rollout_variant", var)
        U0 = np.zeros(g.cs)
        for _ in range(3):
            ctrl, U2, its = g.plan(env.state, 0, U0)
        t0 = time.time(); ctrl, U2, its = g.plan(env.state, 0, U0); wall = time.time() - t0
        tm = g.last_timing()
        print(f"K={K} var={var} its={its} wall {wall*1e3:.2f} ms total {tm['total_ms']:.2f} ms rollout {tm['rollout_ms']:.2f} ms -> {K*T*its/(tm['total_ms']*1e-3):.3e} rollout-steps/s (rollout-only {K*T*its/(tm['rollout_ms']*1e-3):.3e}) ctrl={ctrl}")
        g.close()
