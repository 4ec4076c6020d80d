"""Generates tests/golden/golden_v1.npz with the CPU ORACLE (oracle/mpopis_oracle.c).

The reference ships no golden vectors and cannot run here (Julia absent), so these fixtures pin the
CUDA engine (and the oracle itself, as a regression guard) to the oracle's current outputs — they are
NOT outputs of the reference: PARITY UNPINNED. Re-generate with: python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import configure, engine_kwargs, make_env, synthetic_states  # noqa: E402

from oracle import oracle  # noqa: E402

POLICIES = ["mppi", "gmppi", "imppi", "cemppi", "cmamppi", "μaismppi", "μΣaismppi", "pmcmppi"]
ASCII = {"μaismppi": "muaismppi", "μΣaismppi": "musigmaaismppi"}


def noise(seed, shape):
    return np.random.Generator(np.random.Philox(key=seed)).standard_normal(shape)


def main():
    out = {}
    states = synthetic_states()
    out["states"] = np.array(states)
    # --- rollout costs (depth i) -----------------------------------------------------------------
    env = make_env("car")
    K, T = 64, 50
    e = configure(oracle.engine(**engine_kwargs("gmppi", env, K, T)), env, "gmppi")
    E = noise(1, (e.cs, K)) * np.tile([0.25, np.sqrt(0.1)], T)[:, None]
    U = 0.3 * np.sin(np.arange(e.cs) / 7.0)
    out["roll1_E"], out["roll1_U"] = E, U
    out["roll1_costs"] = np.array([e.rollout_costs(s, 0, U, U, E) for s in [env.state] + states[:5]])
    env3 = make_env("car", 3)
    K3 = 32
    e3 = configure(oracle.engine(**engine_kwargs("gmppi", env3, K3, T)), env3, "gmppi")
    E3 = noise(2, (e3.cs, K3)) * np.tile([0.25, np.sqrt(0.1)], 3 * T)[:, None]
    out["roll3_E"] = E3
    out["roll3_costs"] = e3.rollout_costs(env3.state, 0, np.zeros(e3.cs), np.zeros(e3.cs), E3)
    mc = make_env("mc")
    em = configure(oracle.engine(**engine_kwargs("gmppi", mc, 32, 15, lam=0.1)), mc, "gmppi")
    Em = noise(3, (15, 32)) * np.sqrt(1.5)
    out["rollmc_E"] = Em
    out["rollmc_costs"] = np.array([em.rollout_costs(mc.state, t0, np.zeros(15), np.zeros(15), Em) for t0 in (0, 192)])
    # --- within_track integer fixture ---------------------------------------------------------------
    gx, gy = np.meshgrid(np.linspace(1, 255, 40), np.linspace(-156, 141, 40))
    pos = np.stack([gx.ravel(), gy.ravel()], axis=1)
    idx, idx2, dist, within = e.track_query(pos)
    out["trk_pos"], out["trk_idx"], out["trk_idx2"], out["trk_dist"], out["trk_within"] = pos, idx, idx2, dist, within
    # --- one control step per policy (depth iii with injected noise) ----------------------------------
    for pol in POLICIES:
        tag = ASCII.get(pol, pol)
        for name, envk, K, T, N, kw in (("car", make_env("car"), 96, 20, 4, dict(lam=10.0, sigma_est="ss")),
                                        ("mc", make_env("mc"), 20, 15, 5, dict(lam=0.1, lam_ais=0.1, sigma_est="mle"))):
            eng = configure(oracle.engine(**engine_kwargs(pol, envk, K, T, N, **kw)), envk, pol)
            Z = noise(10 + len(tag), (eng.cs, K, eng.N))
            u = np.random.Generator(np.random.Philox(key=99)).uniform(size=(K, max(eng.N - 1, 1)))
            U = np.zeros(eng.cs)
            st = states[2] if name == "car" else envk.state
            ctrl, U2, its = eng.plan(st, 0, U, Z=Z, resample_u=u)
            f = eng.fetch()
            p = f"plan_{tag}_{name}_"
            out[p + "Z"], out[p + "u"], out[p + "control"], out[p + "U"] = Z.astype(np.float64), u, ctrl, U2
            out[p + "its"], out[p + "costs"], out[p + "weights"] = np.array(its), f["costs"], f["weights"]
    path = Path(__file__).resolve().parent / "golden_v1.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
