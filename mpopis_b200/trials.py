"""Independent trials as concurrent device-resident replicas.

The reference runs its trials one after the other (`for k ∈ 1:num_trials`, car_example.jl:170; mountaincar_example.jl:120)
and, inside a trial, one control step after the other with a host round trip per step. At the reference's own sizes
(K = 150 / 375 / 20) a control step keeps a B200 busy for a few percent of its time, so the natural use of the GPU —
and of 8 GPUs — is to run many trials AT ONCE: every trial gets its own engine handle (own stream, own captured CUDA
graph, own Philox key), its env lives on the device (`resident_plan(advance_env=True)` applies the control to the
resident env, car_example.jl:205-207), and the host only enqueues: step s of every trial is launched before step s + 1 of
any, so the per-trial kernels of one step overlap on the device. Trials are dealt round-robin over the given devices;
one host thread drives all of them (a handle switches the device itself). Nothing is read back until the end
(state, U, Σ reward, executed AIS iterations).
"""
from __future__ import annotations

import time

import numpy as np


def run_trial_replicas(make_engine, num_trials: int, num_steps: int, *, devices=(0,), seeds=None, concurrency=None,
                       timing: dict | None = None):
    """make_engine(device) -> (env, engine) with the engine configured (set_*_env, set_sigma, ...).
    Returns a list of per-trial dicts {state, U, control, reward_sum, its, seed, device} and the wall time in s.
    `timing` (optional dict) receives {"create_s", "run_s"}: handle creation vs enqueueing the steps and reading back."""
    seeds = list(seeds) if seeds is not None else list(range(1, num_trials + 1))
    concurrency = num_trials if concurrency is None else max(1, int(concurrency))
    out = [None] * num_trials
    t0 = time.perf_counter()
    t_create = t_run = 0.0
    for first in range(0, num_trials, concurrency):  # waves of `concurrency` resident trials
        batch = []
        tc = time.perf_counter()
        for i in range(first, min(num_trials, first + concurrency)):
            dev = devices[i % len(devices)]
            env, eng = make_engine(dev)
            eng.seed(seeds[i])
            eng.resident_reset(env.state, getattr(env, "t", 0), np.zeros(eng.cs))
            batch.append((i, dev, eng))
        tr = time.perf_counter()
        t_create += tr - tc
        for _ in range(num_steps):
            for _, _, eng in batch:  # breadth first: one control step of every trial is in flight together
                eng.resident_plan(True)
        for i, dev, eng in batch:
            state, U, ctrl, _ = eng.resident_read()
            out[i] = {"state": state, "U": U, "control": ctrl, "reward_sum": eng.resident_reward_sum(),
                      "its": eng.resident_total_its(), "seed": seeds[i], "device": dev}
        t_run += time.perf_counter() - tr
        for _, _, eng in batch:
            eng.close()
    if timing is not None:
        timing["create_s"], timing["run_s"] = t_create, t_run
    return out, time.perf_counter() - t0
