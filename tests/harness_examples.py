"""TEST HARNESS (not product): the reference's own callers of the hot path, restated so that the policy API can be
driven end to end without Julia — src/examples/car_example.jl (EXC), mountaincar_example.jl (EXM) and example_utils.jl
(EXU): `simulate_car_racing`, `simulate_mountaincar`, `quantile_ci`. In a deployment these stay the reference's Julia
functions, unchanged ("kept verbatim as the caller"); the product's own addition on this side of the path
is `mpopis_b200.trials.run_trial_replicas` (many trials as concurrent device-resident replicas).

Same keyword arguments, defaults, console tables and bookkeeping as the reference; `pol(env)`,
`env(act)` and `reward(env)` run on the GPU. Plotting / GIF output (Plots.jl) is out of scope
(SURVEY §2): plot_steps / save_gif raise. Each function additionally RETURNS the per-trial
statistics it prints, which the reference does not.
"""
from __future__ import annotations

import math
import random
import sys
import time
from statistics import NormalDist

import numpy as np

from mpopis_b200.envs import CarRacingEnv, MountainCarEnv, MultiCarRacingEnv, calculate_β, exceed_β, reward, within_track
from mpopis_b200.policies import block_diagm, get_policy


def quantile_ci(x, p=0.05, q=0.5):
    """quantile_ci (EXU:2-10)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    zm = NormalDist().inv_cdf(p / 2)
    zp = NormalDist().inv_cdf(1 - p / 2)
    j = max(int(math.ceil(n * q + zm * math.sqrt(n * q * (1 - q)))), 1)
    k = min(int(math.ceil(n * q + zp * math.sqrt(n * q * (1 - q)))), n)
    xs = np.sort(x)
    return xs[j - 1], float(np.quantile(x, q)), xs[k - 1]


def _p(fmt, *a, out=sys.stdout):
    out.write(fmt % a)


def _summary_rows(cols, out):
    """The AVE/STD/MED/L95/U95/MIN/MAX block (EXC:328-410, EXM:170-200)."""
    def std(v):
        return float(np.std(v, ddof=1)) if len(v) > 1 else float("nan")
    rows = [("AVE", np.mean), ("STD", std), ("MED", lambda v: quantile_ci(v)[1]), ("L95", lambda v: quantile_ci(v)[0]),
            ("U95", lambda v: quantile_ci(v)[2]), ("MIN", np.min), ("MAX", np.max)]
    table = {}
    for name, fn in rows:
        vals = [float(fn(c)) for c in cols.values()]
        table[name] = dict(zip(cols.keys(), vals))
        _p("Trials %3s: " + " : ".join(["%12.2f"] * len(vals)) + "\n", name, *vals, out=out)
    return table


def simulate_car_racing(*, num_trials=1, num_steps=200, num_cars=1, policy_type="cemppi", laps=2, num_samples=150,
                        horizon=50, λ=10.0, α=1.0, U0=None, cov_mat=None, ais_its=10, λ_ais=20.0,
                        ce_elite_threshold=0.8, ce_Σ_est="ss", cma_σ=0.75, cma_elite_threshold=0.8,
                        state_x_sigma=0.0, state_y_sigma=0.0, state_ψ_sigma=0.0, seed=None, log_runs=True,
                        plot_steps=False, pol_log=False, plot_traj=False, plot_traj_perc=1.0, text_with_plot=True,
                        text_on_plot_xy=(80.0, -60.0), save_gif=False, out=sys.stdout, **engine_kwargs):
    """simulate_car_racing(; kwargs...) EXC:51-416."""
    if plot_steps or save_gif:
        raise NotImplementedError("plotting / GIF output (Plots.jl recipes) is outside the B200 hot path")
    if U0 is None:
        U0 = np.zeros(num_cars * 2)  # EXC:61
    if cov_mat is None:
        cov_mat = block_diagm([0.0625, 0.1], num_cars)  # EXC:62
    if seed is None:
        seed = random.randint(1, int(10e10))  # EXC:72
    policy_type = str(policy_type).lstrip(":")
    sim_type = "mcr" if num_cars > 1 else "cr"
    _p("\n%-30s%s\n%-30s%d\n%-30s%d\n%-30s%d\n%-30s%d\n%-30s%s\n%-30s%d\n%-30s%d\n%-30s%.2f\n%-30s%.2f\n",
       "Sim Type:", sim_type, "Num Cars:", num_cars, "Num Trails:", num_trials, "Num Steps:", num_steps,
       "Max Num Laps:", laps, "Policy Type:", policy_type, "Num samples", num_samples, "Horizon", horizon,
       "λ (inverse temp):", λ, "α (control cost param):", α, out=out)
    if policy_type not in ("mppi", "gmppi"):
        _p("%-30s%d\n", "# AIS Iterations:", ais_its, out=out)
        if policy_type in ("μΣaismppi", "μaismppi", "pmcmppi", "musigmaaismppi", "muaismppi"):
            _p("%-30s%.2f\n", "λ_ais (ais inverse temp):", λ_ais, out=out)
        elif policy_type == "cemppi":
            _p("%-30s%.2f\n%-30s%s\n", "CE Elite Threshold:", ce_elite_threshold, "CE Σ Est Method:", ce_Σ_est, out=out)
        elif policy_type == "cmamppi":
            _p("%-30s%.2f\n%-30s%.2f\n", "CMA Step Factor (σ):", cma_σ, "CMA Elite Perc Thres:", cma_elite_threshold, out=out)
    _p("%-30s[%.4f, ..., %.4f]\n", "U₀", U0[0], U0[-1], out=out)
    cm = np.asarray(cov_mat)
    _p("%-30s%s([%.4f %.4f; %.4f %.4f], %d)\n", "Σ", "block_diagm", cm[0, 0], cm[0, 1], cm[1, 0], cm[1, 1], num_cars, out=out)
    if num_cars == 1:
        _p("%-30s%.4f\n%-30s%.4f\n%-30s%.4f\n", "Noise, State X σ:", state_x_sigma, "Noise, State Y σ:", state_y_sigma,
           "Noise, Heading σ:", state_ψ_sigma, out=out)
    _p("%-30s%d\n\n", "Seed:", seed, out=out)
    if plot_traj:
        pol_log = True  # EXC:123-126

    stats = {k: np.zeros(num_trials) for k in ("rews", "steps", "rews_per_step", "mean_vs", "max_vs", "mean_βs",
                                               "max_βs", "β_viols", "T_viols", "C_viols", "exec_times")}
    lap_ts = [np.zeros(num_trials) for _ in range(laps)]
    _p("Trial    #: %12s : %7s: %12s", "Reward", "Steps", "Reward/Step", out=out)
    for ii in range(1, laps + 1):
        _p(" : %6s%d", "lap ", ii, out=out)
    _p(" : %7s : %7s : %7s : %7s : %7s : %7s", "Mean V", "Max V", "Mean β", "Max β", "β Viol", "T Viol", out=out)
    if sim_type == "mcr":
        _p(" : %7s", "C Viol", out=out)
    _p(" : %7s\n", "Ex Time", out=out)

    for k in range(1, num_trials + 1):
        env = CarRacingEnv() if sim_type == "cr" else MultiCarRacingEnv(num_cars)
        pol = get_policy(policy_type, env, num_samples, horizon, λ, α, U0, cov_mat, pol_log, ais_its, λ_ais,
                         ce_elite_threshold, ce_Σ_est, cma_σ, cma_elite_threshold, **engine_kwargs)
        env.seed(seed + k)
        pol.seed(seed + k)
        time_start = time.time()
        lap_time = np.zeros(laps, dtype=int)
        v_mean_log, v_max_log, β_mean_log, β_max_log = [], [], [], []
        rew, cnt, lap, prev_y = 0.0, 0, 0, 0.0
        trk_viol, β_viol, crash_viol = 0, 0, 0
        while not env.done and cnt <= num_steps:  # EXC:203 (runs num_steps+1 control steps, App. B-6)
            act = pol(env)
            env(act)
            cnt += 1
            step_rew = reward(env)
            rew += step_rew
            if sim_type == "cr":  # process noise, EXC:224-236
                env.state[0] += state_x_sigma * env.rng.standard_normal()
                env.state[1] += state_y_sigma * env.rng.standard_normal()
                δψ = state_ψ_sigma * env.rng.standard_normal()
                env.state[2] += δψ
                Vx, Vy = env.state[3], env.state[4]
                env.state[3] = math.cos(δψ) * Vx + math.sin(δψ) * Vy
                env.state[4] = -math.sin(δψ) * Vx + math.cos(δψ) * Vy
            curr_y = env.state[1]
            if sim_type == "mcr":
                cars = env.state.reshape(env.N, 8)
                curr_y = float(np.min(cars[:, 1]))
                vs = np.hypot(cars[:, 3], cars[:, 4])
                βs = np.abs(np.arctan2(cars[:, 4], cars[:, 3]))
            else:
                vs = np.array([math.hypot(env.state[3], env.state[4])])
                βs = np.array([abs(calculate_β(env))])
            v_mean_log.append(float(np.mean(vs)))
            v_max_log.append(float(np.max(vs)))
            β_mean_log.append(float(np.mean(βs)))
            β_max_log.append(float(np.max(βs)))
            if step_rew < -4000:  # EXC:257-264
                ex_β = exceed_β(env)
                within_t = within_track(env)[0] if sim_type == "cr" else within_track(env)
                β_viol += int(ex_β)
                trk_viol += int(not within_t)
                temp_rew = step_rew + ex_β * 5000 + (not within_t) * 1000000
                if temp_rew < -10500:
                    crash_viol += 1
            if sim_type == "mcr":
                d = float(np.min(np.hypot(cars[:, 0], cars[:, 1])))
            else:
                d = math.hypot(env.state[0], env.state[1])
            if prev_y < 0.0 and curr_y >= 0.0 and d <= 15.0:  # EXC:273-276
                lap += 1
                lap_time[lap - 1] = cnt
            if lap >= laps or trk_viol > 10 or β_viol > 50:
                env.done = True
            prev_y = curr_y
        seconds_ran = time.time() - time_start
        i = k - 1
        stats["rews"][i], stats["steps"][i] = rew, cnt - 1
        stats["rews_per_step"][i] = rew / (cnt - 1) if cnt > 1 else float("nan")
        stats["exec_times"][i] = seconds_ran
        for ii in range(laps):
            lap_ts[ii][i] = lap_time[ii]
        stats["mean_vs"][i], stats["max_vs"][i] = np.mean(v_mean_log), np.max(v_max_log)
        stats["mean_βs"][i], stats["max_βs"][i] = np.mean(β_mean_log), np.max(β_max_log)
        stats["β_viols"][i], stats["T_viols"][i], stats["C_viols"][i] = β_viol, trk_viol, crash_viol
        if log_runs:
            _p("Trial %4d: %12.2f : %7d: %12.2f", k, rew, cnt - 1, stats["rews_per_step"][i], out=out)
            for ii in range(laps):
                _p(" : %7d", lap_time[ii], out=out)
            _p(" : %7.2f : %7.2f : %7.2f : %7.2f : %7d : %7d", np.mean(v_mean_log), np.max(v_max_log),
               np.mean(β_mean_log), np.max(β_max_log), β_viol, trk_viol, out=out)
            if sim_type == "mcr":
                _p(" : %7d", crash_viol, out=out)
            _p(" : %7.2f\n", seconds_ran, out=out)
    _p("-----------------------------------\n", out=out)
    cols = {"Reward": stats["rews"], "Steps": stats["steps"], "Reward/Step": stats["rews_per_step"]}
    for ii in range(laps):
        cols[f"lap {ii + 1}"] = lap_ts[ii]
    cols.update({"Mean V": stats["mean_vs"], "Max V": stats["max_vs"], "Mean β": stats["mean_βs"],
                 "Max β": stats["max_βs"], "β Viol": stats["β_viols"], "T Viol": stats["T_viols"]})
    if sim_type == "mcr":
        cols["C Viol"] = stats["C_viols"]
    cols["Ex Time"] = stats["exec_times"]
    summary = _summary_rows(cols, out)
    return {"trials": {**stats, "lap_ts": lap_ts}, "summary": summary, "seed": seed}


def simulate_mountaincar(*, num_trials=1, num_steps=200, policy_type="cemppi", num_samples=20, horizon=15, λ=0.1,
                         α=1.0, U0=(0.0,), cov_mat=(1.5,), ais_its=5, λ_ais=0.1, ce_elite_threshold=0.8,
                         ce_Σ_est="mle", cma_σ=0.75, cma_elite_threshold=0.8, seed=None, log_runs=True,
                         plot_steps=False, pol_log=False, save_gif=False, out=sys.stdout, x0=None,
                         **engine_kwargs):
    """simulate_mountaincar(; kwargs...) EXM:49-207. `x0` (not in the reference) fixes the start
    position, which the reference draws from an unseeded RNG."""
    if plot_steps or save_gif:
        raise NotImplementedError("plotting / GIF output (Plots.jl recipes) is outside the B200 hot path")
    if seed is None:
        seed = random.randint(1, int(10e10))
    policy_type = str(policy_type).lstrip(":")
    _p("\n%-30s%s\n%-30s%d\n%-30s%d\n%-30s%s\n%-30s%d\n%-30s%d\n%-30s%.2f\n%-30s%.2f\n", "Sim Type:", "MountainCar",
       "Num Trails:", num_trials, "Num Steps:", num_steps, "Policy Type:", policy_type, "Num samples", num_samples,
       "Horizon", horizon, "λ (inverse temp):", λ, "α (control cost param):", α, out=out)
    if policy_type not in ("mppi", "gmppi"):
        _p("%-30s%d\n", "# AIS Iterations:", ais_its, out=out)
    _p("%-30s[%.4f, ..., %.4f]\n%-30s[%.4f]\n%-30s%d\n\n", "U₀", U0[0], U0[-1], "Σ", np.asarray(cov_mat).reshape(-1)[0],
       "Seed:", seed, out=out)
    rews, steps, rps, exec_times = (np.zeros(num_trials) for _ in range(4))
    _p("Trial    #: %12s : %7s: %12s : %7s\n", "Reward", "Steps", "Reward/Step", "Ex Time", out=out)
    for k in range(1, num_trials + 1):
        env = MountainCarEnv(continuous=True)
        if x0 is not None:
            env.reset(np.array([x0, 0.0]))
        pol = get_policy(policy_type, env, num_samples, horizon, λ, α, U0, cov_mat, pol_log, ais_its, λ_ais,
                         ce_elite_threshold, ce_Σ_est, cma_σ, cma_elite_threshold, **engine_kwargs)
        env.seed(seed + k)
        pol.seed(seed + k)
        time_start = time.time()
        rew, cnt = 0.0, 0
        while not env.done and cnt <= num_steps:  # EXM:144
            act = pol(env)
            env(act)
            cnt += 1
            rew += reward(env)
        seconds_ran = time.time() - time_start
        rews[k - 1], steps[k - 1], exec_times[k - 1] = rew, cnt - 1, seconds_ran
        rps[k - 1] = rew / (cnt - 1) if cnt > 1 else float("nan")
        if log_runs:
            _p("Trial %4d: %12.2f : %7d: %12.2f : %7.2f\n", k, rew, cnt - 1, rps[k - 1], seconds_ran, out=out)
    _p("-----------------------------------\n", out=out)
    summary = _summary_rows({"Reward": rews, "Steps": steps, "Reward/Step": rps, "Ex Time": exec_times}, out)
    return {"trials": {"rews": rews, "steps": steps, "rews_per_step": rps, "exec_times": exec_times},
            "summary": summary, "seed": seed}
