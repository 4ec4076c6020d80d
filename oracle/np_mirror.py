"""Independent numpy restatement of pieces of the MPOPIS hot path, written from the Julia sources
separately from oracle/mpopis_oracle.c. TEST INFRASTRUCTURE ONLY: it exists so that a transcription
slip in the C oracle shows up as a disagreement (tests/test_oracle_*.py). It cannot catch a shared
misreading of the reference — PARITY UNPINNED applies to both.

Citations: CAR = src/envs/car_racing.jl, TRK = src/envs/car_racing_tracks/car_racing_tracks.jl,
MCR = src/envs/multi-car_racing.jl, UTL = src/utils.jl, POL = src/mppi_mpopi_policies.jl.
"""
from __future__ import annotations

import math

import numpy as np


def sign(x):
    return float(x > 0) - float(x < 0)


def tire_fy(α, μ, C, fz, fx):  # CAR:252-260
    fy_max = math.sqrt(max((μ * fz) ** 2 - fx ** 2, 1e-8))
    ta = math.tan(α)
    if abs(α) < math.atan(3 * fy_max / C):
        return -C * ta + (C ** 2 / (3 * fy_max)) * abs(ta) * ta - (C ** 3 / (27 * fy_max ** 2)) * ta ** 3
    return -fy_max * sign(α)


def car_step(P, dt, δt, s, a):  # CAR:282-344; P = 18 params in declaration order
    m, Izz, h, l_f, l_r, CD0, CD1, Cf, Cr, μf, μr, δmax, δdmax, Fxmax, Fxmin, λb, λd, _ = P
    x, y, Ψ, Vx, Vy, Ψd, δ = s[:7]
    rate = min(abs(a[0] * δmax - δ) / dt, δdmax) * sign(a[0] * δmax - δ)
    pedal = a[1]
    for _ in range(int(np.rint(dt / δt))):
        δ += rate * δt
        αf = math.atan2(Vy + l_f * Ψd, Vx) - δ
        αr = math.atan2(Vy - l_r * Ψd, Vx)
        aero = (CD0 + CD1 * abs(Vx)) * sign(Vx)
        fx = Fxmax * max(pedal, 0.0) + Fxmin * min(pedal, 0.0) * sign(Vx)
        split = λb if pedal <= 0 else λd
        fxf, fxr = split * fx, (1 - split) * fx
        L = l_r + l_f
        fzf = (m * l_r * 9.81 - h * fx) / L
        fzr = (m * l_f * 9.81 + h * fx) / L
        fyf = tire_fy(αf, μf, Cf, fzf, fxf)
        fyr = tire_fy(αr, μr, Cr, fzr, fxr)
        Ψdd = (1 / Izz) * (l_f * (fxf * math.sin(δ) + fyf * math.cos(δ)) - l_r * fyr)
        Vyd = (1 / m) * (fyf * math.cos(δ) + fxf * math.sin(δ) + fyr) - Ψd * Vx
        Vxd = (1 / m) * (fxf * math.cos(δ) - fyf * math.sin(δ) + fxr - aero) + Ψd * Vy
        Ψd += Ψdd * δt
        Vx += Vxd * δt
        Vy += Vyd * δt
        Ψ += Ψd * δt
        Ψ = math.atan2(math.sin(Ψ), math.cos(Ψ))
        x += (Vx * math.cos(Ψ) - Vy * math.sin(Ψ)) * δt
        y += (Vx * math.sin(Ψ) + Vy * math.cos(Ψ)) * δt
    return np.array([x, y, Ψ, Vx, Vy, Ψd, δ, pedal])


def within_track(tx, ty, tw, pos):  # TRK:68-92 (0-based indices)
    d = (tx - pos[0]) ** 2 + (ty - pos[1]) ** 2
    i = int(np.argmin(d))  # first minimum
    n = len(tx)
    im, ip = (i - 1) % n, (i + 1) % n
    dm = math.hypot(tx[im] - pos[0], ty[im] - pos[1])
    dp = math.hypot(tx[ip] - pos[0], ty[ip] - pos[1])
    j = im if dm <= dp else ip
    p1, p2, p3 = np.array([tx[i], ty[i]]), np.array([tx[j], ty[j]]), np.asarray(pos, dtype=float)
    t = np.dot(p3 - p1, p2 - p1) / np.dot(p2 - p1, p2 - p1)
    dist = float(np.linalg.norm(p1 + t * (p2 - p1) - p3))
    return i, j, dist, dist < tw[i]


def car_reward(P, tx, ty, tw, s):  # CAR:201-213
    _, _, dist, within = within_track(tx, ty, tw, s[:2])
    rew = 0.0
    if not within:
        rew += -1000000.0
    if abs(math.atan2(s[4], s[3])) > P[17]:
        rew += -5000.0
    return rew - dist + 2.0 * math.hypot(s[3], s[4])


def multicar_reward(Ps, tx, ty, tw, s):  # MCR:145-158
    N = len(Ps)
    cars = np.asarray(s).reshape(N, 8)
    rew = 0.0
    for i in range(N):
        rew += car_reward(Ps[i], tx, ty, tw, cars[i])
        for j in range(i + 1, N):
            Δd = math.hypot(cars[j, 0] - cars[i, 0], cars[j, 1] - cars[i, 1])
            rew += -Δd - (11000.0 if Δd <= 4.0 else 0.0)
    return rew


def rollout_cost(Ps, dt, δt, tx, ty, tw, state, V):  # UTL:129-144 + POL:271-274 (γ = 0)
    N = len(Ps)
    s = np.asarray(state, dtype=float).copy().reshape(N, 8)
    A = np.clip(np.asarray(V).reshape(-1, 2 * N), -1.0, 1.0)
    cost = 0.0
    for a in A:
        for c in range(N):
            s[c] = car_step(Ps[c], dt, δt, s[c], a[2 * c:2 * c + 2])
        cost -= multicar_reward(Ps, tx, ty, tw, s.reshape(-1)) if N > 1 else car_reward(Ps[0], tx, ty, tw, s[0])
    return cost


def mountaincar_rollout(mc, max_steps, state, t0, V):  # RLEnvs _step! + EXM:10-22
    min_pos, max_pos, max_speed, goal_pos, goal_vel, power, gravity = mc
    x, v, t, cost = state[0], state[1], t0, 0.0
    for a in np.clip(V, -1.0, 1.0):
        t += 1
        v += a * power + math.cos(3 * x) * (-gravity)
        v = min(max(v, -max_speed), max_speed)
        x += v
        x = min(max(x, min_pos), max_pos)
        if x == min_pos and v < 0:
            v = 0
        done = (x >= goal_pos and v >= goal_vel) or t >= max_steps
        rew = (100000 if (x >= goal_pos and v >= goal_vel) else 0) + abs(v) + (0.0 if done else -1.0)
        cost -= rew
    return cost


def weights(costs, λ):  # UTL:79-86
    w = np.exp(-1 / λ * (costs - np.min(costs)))
    return w / np.sum(w)


def cov_estimate(X, method):
    """cov(method, X') with X = p x n columns; CovarianceEstimation.jl semantics."""
    p, n = X.shape
    μ = X.mean(axis=1)
    Xc = (X - μ[:, None]).T  # n x p
    S = Xc.T @ Xc / n
    if method == "mle":
        return μ, S, 0.0
    if method in ("lw", "ss"):
        if method == "ss":
            d = 1 / np.sqrt(np.diag(S))
            Z, R = Xc * d, S * np.outer(d, d)
        else:
            Z, R = Xc, S
        W2 = (Z ** 2).T @ (Z ** 2)  # Σ_k (z_ki z_kj)²
        off = ~np.eye(p, dtype=bool)
        num = (W2[off] - n * R[off] ** 2).sum() * n / ((n - 1) * n ** 2)
        λ = float(np.clip(num / (R[off] ** 2).sum(), 0, 1))
        return μ, (1 - λ) * S + λ * np.diag(np.diag(S)), λ
    trS2, tr2S = float((S ** 2).sum()), float(np.trace(S) ** 2)
    if method == "rblw":
        λ = ((n - 2) / n * trS2 + tr2S) / ((n + 2) * (trS2 - tr2S / p))
    else:
        λ = ((1 - 2 / p) * trS2 + tr2S) / ((n + 1 - 2 / p) * (trS2 - tr2S / p))
    λ = float(np.clip(λ, 0, 1))
    return μ, (1 - λ) * S + λ * np.trace(S) / p * np.eye(p), λ


def ce_plan(Ps, dt, δt, tx, ty, tw, state, U, Σ, Z, λ, N, elite_thr, method, early_stop=True):
    """calculate_trajectory_costs(::CEMPPI_Policy) POL:434-472 + functor POL:221-238 (α = 1)."""
    cs, K = Z.shape[0], Z.shape[1]
    m = int(np.rint(K * (1 - elite_thr)))
    U_orig, U_cur, Σp = U.copy(), U.copy(), Σ.copy()
    its = 0
    for n in range(1, N + 1):
        its = n
        E = np.linalg.cholesky(Σp) @ Z[:, :, n - 1]
        costs = np.array([rollout_cost(Ps, dt, δt, tx, ty, tw, state, U_cur + E[:, k]) for k in range(K)])
        if n < N:
            order = np.argsort(costs, kind="stable")
            ec = costs[order[:m]]
            if early_stop and np.max(np.abs(np.diff(ec))) < 10e-3:
                break
            μ, S, _ = cov_estimate(E[:, order[:m]], method)
            Σp = S + 10e-9 * np.eye(cs)
            U_cur = U_cur + μ
    E = E + (U_cur - U_orig)[:, None]
    w = weights(costs, λ)
    wc = U_orig + E @ w
    as_ = 2 * len(Ps)
    control = np.clip(wc[:as_], -1, 1)
    U_next = U_orig.copy()
    U_next[:-as_] = wc[as_:]
    return control, U_next, its, costs, w
