#!/usr/bin/env python
"""bench.py — headline benchmark of the MPPI/MPOPI hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one control step `pol(env)` (POL:221-238) of CarRacing :cemppi, T = 50, ais_its = 10,
λ = 10, Σ_est = :ss (simulate_car_racing's defaults, car_example.jl:51-81) with K = 65 536 samples
PER GPU (the north-star target configuration at N = 1; weak scaling: K = 65 536·N sharded by sample
across N GPUs, one NCCL all-gather of the costs + all-reduces of the elite moments per AIS
iteration). Metric: rollout-steps/s = K·T·(AIS iterations executed) / time.

  value : device-resident loop (state, U and the env stay in HBM; the control is applied to the
          resident env so consecutive steps differ), CUDA events on the engine's stream, L2 flushed
          between steps (outside the timed intervals), max over ranks.
  e2e   : the same steps through the public C-ABI call a user makes (mpopis_b200_plan) with HOST
          buffers: H2D of state + U and D2H of control + U + status inside the timed region.
  --scaling strong --total-samples 1048576 : BASELINE config 5 — K = 2^20 fixed, sharded over the N GPUs.
  N > 1 lines carry "parity": one seeded control step at K = 8192·N, sharded vs unsharded on rank 0 (outside the
  timed region): max relative control / U difference, identical AIS iteration counts, ranks bit-identical.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

T, N_ITS, LAMBDA, K_PER_GPU = 50, 10, 10.0, 65536
ALG_BYTES_PER_ROLLOUT_STEP = 8 * 2 + 8.0 / T  # reads E[:,k] (as = 2 doubles per step), writes cost_k


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (nvidia-smi's fields via NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        alt = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
               "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k in names:
                    bit = getattr(nv, names[k], None) or getattr(nv, alt[k], None)
                    if bit and mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.th.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# the files whose text IS the rollout kernels (engine.cuh is left out on purpose: it declares every launcher of the library,
# so unrelated signature changes would invalidate the stamp; the kernel parameter structs it holds changed last in round 1)
ROLLOUT_SOURCES = ["mpopis_b200/csrc/rollout.cu", "mpopis_b200/csrc/rollout_split.cu", "mpopis_b200/csrc/rollout_kernels.cuh",
                   "mpopis_b200/csrc/car_model.cuh"]


def rollout_source_hash() -> str:
    """sha256 over the sources that define the rollout kernels. profiles/rollout_kernel_metrics.json (instruction
    counts from an ncu capture) is stamped with it by tools/update_rollout_metrics.py; a stale stamp means the
    counts describe another kernel and roofline_fp64 is refused rather than silently wrong."""
    h = hashlib.sha256()
    for rel in ROLLOUT_SOURCES:
        h.update((ROOT / rel).read_bytes())
    return h.hexdigest()[:16]


def make_engine(bound, K, rank, world, device, early_stop=True, **extra):
    from mpopis_b200 import _abi
    from mpopis_b200.engine import Engine
    from mpopis_b200.envs import CarRacingEnv
    from mpopis_b200.policies import block_diagm
    env = CarRacingEnv()
    eng = Engine(bound, policy="cemppi", env=_abi.ENV_CAR_RACING, num_samples=K, horizon=T, opt_its=N_ITS,
                 lam=LAMBDA, alpha=1.0, ce_elite_threshold=0.8, sigma_est="ss", early_stop=early_stop,
                 device=device, rank=rank, world_size=world, **extra)
    env.configure_engine(eng)
    eng.set_sigma(block_diagm([0.0625, 0.1], 1))
    eng.seed(20260917)
    return env, eng


CPU_NOTE = ("C restatement of the reference (oracle/mpopis_oracle.c, gcc -O2 -fopenmp -ffp-contract=off), OpenMP over k on "
            "all host cores like Threads.@threads POL:269; per sample it copies 8 doubles of env state where the "
            "reference deep-copies the whole env incl. its track (POL:270), so it is a conservative (fast) baseline; "
            "Julia is not installed in this image")


def cpu_reference(steps, warmup, K_sample, threads, make=None):
    """The reference's CPU path (Threads.@threads over k, POL:269) as restated by oracle/ — the only CPU arm
    available: Julia is not installed here, so kind = "port"."""
    from oracle import oracle
    env, eng = (make or (lambda b, K: make_engine(b, K, 0, 1, 0)))(oracle.bound(), K_sample)
    eng.b.set_threads(eng.h, threads)
    state, U = env.state.copy(), np.zeros(eng.cs)
    its_total, t_total = 0, 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ctrl, U, its = eng.plan(state, i, U)
        dt = time.perf_counter() - t0
        state, _, _, _ = eng.env_step(state, ctrl, i)
        if i >= warmup:
            its_total += its
            t_total += dt
    eng.close()
    return K_sample * T * its_total / t_total, t_total / max(steps, 1) * 1e3


def make_sweep_engine(bound, label, K, device=0):
    """C2 = :cemppi K=150, C3 = :μΣaismppi K=4096 ais_its=5, C4 = 3-car :cmamppi K=375 (car_example.jl defaults)."""
    from mpopis_b200 import _abi
    from mpopis_b200.engine import Engine
    from mpopis_b200.envs import CarRacingEnv, MultiCarRacingEnv
    from mpopis_b200.policies import block_diagm, cma_constants
    if label == "C3":
        env = CarRacingEnv()
        eng = Engine(bound, policy="μΣaismppi", env=_abi.ENV_CAR_RACING, num_samples=K, horizon=T, opt_its=5, lam=LAMBDA,
                     alpha=1.0, lambda_ais=20.0, device=device)
        env.configure_engine(eng)
        eng.set_sigma(block_diagm([0.0625, 0.1], 1))
    elif label == "C4":
        env = MultiCarRacingEnv(3)
        eng = Engine(bound, policy="cmamppi", env=_abi.ENV_CAR_RACING, n_cars=3, num_samples=K, horizon=T, opt_its=N_ITS,
                     lam=LAMBDA, alpha=1.0, device=device)
        env.configure_engine(eng)
        eng.set_sigma(block_diagm([0.0625, 0.1], 3))
        c = cma_constants(K, eng.cs, 0.8)
        eng.set_cma(sigma=0.75, m_elite=c["m_elite"], mu_eff=c["μ_eff"], c_sigma=c["cσ"], d_sigma=c["dσ"],
                    c_Sigma=c["cΣ"], c1=c["c1"], c_mu=c["cμ"], E_norm=c["E"], ws=c["ws"])
    else:
        return make_engine(bound, K, 0, 1, device)
    eng.seed(20260917)
    return env, eng


def measure_trial_replicas(make, devices, trials=16, steps=30):
    """SURVEY §8f-2 in numbers: `trials` independent trials of BASELINE config C2 (K = 150, the reference's own size), each
    `steps` control steps with the env resident on the device, run one after the other (the reference's
    `for k ∈ 1:num_trials`, car_example.jl:170) and as concurrent replicas (mpopis_b200/trials.py) over the visible
    devices. Wall-clock trials/s including the creation of every trial's handle. Never raises: a failure is reported in
    the object so that the bench line survives."""
    try:
        from mpopis_b200.trials import run_trial_replicas
        run_trial_replicas(make, min(2, trials), 3, devices=devices)  # warm-up: module load, first graph capture
        tm_seq, tm_con = {}, {}
        _, t_seq = run_trial_replicas(make, trials, steps, devices=devices[:1], concurrency=1, timing=tm_seq)
        res, t_con = run_trial_replicas(make, trials, steps, devices=devices, timing=tm_con)
        return {"config": "C2 :cemppi K=150, env resident on the device", "trials": trials, "steps_per_trial": steps,
                "devices": len(devices), "sequential_trials_per_s": trials / t_seq, "replica_trials_per_s": trials / t_con,
                "speedup": t_seq / t_con,
                "control_steps_per_s": {"sequential": trials * steps / tm_seq["run_s"], "replicas": trials * steps / tm_con["run_s"]},
                "speedup_stepping_only": tm_seq["run_s"] / tm_con["run_s"],
                "handle_creation_s_per_trial": tm_con["create_s"] / trials,
                "its_per_trial": float(np.mean([r["its"] for r in res])),
                "note": "wall clock; trials/s include the creation of every trial's handle, control_steps_per_s only the "
                        "stepping + read-back; replicas: one handle / stream / CUDA graph / Philox key per trial, step s of "
                        "every trial enqueued before step s + 1 of any"}
    except Exception as e:  # noqa: BLE001 - the measurement is an extra, the line must survive
        return {"error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples-per-gpu", type=int, default=K_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the K = 150 / 4096 / 2^20 side measurements")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --samples-per-gpu per GPU (default); strong: --total-samples sharded over the GPUs")
    ap.add_argument("--total-samples", type=int, default=1 << 20, help="K of --scaling strong (BASELINE config 5: 2^20)")
    ap.add_argument("--rollout-variant", type=int, default=None, help="override the engine's rollout kernel (A/B)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    if args.scaling == "strong":
        if args.total_samples % max(world, 1):
            raise SystemExit("--total-samples must be divisible by the number of GPUs")
        K = args.total_samples
    else:
        K = args.samples_per_gpu * max(world, 1)
    workload = (f"CarRacing 1-car :cemppi K={K} ({K // max(world, 1)}/GPU) H={T} ais_its={N_ITS} λ={LAMBDA} "
                f"Σ_est=:ss early-stop on (reference defaults, car_example.jl:51-81)")
    base = {"metric": "rollout-steps/sec CarRacing :cemppi K×H", "unit": "rollout-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the stated configuration itself (same K): ≈4 s per control step on 16 cores, so the number of timed steps is
        # bounded (the CPU rate does not depend on how many steps are averaged); one untimed warm-up step
        # N = 1 (K = 65 536) with up to ~20 timed steps: the full configuration. Larger K or more steps: each step is a
        # bounded sample of the K samples, sized so that the whole run stays near two minutes of CPU time (the port
        # sustains ≈ 5.5·10⁶ rollout-steps/s on 16 cores, independent of K)
        budget_s = 120.0 / max(args.steps + 1, 1)
        K_fit = max(4096, int(budget_s * 5.5e6 * (threads / 16.0) / (T * N_ITS)) // 32 * 32)
        K_ref = min(K, 131072, max(K_fit, 4096))
        if K_ref > 60000 and K == 65536:
            K_ref = K
        v, ms = cpu_reference(args.steps, 1, K_ref, threads)
        sample = (f"{args.steps} timed control steps of the stated workload at " +
                  (f"the full K={K}" if K_ref == K else f"K={K_ref} of {K} (the CPU rate does not depend on K)") +
                  ", 1 warm-up step")
        line = dict(base, impl="reference", value=v, ms_per_step=ms, warmup=1,
                    config={"workload": workload, "sample": sample, "parallelism": f"OpenMP x{threads} (host cores)"},
                    cpu_baseline={"value": v, "unit": "rollout-steps/s", "cores": threads, "kind": "port", "sample": sample},
                    e2e={"value": v, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0, note=CPU_NOTE)
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from mpopis_b200 import _lib, sharding
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    bound = _lib.product()  # raises if the CUDA library is missing: no fallback
    transport = []  # of the per-iteration exchanges: "peer" (csrc/comm.cu kernels over NVLink) or "nccl"
    def sharded_engine(Kx, **extra):
        envx, engx = make_engine(bound, Kx, rank, world, local_rank, **extra)
        if world > 1:
            transport.append(sharding.connect(engx, dist, rank, world))
        if args.rollout_variant is not None:
            engx.set_option("rollout_variant", args.rollout_variant)
        return envx, engx

    # ---------------- parity of the sharded path (N > 1), outside the timed region ----------------
    parity = None
    if world > 1:
        Kp = 8192 * world
        envp, engp = sharded_engine(Kp)
        Up, stp, res = np.zeros(engp.cs), envp.state.copy(), []
        for i in range(2):
            ctrl, Up, its = engp.plan(stp, i, Up)
            res.append((ctrl.copy(), Up.copy(), int(its)))
        everyone = [None] * world
        dist.all_gather_object(everyone, res)
        if rank == 0:
            _, one = make_engine(bound, Kp, 0, 1, local_rank)
            if args.rollout_variant is not None:
                one.set_option("rollout_variant", args.rollout_variant)
            U1, ref = np.zeros(one.cs), []
            for i in range(2):
                c1, U1, i1 = one.plan(stp, i, U1)
                ref.append((c1.copy(), U1.copy(), int(i1)))
            one.close()
            relmax = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))
            parity = {"K": Kp, "steps": 2, "what": "sharded (this run's communicator) vs unsharded engine on rank 0, same seed",
                      "max_rel_control": max(relmax(r[0], q[0]) for r, q in zip(res, ref)),
                      "max_rel_U": max(relmax(r[1], q[1]) for r, q in zip(res, ref)),
                      "its_equal": all(r[2] == q[2] for r, q in zip(res, ref)),
                      "ranks_bit_identical": all(all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
                                                     for a, b in zip(everyone[0], o)) for o in everyone[1:]),
                      "tolerance": 1e-5}
            parity["ok"] = bool(parity["max_rel_control"] <= 1e-5 and parity["max_rel_U"] <= 1e-5 and parity["its_equal"]
                                and parity["ranks_bit_identical"])
        engp.close()

    env, eng = sharded_engine(K)

    stream = torch.cuda.ExternalStream(eng.b.stream(eng.h), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    state0, U0 = env.state.copy(), np.zeros(eng.cs)
    # ---------------- value: device-resident loop ----------------
    eng.resident_reset(state0, 0, U0)
    for _ in range(warmup):
        eng.resident_plan(True)
    eng.resident_read()
    eng.resident_reset(state0, 0, U0)
    launches0 = eng.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    roll_ms, roll_launches = 0.0, 0
    barrier()
    with ClockSampler(local_rank) as clocks:
        with torch.cuda.stream(stream):
            for a, b in evs:
                flush.zero_()  # L2 flush, outside the timed interval
                a.record(stream)
                eng.resident_plan(True)
                b.record(stream)
        barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = eng.launch_count() - launches0
    its_total = eng.resident_total_its()
    eng.resident_read()
    tm = eng.last_timing()  # rollout-kernel CUDA events of the last step (live, on the engine's stream)
    dev_ms = max_over_ranks(dev_ms)
    value = K * T * its_total / (dev_ms * 1e-3)

    # ---------------- e2e: public C-ABI call with host buffers ----------------
    state, U = state0.copy(), U0.copy()
    for i in range(warmup):
        ctrl, U, its = eng.plan(state, i, U)
    state, U = state0.copy(), U0.copy()
    its_e2e = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ctrl, U, its = eng.plan(state, i, U)  # H2D state+U, D2H control+U+flags inside
        its_e2e += its
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = K * T * its_e2e / e2e_s
    h2d = 8 * (eng.ss + eng.cs) + 8
    d2h = 8 * (eng.as_ + eng.cs) + 12

    # ---------------- rooflines ----------------
    hbm_peak, peak_src = peaks()
    variant = int(eng.get_option("rollout_variant_used"))  # what the automatic choice (6) resolved to at this K
    kname = {4: "rollout_car_split_kernel<1> (v5: velocity warps + pose/reward warps)",
             5: "rollout_car_split_kernel<1> wide (v5: 1 velocity warp + 2 pose/reward warps per 32 rollouts)",
             3: "rollout_car_kernel<1,3,0> (v4: one thread per rollout)"}.get(variant, f"rollout variant {variant}")
    roll_ms_per_launch = tm["rollout_ms"] / max(tm["rollout_launches"], 1)
    alg_bytes = ALG_BYTES_PER_ROLLOUT_STEP * eng.Kloc * T
    achieved = alg_bytes / (roll_ms_per_launch * 1e-3) / 1e9
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "ms_per_launch": roll_ms_per_launch, "share_of_step": tm["rollout_ms"] / tm["total_ms"],
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "the rollout kernel is FP64-issue bound (≈16 B of HBM traffic per ≈1.4·10³ instructions); "
                        "see roofline_fp64 for the binding roofline and DESIGN.md §5"}
    # whole control step: every kernel's algorithmic HBM bytes (≈37 B per rollout-step) at the measured rate
    roofline_step_hbm = {"bytes_per_rollout_step": 37.0, "achieved": 37.0 * value / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": 37.0 * value / 1e9 / hbm_peak}
    prof = ROOT / "profiles" / "rollout_kernel_metrics.json"
    fp64_peak = eng.measure_fp64_peak() if rank == 0 else 0.0
    roofline_fp64 = None
    if prof.exists() and rank == 0:
        pm = json.loads(prof.read_text())
        ent = pm.get("variants", {}).get(str(variant))
        sha = rollout_source_hash()
        if ent is None or pm.get("rollout_source_sha256_16") != sha:
            roofline_fp64 = {"bound": "fp64-pipe", "frac": None,
                             "stale": f"profiles/rollout_kernel_metrics.json has no entry for variant {variant} of the "
                                      f"current kernel sources (stamp {pm.get('rollout_source_sha256_16')} != {sha}); "
                                      "re-run tools/update_rollout_metrics.py on a fresh ncu capture"}
        else:
            roofline["traffic"] = ent.get("dram_bytes_per_launch")
            ipr = ent["fp64_thread_instr_per_rollout_step"]
            rate = ipr * eng.Kloc * T / (roll_ms_per_launch * 1e-3)
            roofline_fp64 = {"bound": "fp64-pipe", "achieved": rate / 1e12, "peak": fp64_peak / 1e12,
                             "unit": "T thread-instr/s", "frac": rate / fp64_peak if fp64_peak else None,
                             "fp64_instr_per_rollout_step": ipr,
                             "peak_method": "DFMA micro-benchmark in this run (mpopis_b200_measure_fp64_peak: 8 independent "
                                            "DFMA chains per thread, 2048 threads/SM, best of 5); nominal 148 SM x 64 "
                                            "lanes x SM clock",
                             "source": f"ncu instruction counts of {ent.get('capture')} (kernel sources {sha}) x live "
                                       "CUDA-event launch time"}

    # the one HBM-bound kernel of the path (weighted-noise reduction, POL:226-229) on an operand larger than L2
    roofline_g8 = None
    if rank == 0 and world == 1:
        from mpopis_b200 import _abi
        from mpopis_b200.engine import Engine
        g8 = Engine(bound, policy="gmppi", env=_abi.ENV_CAR_RACING, num_samples=1 << 20, horizon=T, lam=LAMBDA,
                    device=local_rank)
        ms8, bytes8 = g8.bench_rowsum(20)
        g8.close()
        ach8 = bytes8 / (ms8 * 1e-3) / 1e9
        roofline_g8 = {"kernel": "rowsum_partial_kernel (Σ_k w_k E[r,k], K=2^20, cs=100: 839 MB > L2)", "bound": "hbm",
                       "achieved": ach8, "peak": hbm_peak, "unit": "GB/s", "frac": ach8 / hbm_peak,
                       "ms_per_launch": ms8, "peak_source": peak_src}

    # the other BASELINE.json configs and sizes the north star names (T = 50): short device-resident runs with the same
    # timing rules, the CPU port beside each (bounded: 2 timed steps; K = 2^20 would take minutes and is left out)
    k_sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        k_sweep = []
        for label, Ks in (("C2 :cemppi", 150), ("C4 3-car :cmamppi", 375), ("C3 :μΣaismppi", 4096), (":cemppi", 4096),
                          (":cemppi", 1 << 20)):
            tag = label.split()[0]
            env_s, eng_s = make_sweep_engine(bound, tag, Ks, local_rank)
            if args.rollout_variant is not None:
                eng_s.set_option("rollout_variant", args.rollout_variant)
            st0_s = env_s.state.copy()
            st_s = torch.cuda.ExternalStream(eng_s.b.stream(eng_s.h), device=torch.device("cuda", local_rank))
            eng_s.resident_reset(st0_s, 0, np.zeros(eng_s.cs))
            for _ in range(3):
                eng_s.resident_plan(True)
            eng_s.resident_read()
            eng_s.resident_reset(st0_s, 0, np.zeros(eng_s.cs))
            n_s = 10 if Ks <= 4096 else 4
            ev_s = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_s)]
            torch.cuda.synchronize()
            with torch.cuda.stream(st_s):
                for a, b in ev_s:
                    flush.zero_()
                    a.record(st_s)
                    eng_s.resident_plan(True)
                    b.record(st_s)
            torch.cuda.synchronize()
            ms_s = sum(a.elapsed_time(b) for a, b in ev_s)
            its_s = eng_s.resident_total_its()
            eng_s.resident_read()  # synchronises and completes the engine's own CUDA-event timings
            tm_s = eng_s.last_timing()
            row = {"config": f"{label} K={Ks}", "K": Ks, "ms_per_step": ms_s / n_s, "its_per_step": its_s / n_s,
                   "rollout_steps_per_s": Ks * T * its_s / (ms_s * 1e-3),
                   "rollout_ms_per_launch": tm_s["rollout_ms"] / max(tm_s["rollout_launches"], 1)}
            eng_s.close()
            if Ks <= 4096 and not args.no_cpu_baseline:
                v_c, ms_c = cpu_reference(2, 1, Ks, threads, make=lambda b, Kx, tag=tag: make_sweep_engine(b, tag, Kx))
                row["cpu_port_ms_per_step"], row["cpu_port_rollout_steps_per_s"], row["cpu_cores"] = ms_c, v_c, threads
            k_sweep.append(row)
    trial_replicas = None
    if rank == 0 and world == 1 and not args.no_sweep:
        devs = tuple(range(torch.cuda.device_count())) if local_rank == 0 else (local_rank,)
        trial_replicas = measure_trial_replicas(lambda dev: make_sweep_engine(bound, "C2", 150, dev), devs)

    line = dict(base, value=value, ms_per_step=dev_ms / args.steps,
                config={"workload": workload, "l2": "flushed between steps (256 MiB memset outside the timed intervals)",
                        "parallelism": f"sample-sharded x{world}" if world > 1 else "single GPU",
                        "collectives": (transport[-1] if transport else "none"), "rollout_variant": variant},
                e2e={"value": e2e_value, "unit": "rollout-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "ms_per_step": e2e_s / args.steps * 1e3},
                gpu_launches=int(launches), clocks=clocks.summary(), roofline=roofline, roofline_fp64=roofline_fp64,
                roofline_g8=roofline_g8, roofline_step_hbm=roofline_step_hbm, parity=parity,
                fp64_peak_dfma_per_s=fp64_peak, its_per_step=its_total / args.steps, k_sweep=k_sweep,
                trial_replicas=trial_replicas)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms = cpu_reference(2, 1, K, threads)  # the stated configuration itself: ≈4 s per control step on 16 cores
        line["cpu_baseline"] = {"value": v, "unit": "rollout-steps/s", "cores": threads, "kind": "port",
                                "ms_per_step": ms,
                                "sample": f"2 timed control steps (+1 warm-up) of the same :cemppi workload at the full K={K}",
                                "note": CPU_NOTE}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
