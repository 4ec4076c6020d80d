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
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
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


def cpu_reference(steps, warmup, K_sample, threads):
    """The reference's CPU path (Threads.@threads over k, POL:269) as restated by oracle/ — the only CPU arm
    available: Julia is not installed here, so kind = "port"."""
    from oracle import oracle
    env, eng = make_engine(oracle.bound(), K_sample, 0, 1, 0)
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
    return K_sample * T * its_total / t_total, t_total / max(steps, 1) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples-per-gpu", type=int, default=K_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the K = 150 / 4096 / 2^20 side measurements")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    K = args.samples_per_gpu * max(world, 1)
    workload = (f"CarRacing 1-car :cemppi K={K} ({args.samples_per_gpu}/GPU) H={T} ais_its={N_ITS} λ={LAMBDA} "
                f"Σ_est=:ss early-stop on (reference defaults, car_example.jl:51-81)")
    base = {"metric": "rollout-steps/sec CarRacing :cemppi K×H", "unit": "rollout-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        K_ref = 8192  # bounded sample: CPU throughput does not depend on K; keeps the run to ~1 s per step
        v, ms = cpu_reference(args.steps, min(warmup, 1), K_ref, threads)
        sample = f"{args.steps} control steps of the same :cemppi workload at K={K_ref} (of {K}) per step"
        line = dict(base, impl="reference", value=v, ms_per_step=ms,
                    config={"workload": workload, "sample": sample},
                    cpu_baseline={"value": v, "unit": "rollout-steps/s", "cores": threads, "kind": "port", "sample": sample},
                    e2e={"value": v, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0, note="C restatement of the reference (oracle/), OpenMP over k on all host cores; "
                                         "Julia is not installed in this image")
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from mpopis_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    bound = _lib.product()  # raises if the CUDA library is missing: no fallback
    env, eng = make_engine(bound, K, rank, world, local_rank)
    if world > 1:
        ids = [_lib.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0])

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
    roll_ms_per_launch = tm["rollout_ms"] / max(tm["rollout_launches"], 1)
    alg_bytes = ALG_BYTES_PER_ROLLOUT_STEP * eng.Kloc * T
    achieved = alg_bytes / (roll_ms_per_launch * 1e-3) / 1e9
    roofline = {"kernel": "rollout_car_kernel<1,fast>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                "ms_per_launch": roll_ms_per_launch, "share_of_step": tm["rollout_ms"] / tm["total_ms"],
                "note": "the rollout kernel is FP64-issue bound (≈16 B of HBM traffic per ≈2·10³ FP64 instructions); "
                        "see roofline_fp64 for the binding roofline and DESIGN.md §5"}
    prof = ROOT / "profiles" / "rollout_kernel_metrics.json"
    fp64_peak = eng.measure_fp64_peak() if rank == 0 else 0.0
    roofline_fp64 = None
    if prof.exists():
        pm = json.loads(prof.read_text())
        roofline["traffic"] = pm.get("dram_bytes_per_launch")
        ipr = pm.get("fp64_thread_instr_per_rollout_step")
        if ipr:
            rate = ipr * eng.Kloc * T / (roll_ms_per_launch * 1e-3)
            roofline_fp64 = {"bound": "fp64-pipe", "achieved": rate / 1e12, "peak": fp64_peak / 1e12,
                             "unit": "T thread-instr/s", "frac": rate / fp64_peak if fp64_peak else None,
                             "fp64_instr_per_rollout_step": ipr, "source": "profiles/rollout_kernel_metrics.json (ncu) "
                             "× live CUDA-event launch time; peak = DFMA micro-benchmark in this run"}

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

    # the other sizes the north star names (K ∈ {150 … 2^20}, T = 50): short device-resident runs, same timing rules
    k_sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        k_sweep = []
        for Ks in (150, 4096, 1 << 20):
            env_s, eng_s = make_engine(bound, Ks, 0, 1, local_rank)
            st_s = torch.cuda.ExternalStream(eng_s.b.stream(eng_s.h), device=torch.device("cuda", local_rank))
            eng_s.resident_reset(state0, 0, np.zeros(eng_s.cs))
            for _ in range(3):
                eng_s.resident_plan(True)
            eng_s.resident_read()
            eng_s.resident_reset(state0, 0, np.zeros(eng_s.cs))
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
            k_sweep.append({"K": Ks, "ms_per_step": ms_s / n_s, "its_per_step": its_s / n_s,
                            "rollout_steps_per_s": Ks * T * its_s / (ms_s * 1e-3),
                            "rollout_ms_per_launch": tm_s["rollout_ms"] / max(tm_s["rollout_launches"], 1)})
            eng_s.close()

    line = dict(base, value=value, ms_per_step=dev_ms / args.steps,
                config={"workload": workload, "l2": "flushed between steps (256 MiB memset outside the timed intervals)",
                        "parallelism": f"sample-sharded x{world}" if world > 1 else "single GPU"},
                e2e={"value": e2e_value, "unit": "rollout-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "ms_per_step": e2e_s / args.steps * 1e3},
                gpu_launches=int(launches), clocks=clocks.summary(), roofline=roofline, roofline_fp64=roofline_fp64,
                roofline_g8=roofline_g8,
                fp64_peak_dfma_per_s=fp64_peak, its_per_step=its_total / args.steps, k_sweep=k_sweep)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        K_cpu = 8192
        v, ms = cpu_reference(3, 1, K_cpu, threads)
        line["cpu_baseline"] = {"value": v, "unit": "rollout-steps/s", "cores": threads, "kind": "port",
                                "sample": f"3 control steps of the same :cemppi workload at K={K_cpu} (of {K}), "
                                          f"oracle/ C restatement with OpenMP over k"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
