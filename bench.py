#!/usr/bin/env python
"""bench.py -- Heston FFT slices/sec (param x maturity, N=4096) on 1..8 B200, CPU beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sets P]

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE.json
config 3, the batched calibration objective + finite-difference Jacobian (as J^T J / J^T r
normal-equation blocks) over P = 65,536 parameter sets x 32 maturities x 50 strikes,
N = 4096, eta = 0.25, alpha = 0.75, per GPU (weak scaling: every rank prices its own P sets;
the only collective is the NCCL all-gather of the 22-double result rows).  One step = one
hb_normal_eq call (prefix scan, direct-sum job kernel, transform job kernel: DESIGN.md 4.0) = P x 32 x (1 + 5) slice
evaluations; `value` counts all six.

Prints ONE JSON line (rank 0).  `--impl reference` instead times the reference's own CPU
implementation (oracle/_ref = its heston.cpp compiled unmodified, OpenMP over options) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S0, R, Q = 100.0, 0.05, 0.02
TRUTH = np.array([2.0, 0.04, 0.3, -0.7, 0.04])
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])  # HestonCalibrator.DEFAULT_BOUNDS
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])
N_GRID, ETA, ALPHA = 4096, 0.25, 0.75
N_MAT, N_STRIKE = 32, 50
METRIC = "Heston FFT slices/sec (param x maturity, N=4096)"
# SURVEY.md 8(d), frozen convention W1: algorithmic FLOPs of one slice (N=4096, S=50)
FLOPS_PER_SLICE = 4096 * 700 + 5 * 4096 * 12 + 12 * 50  # = 3,113,560
# algorithmic HBM bytes per slice evaluation: 40 B of parameters per set and 22 doubles out per set,
# spread over the 6 x 32 slice evaluations of that set
BYTES_PER_SLICE = (40.0 + 22 * 8.0) / (6 * N_MAT)


def surface():
    K = np.tile(np.linspace(80.0, 120.0, N_STRIKE), N_MAT)
    T = np.repeat(np.linspace(0.1, 1.0, N_MAT), N_STRIKE)
    return K, T


def sobol_sets(P, skip=0):
    from scipy.stats import qmc

    s = qmc.Sobol(d=5, seed=42)
    if skip:
        s.fast_forward(skip)
    return LB + (UB - LB) * s.random(P)


def slow_decay_sets(P, skip=0):
    """A population where nothing can be elided: sigma near 2, rho near -0.99, small kappa theta (SURVEY.md
    App. D "no decay": |phi| is still O(1e3) at the end of the grid), so every grid point of every slice runs the
    full stage B and F."""
    from scipy.stats import qmc

    lo = np.array([0.1, 0.01, 1.5, -0.99, 0.01])
    hi = np.array([0.5, 0.05, 2.0, -0.90, 0.05])
    s = qmc.Sobol(d=5, seed=11)
    if skip:
        s.fast_forward(skip)
    return lo + (hi - lo) * s.random(P)


def fd_variants(X):
    """Base + 5 forward-difference perturbed sets per row (SciPy 2-point rule with bounds)."""
    rstep = 1.4901161193847656e-08
    out = np.repeat(X[:, None, :], 6, axis=1)
    for c in range(5):
        x = X[:, c]
        h = rstep * np.where(x >= 0, 1.0, -1.0) * np.maximum(1.0, np.abs(x))
        lower, upper = x - LB[c], UB[c] - x
        xn = x + h
        violated = (xn < LB[c]) | (xn > UB[c])
        fitting = np.abs(h) <= np.maximum(lower, upper)
        h = np.where(violated & fitting, -h, h)
        h = np.where(~fitting, np.where(upper >= lower, upper, -lower), h)
        out[:, c + 1, c] = x + h
    return out.reshape(-1, 5)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(seconds_budget, n_sets_hint=None):
    """Reference CPU path (oracle/_ref: heston.cpp unmodified, price_options with OpenMP over options)
    on whole (set x 32-maturity) surfaces until `seconds_budget` is spent.  -> (slices/s, info)."""
    from oracle.oracle import Reference

    ref = Reference()
    K, T = surface()
    cores = ref.use_all_cores()
    rows = fd_variants(sobol_sets(4096)[:512])
    ref.price_surface_batch(rows[:1], K, T, S0, R, Q)  # warm
    done, t0 = 0, time.perf_counter()
    chunk = n_sets_hint or max(1, cores // 8)
    while True:
        ref.price_surface_batch(rows[done:done + chunk], K, T, S0, R, Q)
        done += chunk
        el = time.perf_counter() - t0
        if el >= seconds_budget or done + chunk > len(rows):
            break
    return done * N_MAT / el, {"cores": cores, "kind": "reference", "seconds": round(el, 3), "sets_done": done,
                               "sample": f"{done} parameter-set evaluations x {N_MAT} maturities x {N_STRIKE} strikes "
                                         f"through the reference's HestonModel::price_options (1023-point quadrature "
                                         f"per option, OpenMP over options)"}


def run_reference(args):
    """`--impl reference`: the reference's own C++ (oracle/_ref: heston.cpp compiled unmodified) on the host cores.
    A step is a time-bounded SAMPLE of the workload (whole parameter sets x 32 maturities x 50 strikes through
    HestonModel::price_options); `value` is the sampled rate, `ms_per_step` the measured duration of a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, secs, done = [], [], []
    info = None
    for i in range(args.warmup + args.steps):
        rate, info = cpu_reference_rate(per_step)
        if i >= args.warmup:
            rates.append(rate)
            secs.append(info["seconds"])
            done.append(info["sets_done"])
    value = float(np.mean(rates))
    cfg = workload_config(args.sets)
    cfg["reference_sample"] = (f"each step prices {int(np.mean(done))} of the {args.sets} parameter sets (time-bounded "
                               f"sample, {per_step:.0f} s budget); the rate is a sampled rate of the same workload")
    cfg["reference_arithmetic"] = ("the reference has no FFT: price_options is a 1023-point quadrature PER OPTION "
                                   "(51,150 CF evaluations per slice against 4,096 in fft mode); the like-for-like CPU "
                                   "number is cpu_baseline.port_fft of the default arm")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "slices/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def kernel_source_hash():
    """Content hash of the CUDA sources the library is built from (pde_b200/csrc/build.py)."""
    from pde_b200.csrc.build import source_hash

    return source_hash()


def _ncu_capture(P):
    """The committed `ncu --set full` capture of ONE launch of the bench step (profiles/r02_traffic.json).  It is
    only quoted when it was taken on THIS kernel source (content hash) and launch shape; otherwise null."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None
    t = json.load(open(path))
    if (t["sets_per_launch"], t["maturities"], t["strikes"]) != (P, N_MAT, N_STRIKE):
        return None
    return t if t.get("kernel_source_hash") == kernel_source_hash() else None


def workload_config(P):
    return {"workload": f"C3: batched calibration objective + finite-difference Jacobian (J^T J, J^T r blocks), "
                        f"{P} parameter sets x {N_MAT} maturities x {N_STRIKE} strikes per GPU, N={N_GRID}, "
                        f"eta={ETA}, alpha={ALPHA}; 6 slice evaluations per (set, maturity)",
            "sets_per_gpu": P, "maturities": N_MAT, "strikes": N_STRIKE, "n_grid": N_GRID,
            "slice_evals_per_step_per_gpu": P * N_MAT * 6, "mode": "fft",
            "population": "scrambled Sobol (seed 42) over HestonCalibrator.DEFAULT_BOUNDS: the rate is INPUT-DEPENDENT "
                          "-- the integrand of most sets decays within the first ~10 % of the grid (plan's significance "
                          "cut); a rigorous bound of that live prefix routes a set to the direct-sum kernel (cost ~ prefix) "
                          "or, above a mean prefix of 1200 points, to the transform kernel (cost ~ N log N); "
                          "value_noskip gives the same step where nothing can be or is elided",
            "l2": "L2 flushed (256 MiB write) between timed steps; per-step CUDA events on the launch stream"}


def feasible_surface(pricer):
    """C4's calibration surface: the 32 x 50 grid minus the options whose fft-mode price at the truth is below the
    generator's 0.01 floor (deep OTM at the shortest maturities: the FFT value clamps to 0 there, SURVEY.md
    App. B, so the truth itself would score the 1e10 sentinel).  -> (K, T, market)."""
    K, T = surface()
    pricer.set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    p = pricer.price_host(TRUTH[None, :])[0]
    keep = p >= 0.05
    K, T, p = K[keep], T[keep], p[keep]
    mk = np.maximum(p * (1 + 0.001 * np.random.default_rng(42).normal(size=p.size)), 0.01)
    return K, T, mk


def timed_steps(torch, dev, fn, steps, flush):
    """Per-step CUDA events on the current stream of `dev`, L2 flushed before each step -> total ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return sum(a.elapsed_time(b) for a, b in ev)


def noelide_rate(sets):
    """The bench step on the Sobol box through the no-elision build (pde_b200/csrc/build.py NOELIDE_DEFINES:
    no tail skip, no asymptotic / series stage B, no zero-aware passes) in exact mode, in a subprocess."""
    lib = os.path.join(ROOT, "pde_b200", "csrc", "libheston_b200_noelide.so")
    if not os.path.exists(lib):
        return {"error": "libheston_b200_noelide.so not built"}
    env = dict(os.environ, PDE_B200_LIB=lib)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "step_rate.py"), "--sets", str(sets),
                            "--trunc", "0", "--steps", "2", "--tag", "noelide"], env=env, capture_output=True, text=True,
                           timeout=300)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return {"value": d["neq"]["slice_evals_per_s"], "ms_per_step": d["neq"]["ms_per_step"], "sets": sets}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)[:200]}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pde_b200 import BatchPricer, launch_count, measure_fp64_peak
    from pde_b200.sharding import init_from_env, sharded_map

    rank, world, local = init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    dev = torch.device("cuda", local)
    P = args.sets
    K, T = surface()
    pricer = BatchPricer("fft", N_GRID, ETA, ALPHA, device=local)
    # synthetic market: model prices at the truth, 0.1% relative noise, 0.01 floor (generator's convention)
    pricer.set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    mk = pricer.price_host(TRUTH[None, :])[0]
    mk = np.maximum(mk * (1 + 0.001 * np.random.default_rng(42).normal(size=mk.size)), 0.01)
    pricer.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)

    X_host = torch.from_numpy(sobol_sets(P, skip=rank * P)).pin_memory()
    X = X_host.to(dev)
    gathered = torch.empty((world * P, 22), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        out = pricer.normal_equations(X)
        if world > 1:  # loss + Jacobian-block all-gather over NVLink (multi-start / population drivers)
            dist.all_gather_into_tensor(gathered, out)
        return out

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak = measure_fp64_peak(local, 0.5)
    for _ in range(args.warmup):
        step()
    barrier()
    # multi-GPU correctness bit: every rank's gathered rows == the same sets priced alone on rank 0
    shard_check = None
    if world > 1:
        n_chk = 4096 // world
        ok = True
        if rank == 0:
            for rk in range(world):
                Xr = torch.from_numpy(sobol_sets(n_chk, skip=rk * P)).to(dev)
                mine = pricer.normal_equations(Xr)
                ok = ok and torch.equal(mine.view(torch.int64), gathered[rk * P: rk * P + n_chk].view(torch.int64))
        shard_check = "bitwise" if ok else "MISMATCH"
        barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    pricer.profile(True)  # CUDA-event pairs around each kernel of the step, on the launch stream
    t_wall = time.perf_counter()
    ms = timed_steps(torch, dev, step, args.steps, flush)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = launch_count() - l0
    prof = pricer.profile_read()
    pricer.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    slices_step = world * P * N_MAT * 6
    value = slices_step * args.steps / (ms * 1e-3)

    # e2e: host buffers in, host buffers out, copies inside the timed region.  N = 1: the public C-ABI host call
    # (hb_normal_eq_host: pinned staging, H2D, kernel, D2H).  N > 1: the whole multi-GPU step -- H2D of this rank's
    # candidates from pinned memory, kernel, NCCL all-gather of the result rows, D2H of the gathered block.
    Xh = X_host.numpy()
    if world == 1:
        pricer.normal_equations_host(Xh)  # warm-up at full size: the plan's pinned staging buffers grow on first use
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pricer.normal_equations_host(Xh)
        torch.cuda.synchronize(dev)
        te = time.perf_counter() - t0
        h2d, d2h = P * 5 * 8, P * 22 * 8
    else:
        gathered_host = torch.empty((world * P, 22), dtype=torch.float64).pin_memory()

        def e2e_step():
            X.copy_(X_host, non_blocking=True)
            out = pricer.normal_equations(X)
            dist.all_gather_into_tensor(gathered, out)
            gathered_host.copy_(gathered, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        te = time.perf_counter() - t0
        res = gathered_host.numpy()[rank * P: (rank + 1) * P]
        h2d, d2h = world * P * 5 * 8, world * world * P * 22 * 8
    te = max_over_ranks(te)
    e2e = slices_step * args.steps / te
    # inf losses are legitimate (moment explosion of E[S^1.75]: the reference returns inf there too)
    assert res.shape == (P, 22) and not np.isnan(res[:, 0]).any()

    # value_noskip (a): the same step on a population where no grid point can be elided
    Xs = torch.from_numpy(slow_decay_sets(min(P, 16384), skip=rank * P)).to(dev)
    for _ in range(2):
        pricer.normal_equations(Xs)
    torch.cuda.synchronize(dev)
    ms_slow = max_over_ranks(timed_steps(torch, dev, lambda: pricer.normal_equations(Xs), 2, flush) / 2)
    noskip = {"slow_decay_population": {"value": world * Xs.shape[0] * N_MAT * 6 / (ms_slow * 1e-3),
                                        "ms_per_step": ms_slow, "sets_per_gpu": int(Xs.shape[0]),
                                        "what": "sigma in [1.5, 2], rho in [-0.99, -0.9], kappa theta tiny: |phi| never "
                                                "decays on the grid, every point runs the full stage B and F"}}

    # C4 (BASELINE.json config 4): 1,048,576 candidates of ONE population sharded over the ranks, objective +
    # all-gather of the losses; strong scaling = this rate against world x the single-GPU objective rate of the
    # same run; then the whole calibration (population + batched LM) on a surface the truth can price.
    n_cand = args.c4_candidates
    c4 = None
    if n_cand > 0:
        from pde_b200.calibration.population import sobol_population

        Xc = torch.from_numpy(sobol_population(n_cand, LB, UB, seed=42)).to(dev)
        obj = lambda Z: pricer.objective(Z)  # noqa: E731
        sharded_map(obj, Xc[: 4096 * world])
        sharded_map(obj, Xc)  # warm-up at full size: the plan's routing tables grow to the batch on first use
        barrier()
        ms_c4 = max_over_ranks(timed_steps(torch, dev, lambda: sharded_map(obj, Xc), 2, flush) / 2)
        n1 = min(n_cand, 131072)
        pricer.objective(Xc[:n1])
        torch.cuda.synchronize(dev)
        ms_n1 = max_over_ranks(timed_steps(torch, dev, lambda: pricer.objective(Xc[:n1]), 2, flush) / 2)
        rate_c4, rate_1 = n_cand * N_MAT / (ms_c4 * 1e-3), n1 * N_MAT / (ms_n1 * 1e-3)
        c4 = {"candidates": n_cand, "objective_allgather_ms": ms_c4, "slices_per_s": rate_c4,
              "single_gpu_objective_slices_per_s": rate_1, "strong_scaling_efficiency": rate_c4 / (world * rate_1),
              "allgather_bytes": n_cand * 8}
        if rank == 0 or world > 1:
            import pandas as pd

            from pde_b200.calibration import PopulationCalibrator

            cal_pr = BatchPricer("fft", N_GRID, ETA, ALPHA, device=local)
            Kc, Tc, mkc = feasible_surface(cal_pr)
            df = pd.DataFrame({"strike": Kc, "maturity": Tc, "mid_price": mkc, "is_call": True})
            cal = PopulationCalibrator(mode="fft", device=local)
            cal.calibrate(df, S0, R, Q, n_candidates=4096 * world, n_starts=8, lm_iters=3)  # warm
            barrier()
            t0 = time.perf_counter()
            r_ = cal.calibrate(df, S0, R, Q, n_candidates=n_cand, n_starts=32, lm_iters=30)
            torch.cuda.synchronize(dev)
            t_cal = max_over_ranks(time.perf_counter() - t0)
            c4["calibration"] = {"seconds": t_cal, "rmse": float(r_.rmse), "n_options": int(Kc.size),
                                 "params": [float(v) for v in r_.params.to_array()], "truth": TRUTH.tolist(),
                                 "noise_rmse_expected": float(np.sqrt(np.mean((mkc * 0.001) ** 2))),
                                 "best_population_loss": r_.convergence["best_population_loss"],
                                 "surface": "32 x 50 grid minus the options whose fft price at the truth is below 0.05"}

    if rank == 0:
        per_launch_s = ms / args.steps * 1e-3
        # The step is three kernels: prefix scan (routing + live-prefix table), the direct-sum job kernel (the
        # dominant one: the sets whose integrand has decayed) and the transform job kernel (long-prefix sets).
        # Roofline of the dominant kernel: W1 flops of the sets it priced / its own duration (CUDA events around
        # the kernel on the launch stream, averaged over the timed steps).
        k_ms = {k: prof[k] / args.steps for k in ("scan_ms", "direct_ms", "transform_ms")}
        n_direct = prof["sets_direct"] if prof["sets_direct"] >= 0 else P
        dom_s = (k_ms["direct_ms"] if k_ms["direct_ms"] > 0 else ms / args.steps) * 1e-3
        achieved = (n_direct * N_MAT * 6 * FLOPS_PER_SLICE) / dom_s / 1e12  # per GPU, TFLOP/s, W1 convention
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
        hbm_achieved = (P * N_MAT * 6 * BYTES_PER_SLICE) / per_launch_s / 1e9
        cap = _ncu_capture(P)
        executed = cap["executed_flop_per_launch"] / dom_s / 1e12 if cap else None
        line = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(P),
            "base_slices_per_s": value / 6.0,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "slices/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": "hb_normal_eq_host (C ABI, host buffers)" if world == 1 else
                            "pinned H2D + hb_normal_eq + NCCL all-gather of the [P,22] rows + D2H of the gathered block"},
            "gpu_launches": int(launches),
            "wall_s_timed_region": t_wall,
            "significance_cut": {"log_cut": pricer.log_cut, "abs_price_error_budget": 2.0 ** -80},
            "value_noskip": noskip,
            "kernels_ms_per_step": dict(k_ms, step_ms=ms / args.steps,
                                        sets_direct=prof["sets_direct"], sets_transform=prof["sets_transform"],
                                        note="CUDA events around each kernel; their sum is the step (profiles/ ncu "
                                             "launch list: same shares)"),
            "roofline": {"bound": "fp64", "kernel": "direct_job_kernel<false> (live prefix + direct sums)",
                         "kernel_ms": dom_s * 1e3, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         "frac_is": "W1 algorithmic flops (SURVEY.md 8d) / measured DFMA peak: exceeds 1 because shared "
                                    "and elided work is not executed; frac_executed and fp64_pipe_active_ncu are the "
                                    "utilisation figures",
                         "traffic": cap["dram_bytes_per_launch"] if cap else None,
                         "executed_tflops": executed, "frac_executed": executed / peak if executed else None,
                         "fp64_pipe_active_ncu": cap["fp64_pipe_active_pct"] / 100.0 if cap else None,
                         "ncu_capture": ({"file": "profiles/r02_traffic.json", "kernel_source_hash": cap["kernel_source_hash"]}
                                         if cap else "no capture of this kernel source: ncu-derived fields are null"),
                         "peak_source": "DFMA probe (hb_measure_fp64_peak) run in this process; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_slice_W1": FLOPS_PER_SLICE,
                         "hbm": {"achieved_gbs": hbm_achieved, "peak_gbs": hbm_peak,
                                 "frac": hbm_achieved / hbm_peak, "bytes_per_slice": BYTES_PER_SLICE}},
        }
        if shard_check is not None:
            line["shard_check"] = shard_check
        if c4 is not None:
            line["c4"] = c4
        if world == 1 and not args.no_cpu:
            noskip["noelide_build_exact_mode"] = noelide_rate(16384)
            rate, info = cpu_reference_rate(12.0)
            line["cpu_baseline"] = {"value": rate, "unit": "slices/s", "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"]}
            try:
                from oracle.oracle import MODE_FFT, Oracle

                orc = Oracle()
                orc.use_all_cores()
                rows = fd_variants(sobol_sets(64))
                t0 = time.perf_counter()
                n_done = 0
                while time.perf_counter() - t0 < 6.0 and n_done < len(rows):
                    orc.objective_batch(MODE_FFT, rows[n_done:n_done + 32], K, T, True, mk, S0, R, Q)
                    n_done += 32
                el = time.perf_counter() - t0
                line["cpu_baseline"]["port_fft"] = {"value": n_done * N_MAT / el, "unit": "slices/s",
                                                    "cores": orc.num_threads(), "kind": "port",
                                                    "sample": f"{n_done} parameter-set evaluations x {N_MAT} maturities, "
                                                              f"C restatement of the N=4096 Carr-Madan FFT path, OpenMP over sets"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"]["port_fft"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets", type=int, default=65536, help="parameter sets per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and no-elision legs")
    ap.add_argument("--c4-candidates", type=int, default=1 << 20,
                    help="candidates of the sharded population leg (BASELINE.json config 4); 0 skips it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
