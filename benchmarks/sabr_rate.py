"""SABR path throughput on one B200: implied vols (both reference formulas) and the smile-calibration
objective (all maturities x candidates in one launch), next to the reference formulas on the host cores.

    python benchmarks/sabr_rate.py          # one JSON line
Roofline: FP64 pipe.  Algorithmic work per (candidate, strike) of the objective, counted from the formula
(sabr_calibrator.py:187-222 with the per-strike pieces hoisted): 41 plain flops + 4 DIV + 1 SQRT + 1 LOG,
with the W1 weights of SURVEY.md 8d (DIV = SQRT = 20, LOG = 50): 191 flop-equivalents.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pde_b200 import BatchSABR, measure_fp64_peak  # noqa: E402

FLOPS_PER_EVAL = 191
LB, UB = np.array([0.001, -0.99, 0.001]), np.array([2.0, 0.99, 3.0])
rng = np.random.default_rng(0)
n_smiles, n_strikes, P = 32, 50, 262144
Ts = np.linspace(0.1, 2.0, n_smiles)
Fs = 100.0 * np.exp(0.03 * Ts)
Ks = [np.linspace(80.0, 120.0, n_strikes) for _ in Ts]
eng = BatchSABR(0.5)
Vs = [eng.vols_host([[0.3, -0.3, 0.5]], k, F, T, "py")[0] * (1 + 0.005 * rng.normal(size=n_strikes)) for k, F, T in zip(Ks, Fs, Ts)]
eng.set_smiles(Ks, Vs, Fs, Ts)
X = torch.tensor(LB + (UB - LB) * rng.random((n_smiles, P, 3)), device="cuda:0")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


peak = measure_fp64_peak(0, 0.3)
ms_obj = timed(lambda: eng.objective(X))
evals = n_smiles * P * n_strikes
Xv = X[0, :65536].contiguous()
Kd = torch.tensor(Ks[0], device="cuda:0")
ms_py = timed(lambda: eng.vols(Xv, Kd, 100.0, 1.0, "py"))
ms_cpp = timed(lambda: eng.vols(Xv, Kd, 100.0, 1.0, "cpp"))
out = {
    "objective": {"smiles": n_smiles, "strikes": n_strikes, "candidates_per_smile": P, "ms": round(ms_obj, 3),
                  "candidate_smiles_per_s": round(n_smiles * P / ms_obj * 1e3), "vol_evals_per_s": round(evals / ms_obj * 1e3),
                  "w1_tflops": round(evals * FLOPS_PER_EVAL / ms_obj / 1e9, 2), "fp64_peak_tflops": round(peak, 2),
                  "frac_of_fp64_peak": round(evals * FLOPS_PER_EVAL / ms_obj / 1e9 / peak, 3)},
    "vols_py": {"sets": 65536, "strikes": n_strikes, "ms": round(ms_py, 3), "vols_per_s": round(65536 * n_strikes / ms_py * 1e3),
                "write_gbs": round(65536 * n_strikes * 8 / ms_py / 1e6, 1)},
    "vols_cpp": {"sets": 65536, "strikes": n_strikes, "ms": round(ms_cpp, 3), "vols_per_s": round(65536 * n_strikes / ms_cpp * 1e3),
                 "write_gbs": round(65536 * n_strikes * 8 / ms_cpp / 1e6, 1)},
}
# host baselines (checkers only): the reference's C++ SABRModel compiled unmodified, and our C restatement of its
# Python formula with OpenMP over candidates (the Python formula itself runs at ~2e5 evaluations/s per core)
try:
    from oracle.oracle import Reference, SabrOracle

    xs = LB + (UB - LB) * rng.random((20000, 3))
    ref, so = Reference(), SabrOracle()
    t0 = time.perf_counter()
    ref.sabr_vols(0.5, 100.0, 1.0, Ks[0], xs)
    out["cpu_reference_cpp"] = {"vols_per_s": round(xs.shape[0] * n_strikes / (time.perf_counter() - t0)), "cores": 1,
                                "kind": "reference", "sample": "20,000 sets x 50 strikes through SABRModel::implied_volatility"}
    t0 = time.perf_counter()
    so.objective(0.5, 100.0, 1.0, Ks[0], Vs[0], np.ones(n_strikes) / n_strikes, xs)
    out["cpu_port_objective"] = {"vol_evals_per_s": round(xs.shape[0] * n_strikes / (time.perf_counter() - t0)),
                                 "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                 "sample": "20,000 candidates x 50 strikes, C restatement, OpenMP over candidates"}
except Exception as e:  # pragma: no cover
    out["cpu"] = {"error": str(e)}
print(json.dumps(out))
