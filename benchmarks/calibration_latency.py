"""End-to-end calibration latency on the reference's own test setup (7 strikes x 3 maturities,
noise 0.001, seed 42; tests/python/calibration/test_calibration.py:103-143) and on a 50 x 32 surface.
Reference (one CPU core): 76.5 s, 7,575 objective + 667 residual evaluations, rmse 0.072.

    python benchmarks/calibration_latency.py        # needs a B200
"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pde_b200.calibration import HestonCalibrator, PopulationCalibrator  # noqa: E402

warnings.simplefilter("ignore")
out = {}
for name, (ns, nm) in {"7x3": (7, 3), "50x32": (50, 32)}.items():
    for mode in ("refgrid", "fft"):
        np.random.seed(42)
        df = HestonCalibrator.generate_synthetic_data(n_strikes=ns, n_maturities=nm, noise_std=0.001, mode=mode)
        cal = HestonCalibrator(mode=mode)  # reference defaults: maxiter=100, popsize=15
        t0 = time.time()
        res = cal.calibrate(df, S0=100.0, r=0.05, q=0.02)
        dt = time.time() - t0
        out[f"HestonCalibrator[{mode}] {name}"] = {"seconds": round(dt, 3), "rmse": round(res.rmse, 5),
                                                   "global_nit": int(res.convergence["global_nit"]),
                                                   "local_nfev": int(res.convergence["local_nfev"]),
                                                   "params": [round(float(x), 5) for x in res.params.to_array()]}
        pop = PopulationCalibrator(mode=mode)
        t0 = time.time()
        res = pop.calibrate(df, S0=100.0, r=0.05, q=0.02, n_candidates=65536, n_starts=32, lm_iters=30)
        dt = time.time() - t0
        out[f"PopulationCalibrator[{mode}] {name} (65,536 candidates, 32 starts)"] = {
            "seconds": round(dt, 3), "rmse": round(res.rmse, 5), "lm_evals": int(res.convergence["local_nfev"]),
            "params": [round(float(x), 5) for x in res.params.to_array()]}
print(json.dumps(out, indent=1))
