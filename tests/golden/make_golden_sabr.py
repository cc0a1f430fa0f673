"""Generate tests/golden/ref_sabr.npz from the reference itself (build container only: needs /root/reference
and `make -C oracle ref`).      python tests/golden/make_golden_sabr.py

Sources of truth
  vols_cpp : oracle/_ref/libheston_ref.so = the reference's src/cpp/models/sabr.cpp compiled unmodified
             (SABRModel::implied_volatility, sabr.cpp:130-192; NaN where it throws).
  vols_py, smile_*, cal_* : the reference's Python SABRCalibrator (sabr_calibrator.py) imported from
             /root/reference/src/python with the package root __init__ bypassed:
             sabr_implied_vol on a grid, generate_synthetic_smile with the reference test's arguments
             (tests/python/calibration/test_calibration.py:224-236, seed 42 as tests/python/conftest.py:12),
             calibrate_single_maturity / calibrate results on those smiles.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import import_reference_calibrator  # noqa: E402
from oracle.oracle import Reference, build  # noqa: E402

LB = np.array([0.001, -0.99, 0.001])  # sabr_calibrator.py:130-134
UB = np.array([2.0, 0.99, 3.0])


def main():
    build(ref=True, pyref=True)
    ref = Reference()
    import_reference_calibrator()
    mod = importlib.import_module("quant_trading.calibration.sabr_calibrator")
    from scipy.stats import qmc

    F = 100.0
    corners = np.array([np.where([(m >> b) & 1 for b in range(3)], UB, LB) for m in range(8)])
    params = np.vstack([[0.3, -0.3, 0.5], LB + (UB - LB) * qmc.Sobol(d=3, seed=42).random(32), corners,
                        [0.3, -0.3, 0.0], [0.3, 0.0, 1e-12], [1e-12, 0.5, 0.5]])
    K = np.array([50.0, 80.0, 90.0, 95.0, 99.0, 100.0 - 1e-11, 100.0, 100.0 + 1e-9, 100.00005, 101.0, 105.0, 110.0,
                  125.0, 200.0])
    betas = np.array([0.0, 0.5, 1.0])
    Ts = np.array([0.0, 1e-12, 0.25, 1.0, 2.0])
    vols_cpp = np.array([[ref.sabr_vols(b, F, T, K, params) for T in Ts] for b in betas])  # [beta][T][P][n]
    with np.errstate(all="ignore"):
        vols_py = np.array([[[[mod.SABRCalibrator(beta=b).sabr_implied_vol(F, k, T, a, b, r, n) for k in K]
                              for a, r, n in params] for T in Ts] for b in betas])

    # the reference test's smile and its calibration
    np.random.seed(42)
    smile = mod.SABRCalibrator.generate_synthetic_smile(F=100.0, T=0.25, alpha=0.3, beta=0.5, rho=-0.3, nu=0.5,
                                                         n_strikes=11, noise_std=0.001)
    cal = mod.SABRCalibrator(beta=0.5)
    p1, rmse1 = cal.calibrate_single_maturity(smile["strike"].values, smile["implied_vol"].values, F=100.0, T=0.25)
    import pandas as pd

    np.random.seed(42)
    multi = pd.concat([mod.SABRCalibrator.generate_synthetic_smile(F=100.0, T=T, noise_std=0.001)
                       for T in (0.25, 0.5, 1.0)], ignore_index=True)
    res = cal.calibrate(market_options=multi, F0=100.0)
    multi_params = np.array([[res.params_by_maturity[T].alpha, res.params_by_maturity[T].rho,
                              res.params_by_maturity[T].nu] for T in (0.25, 0.5, 1.0)])
    multi_rmse = np.array([res.rmse_by_maturity[T] for T in (0.25, 0.5, 1.0)])
    np.savez_compressed(os.path.join(HERE, "ref_sabr.npz"), F=F, params=params, K=K, betas=betas, Ts=Ts,
                        vols_cpp=vols_cpp, vols_py=vols_py,
                        smile_K=smile["strike"].values, smile_vol=smile["implied_vol"].values,
                        cal_single=np.array([p1.alpha, p1.rho, p1.nu, rmse1]),
                        multi_K=multi["strike"].values, multi_T=multi["T"].values, multi_vol=multi["implied_vol"].values,
                        multi_params=multi_params, multi_rmse=multi_rmse, multi_total_rmse=res.total_rmse)
    print("ref_sabr.npz:", vols_cpp.shape, "single:", p1, rmse1, "multi total rmse:", res.total_rmse)


if __name__ == "__main__":
    main()
