"""Generate tests/golden/*.npz from the reference itself.  Run in the BUILD container only
(needs /root/reference and `make -C oracle ref pyref`); the fixtures are committed because
/root/reference does not exist on the GPU box.

    python tests/golden/make_golden.py

Sources of truth
  ref_cf.npz, ref_prices.npz, ref_misc.npz : oracle/_ref/libheston_ref.so = the reference's
      src/cpp/models/heston.cpp compiled unmodified (g++ 13.3 -O3 -std=c++17 -fopenmp, x86-64,
      no -march=native).
  ref_calibrator.npz : the reference's Python `HestonCalibrator` (imported from
      /root/reference/src/python with its root __init__ bypassed) driving the
      reference's own pybind11 module oracle/_ref/quant_cpp*.so, plus SciPy's
      approx_derivative exactly as least_squares(jac='2-point', bounds=...) calls it.
  fft_selfcheck.npz : our OWN oracle (liborc.so) in FFT mode -- NOT a reference output
      (the reference has no FFT pricer; parity unpinned).  Kept only so that the restatement
      cannot drift silently; also holds an independent numpy.fft recomputation.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import MODE_FFT, Oracle, Reference, build  # noqa: E402

REF_PY = "/root/reference/src/python"
S0, R_, Q_ = 100.0, 0.05, 0.02
DEFAULT = np.array([2.0, 0.04, 0.3, -0.7, 0.04])
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])  # heston_calibrator.py:201-207
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])


def param_sets(n_sobol=16):
    from scipy.stats import qmc

    sob = qmc.Sobol(d=5, seed=42).random(n_sobol)
    sets = [DEFAULT] + list(LB + (UB - LB) * sob)
    for mask in range(32):  # the 32 box corners
        sets.append(np.where([(mask >> b) & 1 for b in range(5)], UB, LB))
    return np.array(sets)


def import_reference_calibrator():
    """quant_trading.calibration without running quant_trading/__init__.py."""
    import importlib
    import importlib.util

    pkg = types.ModuleType("quant_trading")
    pkg.__path__ = [os.path.join(REF_PY, "quant_trading")]
    sys.modules["quant_trading"] = pkg
    cpp = types.ModuleType("quant_trading.cpp")
    cpp.__path__ = []
    sys.modules["quant_trading.cpp"] = cpp
    ext = [f for f in os.listdir(os.path.join(ROOT, "oracle", "_ref")) if f.startswith("quant_cpp")][0]
    spec = importlib.util.spec_from_file_location("quant_cpp", os.path.join(ROOT, "oracle", "_ref", ext))
    quant_cpp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(quant_cpp)
    sys.modules["quant_trading.cpp.quant_cpp"] = quant_cpp
    cpp.quant_cpp = quant_cpp
    mod = importlib.import_module("quant_trading.calibration.heston_calibrator")
    return mod, quant_cpp


def main():
    build(ref=True, pyref=True)
    ref, orc = Reference(), Oracle()
    sets = param_sets()

    # ---- CF on (a subset of) the damped grids -------------------------------------------
    j_fft = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, 4096, 29), [4095]]))
    v_fft = 0.25 * j_fft
    j_rg = np.unique(np.concatenate([np.arange(1, 32), np.arange(32, 1024, 13), [1023]]))
    v_rg = 0.01 * j_rg
    Ts = np.array([0.1, 0.37, 1.0, 2.0])
    cf_fft = np.empty((len(sets), len(Ts), len(v_fft)), dtype=np.complex128)
    cf_rg = np.empty((len(sets), len(Ts), len(v_rg)), dtype=np.complex128)
    for i, p in enumerate(sets):
        for m, T in enumerate(Ts):
            cf_fft[i, m] = ref.cf_grid(p, v_fft, -1.75, T, S0, R_, Q_)
            cf_rg[i, m] = ref.cf_grid(p, v_rg, -1.75, T, S0, R_, Q_)
    # general complex u (API: characteristic_function(u: complex, ...))
    rng = np.random.default_rng(7)
    u_gen = rng.normal(size=40) * 5 + 1j * rng.uniform(-2.0, 0.5, size=40)
    cf_gen = np.array([[ref.cf(p, u, 0.7, S0, R_, Q_) for u in u_gen] for p in sets[:8]])
    cf_T0 = np.array([ref.cf(DEFAULT, u, 0.0, S0, R_, Q_) for u in u_gen])
    np.savez_compressed(os.path.join(HERE, "ref_cf.npz"), params=sets, T=Ts, v_fft=v_fft, j_fft=j_fft, v_rg=v_rg,
                        j_rg=j_rg, cf_fft=cf_fft, cf_rg=cf_rg, u_gen=u_gen, cf_gen=cf_gen, cf_T0=cf_T0,
                        S0=S0, r=R_, q=Q_, ui=-1.75)

    # ---- refgrid prices ------------------------------------------------------------------
    K50 = np.linspace(80.0, 120.0, 50)
    T32 = np.linspace(0.1, 1.0, 32)
    T8 = T32[::4]
    surf_default = np.array([ref.price_options(DEFAULT, K50, T, S0, R_, Q_) for T in T32])  # [32][50]
    puts_default = np.array([ref.price_options(DEFAULT, K50, T, S0, R_, Q_, False) for T in T8])
    surf_sets = np.array([[ref.price_options(p, K50, T, S0, R_, Q_) for T in T8] for p in sets])  # [S][8][50]
    K_wide = np.array([60.0, 80.0, 100.0, 110.0, 120.0, 150.0, 200.0])
    wide = np.array([[ref.price_option(DEFAULT, K, T, S0, R_, Q_) for K in K_wide] for T in (0.1, 1.0)])
    np.savez_compressed(os.path.join(HERE, "ref_prices.npz"), params=sets, K50=K50, T32=T32, T8=T8,
                        surf_default=surf_default, puts_default=puts_default, surf_sets=surf_sets, K_wide=K_wide,
                        wide=wide, S0=S0, r=R_, q=Q_)

    # ---- misc: T=0, IV, Greeks, error strings ---------------------------------------------
    iv = np.array([[ref.implied_vol(DEFAULT, K, T, S0, R_, Q_) for K in (90.0, 100.0, 110.0)] for T in (0.25, 1.0)])
    gk = np.array([ref.greeks(DEFAULT, K, 1.0, S0, R_, Q_, c) for K in (90.0, 100.0, 110.0) for c in (True, False)])
    t0 = np.array([ref.price_option(DEFAULT, K, 0.0, S0, R_, Q_, c) for K in (90.0, 110.0) for c in (True, False)])
    errs = {}
    for name, bad in [("kappa", [-1, .04, .3, -.7, .04]), ("theta", [2, 0, .3, -.7, .04]),
                      ("sigma", [2, .04, -0.5, -.7, .04]), ("rho", [2, .04, .3, 1.0, .04]),
                      ("v0", [2, .04, .3, -.7, -0.01])]:
        try:
            ref.validate(bad)
        except ValueError as e:
            errs[name] = str(e)
    for name, args in [("strike", (-100.0, 1.0, 100.0)), ("spot", (100.0, 1.0, -100.0)),
                       ("maturity", (100.0, -1.0, 100.0))]:
        try:
            ref.price_option(DEFAULT, args[0], args[1], args[2], R_, Q_)
        except ValueError as e:
            errs[name] = str(e)
    np.savez_compressed(os.path.join(HERE, "ref_misc.npz"), iv=iv, greeks=gk, t0=t0,
                        err_keys=np.array(list(errs.keys())), err_msgs=np.array(list(errs.values())))

    # ---- the reference's Python calibrator ---------------------------------------------------
    mod, quant_cpp = import_reference_calibrator()
    from scipy.optimize._numdiff import approx_derivative

    np.random.seed(42)  # tests/python/conftest.py:12
    df = mod.HestonCalibrator.generate_synthetic_data(n_strikes=7, n_maturities=3, noise_std=0.001)
    np.random.seed(42)
    df_big = mod.HestonCalibrator.generate_synthetic_data(n_strikes=50, n_maturities=32, noise_std=0.001)
    cal = mod.HestonCalibrator()
    K = df["strike"].values.astype(float)
    T = df["maturity"].values.astype(float)
    mkt = df["mid_price"].values.astype(float)
    ic = df["is_call"].values
    import warnings

    xs = np.array([DEFAULT, [1.5, 0.05, 0.4, -0.5, 0.05], [6.3313, 0.034994, 0.36308, -0.98999, 0.054031],
                   [0.1, 0.01, 2.0, -0.99, 0.01], [10.0, 1.0, 0.01, 0.99, 1.0], [3.0, 0.09, 0.2, 0.3, 0.02],
                   [2.0, 0.3, 0.5, -0.5, 0.3], [1.0, 0.2, 1.0, 0.0, 0.25], [4.0, 0.15, 0.8, -0.9, 0.5]])
    obj, res, jac, prices = [], [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for x in xs:
            prices.append(cal._price_options(x, K, T, ic, S0, R_, Q_))
            obj.append(cal._compute_objective(x, K, T, mkt, ic, S0, R_, Q_))
            f0 = cal._compute_residuals(x, K, T, mkt, ic, S0, R_, Q_)
            res.append(f0)
            # least_squares: approx_derivative(fun, x, rel_step=None, method='2-point', f0=f, bounds=bounds)
            jac.append(approx_derivative(lambda y: cal._compute_residuals(y, K, T, mkt, ic, S0, R_, Q_), x,
                                         rel_step=None, method="2-point", f0=f0, bounds=(LB, UB)))
        # mixed calls/puts and an invalid parameter set (-> NaN prices -> 1e10)
        ic_mixed = (np.arange(len(K)) % 2 == 0)
        prices_mixed = cal._price_options(DEFAULT, K, T, ic_mixed, S0, R_, Q_)
        # an invalid parameter set RAISES from the wrapper ctor (models/heston.py:166) before the
        # per-option try/except (heston_calibrator.py:563-584) is reached
        try:
            cal._compute_objective(np.array([-1.0, 0.04, 0.3, -0.7, 0.04]), K, T, mkt, ic, S0, R_, Q_)
            invalid_msg = ""
        except ValueError as e:
            invalid_msg = str(e)
    np.savez_compressed(os.path.join(HERE, "ref_calibrator.npz"), K=K, T=T, market=mkt, is_call=ic, xs=xs,
                        prices=np.array(prices), objective=np.array(obj), residuals=np.array(res),
                        jacobian=np.array(jac), lb=LB, ub=UB, is_call_mixed=ic_mixed, prices_mixed=prices_mixed,
                        invalid_msg=np.array(invalid_msg), big_K=df_big["strike"].values.astype(float),
                        big_T=df_big["maturity"].values.astype(float),
                        big_market=df_big["mid_price"].values.astype(float), S0=S0, r=R_, q=Q_)

    # ---- FFT self-check (our oracle; NOT a reference output) -----------------------------------
    N, eta, alpha = 4096, 0.25, 0.75
    fft_prices = orc.price_batch(MODE_FFT, sets, np.tile(K50, len(T8)), np.repeat(T8, len(K50)), True, S0, R_, Q_)
    # independent numpy.fft recomputation of the same spec from reference CF values
    lam, b = 2 * np.pi / (N * eta), np.pi / eta
    j = np.arange(N)
    v = eta * j
    w = (eta / 3.0) * np.where(j == 0, 1.0, np.where(j % 2 == 1, 4.0, 2.0))
    np_prices = np.empty((4, len(T8), len(K50)))
    for i, p in enumerate(sets[:4]):
        for m, Tm in enumerate(T8):
            phi = ref.cf_grid(p, v, -(alpha + 1), Tm, S0, R_, Q_)
            psi = np.exp(-R_ * Tm) * phi / (alpha * alpha + alpha - v * v + 1j * (2 * alpha + 1) * v)
            X = np.fft.fft(psi * w * np.where(j % 2 == 1, -1.0, 1.0))
            km = -b + lam * j
            Cg = np.exp(-alpha * km) / np.pi * X.real
            k = np.log(K50)
            mm = np.floor((k + b) / lam).astype(int)
            np_prices[i, m] = np.maximum(Cg[mm] + (Cg[mm + 1] - Cg[mm]) * (k - km[mm]) / lam, 0.0)
    np.savez_compressed(os.path.join(HERE, "fft_selfcheck.npz"), params=sets, K50=K50, T8=T8, N=N, eta=eta,
                        alpha=alpha, fft_prices=fft_prices.reshape(len(sets), len(T8), len(K50)),
                        numpy_fft_prices=np_prices, S0=S0, r=R_, q=Q_)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
