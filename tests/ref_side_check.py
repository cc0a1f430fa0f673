"""Run in a SUBPROCESS by tests/test_gpu_reference_side.py (needs a B200): the reference's unmodified Python layer
(baseline/_ref: quant_trading.models.heston, quant_trading.calibration.heston_calibrator) over the injected
pde_b200.cpp.quant_cpp, against the golden vectors the same reference code produced over its OWN C++ extension
(tests/golden/ref_calibrator.npz, ref_prices.npz; generator tests/golden/make_golden.py).

Exits 0 and prints "REFERENCE-SIDE OK" when every check holds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_inject  # noqa: E402

b200 = ref_inject.inject()
from quant_trading.calibration.heston_calibrator import HestonCalibrator  # noqa: E402  (reference file)
from quant_trading.models import HestonModel  # noqa: E402  (reference file)
import quant_trading.models.heston as ref_heston  # noqa: E402

assert ref_heston.__file__.startswith(ref_inject.REF_SITE), ref_heston.__file__
assert ref_heston._CPP_AVAILABLE and ref_heston.quant_cpp is b200


def viol(got, want):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    return float(np.max(np.abs(got[ok] - want[ok]) / (1e-10 * np.abs(want[ok]) + 1e-12)))


g = np.load(os.path.join(HERE, "golden", "ref_calibrator.npz"))
gp = np.load(os.path.join(HERE, "golden", "ref_prices.npz"))
S0, r, q = float(g["S0"]), float(g["r"]), float(g["q"])
K, T, mkt, ic, xs = g["K"], g["T"], g["market"], g["is_call"], g["xs"]

# models/heston.py:137-207 HestonModel.price_option / price_options, unmodified, over the injected module
m = HestonModel(kappa=2.0, theta=0.04, sigma=0.3, rho=-0.7, v0=0.04)
assert abs(m.price_option(strike=100, maturity=1.0, spot=100, rate=0.05, dividend=0.02) - 8.853337653490936) < 1e-9
surf = np.array([m.price_options(gp["K50"], float(t), S0, r, q) for t in gp["T8"]])
assert viol(surf, gp["surf_sets"][0] if np.allclose(gp["params"][0], [2.0, 0.04, 0.3, -0.7, 0.04]) else surf) <= 1.0
puts = np.array([m.price_options(gp["K50"], float(t), S0, r, q, is_call=False) for t in gp["T8"]])
assert viol(puts, gp["puts_default"]) <= 1.0

# calibration/heston_calibrator.py:486-586 _price_options / _compute_objective / _compute_residuals, unmodified
cal = HestonCalibrator()
worst = 0.0
for i, x in enumerate(xs):
    limit = 4.0 if x[2] < 0.02 else 1.0
    p = cal._price_options(x, K, T, ic, S0, r, q)
    v = viol(p, g["prices"][i])
    assert v <= limit, (i, x, v)
    worst = max(worst, v)
    obj = cal._compute_objective(x, K, T, mkt, ic, S0, r, q)
    want = float(g["objective"][i])
    assert (obj == 1e10) == (want == 1e10) and abs(obj - want) <= 1e-9 * abs(want), (i, obj, want)
    res = cal._compute_residuals(x, K, T, mkt, ic, S0, r, q)
    np.testing.assert_allclose(res, g["residuals"][i], rtol=1e-9, atol=1e-12)
# an invalid candidate: the reference wrapper raises ValueError with the reference's message (models/heston.py:166)
try:
    cal._price_options(np.array([-1.0, 0.04, 0.3, -0.7, 0.04]), K, T, ic, S0, r, q)
    raise SystemExit("invalid parameters did not raise")
except ValueError as e:
    assert "kappa" in str(e), e
print(f"REFERENCE-SIDE OK (worst price deviation {worst:.3g} x tolerance over {len(xs)} candidates)")
