"""GPU: the drop-in API.  Re-expresses the reference's own acceptance tests against
pde_b200.cpp.quant_cpp / pde_b200.models / pde_b200.calibration:

  tests/python/test_cpp_bindings.py::TestHestonBindings (:25-163) and ::TestPythonWrappers (:349-375)
  tests/cpp/test_heston.cpp (:98-351)
  tests/python/calibration/test_calibration.py::TestHestonCalibrator (:95-181)

plus numeric pins the reference's tests lack: golden values from the compiled
reference (tests/golden/ref_*.npz).
"""
import math
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

S0, R, Q = 100.0, 0.05, 0.02


@pytest.fixture(scope="module")
def H():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from pde_b200.cpp import quant_cpp

    return quant_cpp.heston


@pytest.fixture(scope="module")
def model(H):
    return H.HestonModel(H.HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04))


def _close(a, b):
    return abs(a - b) <= 1e-10 * abs(b) + 1e-12


# ---- TestHestonBindings ------------------------------------------------------------------------------

def test_model_construction_and_invalid(H):
    assert H.HestonModel(H.HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04)) is not None
    with pytest.raises(ValueError):
        H.HestonModel(H.HestonParameters(-1.0, 0.04, 0.3, -0.7, 0.04))
    m = H.HestonModel(H.HestonParameters())
    m.set_parameters(H.HestonParameters(3.0, 0.05, 0.4, -0.5, 0.06))  # tests/cpp/test_heston.cpp:85-94
    assert m.parameters().kappa == 3.0
    with pytest.raises(ValueError):
        m.set_parameters(H.HestonParameters(1.0, 0.04, 0.3, 1.5, 0.04))


def test_price_call_put_and_parity(model):
    call = model.price_option(strike=100.0, maturity=1.0, spot=100.0, rate=0.05, dividend=0.02, is_call=True)
    assert isinstance(call, float) and 3.0 < call < 20.0
    put = model.price_option(strike=100.0, maturity=1.0, spot=100.0, rate=0.05, dividend=0.02, is_call=False)
    assert 0 < put < 100
    expected = S0 * math.exp(-Q) - 100.0 * math.exp(-R)
    assert abs((call - put) - expected) < 0.5  # the reference's own tolerance (:117)
    # numeric pins
    assert _close(call, 8.853337653490936) and _close(put, 5.9564127728868215)
    assert model.price_option(100.0, 1.0, 100.0, 0.05, 0.02) == call  # is_call defaults to True


def test_price_multiple_options(model, g_prices):
    strikes = [90.0, 95.0, 100.0, 105.0, 110.0]
    prices = model.price_options(strikes, [1.0], 100.0, 0.05, 0.02, True)
    assert isinstance(prices, list) and len(prices) == 5 and all(p > 0 for p in prices)
    assert all(prices[i] < prices[i - 1] for i in range(1, 5))
    got = model.price_options([90.0, 100.0, 110.0], [1.0], 100.0, 0.05, 0.02)
    for g, w in zip(got, [15.118224850104099, 8.853337653490936, 4.515002249852158]):
        assert _close(g, w)
    # per-option maturities, the full golden surface
    K, T32 = g_prices["K50"], g_prices["T32"]
    flat = model.price_options(list(np.tile(K, 32)), list(np.repeat(T32, 50)), S0, R, Q, True)
    want = g_prices["surf_default"].ravel()
    assert np.max(np.abs(np.array(flat) - want) / (1e-10 * np.abs(want) + 1e-12)) <= 1.0
    assert model.price_options([], [1.0], S0, R, Q) == []  # heston.cpp:224-226
    with pytest.raises(ValueError, match="Maturities must have size 1 or match strikes size"):
        model.price_options([90.0, 100.0, 110.0], [1.0, 2.0], S0, R, Q)


def test_invalid_arguments_raise_value_error(model):  # tests/cpp/test_heston.cpp:235-244
    with pytest.raises(ValueError, match="Strike must be positive"):
        model.price_option(-100.0, 1.0, 100.0, 0.05, 0.02)
    with pytest.raises(ValueError, match="Spot must be positive"):
        model.price_option(100.0, 1.0, -100.0, 0.05, 0.02)
    with pytest.raises(ValueError, match="Maturity must be non-negative"):
        model.price_option(100.0, -1.0, 100.0, 0.05, 0.02)


def test_zero_maturity_is_intrinsic(model, g_misc):  # tests/cpp/test_heston.cpp:216-233
    got = [model.price_option(K, 0.0, S0, R, Q, c) for K in (90.0, 110.0) for c in (True, False)]
    assert got == g_misc["t0"].tolist() == [10.0, 0.0, 0.0, 10.0]


def test_characteristic_function(model, g_cf):
    for T in (0.1, 0.5, 1.0, 2.0):  # phi(0) = 1, tests/cpp/test_heston.cpp:98-112
        assert abs(model.characteristic_function(0j, T, S0, R, Q) - 1.0) < 1e-10
    z = model.characteristic_function(1.0 + 0j, 0.0, S0, R, Q)  # :114-130
    assert abs(z - complex(math.cos(math.log(S0)), math.sin(math.log(S0)))) < 1e-10
    for u in (0.1, 0.5, 1.0, 2.0, 5.0, 10.0, 50.0):  # finiteness, :132-153
        z = model.characteristic_function(complex(u, 0.0), 1.0, S0, R, Q)
        assert math.isfinite(z.real) and math.isfinite(z.imag)
    z = model.characteristic_function(1 - 1.75j, 1.0, S0, R, Q)
    assert abs(z - complex(-108.69976080392448, -3357.5708889312546)) / abs(z) < 1e-10
    for u, want in zip(g_cf["u_gen"][:10], g_cf["cf_gen"][0][:10]):
        assert abs(model.characteristic_function(complex(u), 0.7, S0, R, Q) - want) / abs(want) < 1e-10


def test_implied_volatility_and_greeks(model, H, g_misc):
    iv = model.implied_volatility(100.0, 1.0, 100.0, 0.05, 0.02, True)
    assert 0.05 < iv < 1.0 and abs(iv - math.sqrt(0.04)) < 0.1  # :135-146
    got = np.array([[model.implied_volatility(K, T, S0, R, Q) for K in (90.0, 100.0, 110.0)] for T in (0.25, 1.0)])
    np.testing.assert_allclose(got, g_misc["iv"], rtol=1e-7)  # Newton stops at |diff| < 1e-8
    res = model.price_option_with_greeks(100.0, 1.0, 100.0, 0.05, 0.02, True)  # :148-163
    assert res.price > 0 and res.greeks_computed and 0.3 < res.greeks.delta < 0.7 and res.greeks.gamma > 0
    rows = []
    for K in (90.0, 100.0, 110.0):
        for c in (True, False):
            r = model.price_option_with_greeks(K, 1.0, S0, R, Q, c)
            rows.append([r.price, r.greeks.delta, r.greeks.gamma, r.greeks.vega, r.greeks.theta, r.greeks.rho])
    rows, want = np.array(rows), g_misc["greeks"]
    np.testing.assert_allclose(rows[:, 0], want[:, 0], rtol=1e-10)
    # finite differences of prices that agree to ~1e-13: delta/rho/theta/vega amplify by 1/bump
    np.testing.assert_allclose(rows[:, [1, 3, 4, 5]], want[:, [1, 3, 4, 5]], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(rows[:, 2], want[:, 2], rtol=1e-4, atol=1e-8)  # gamma: second difference
    assert isinstance(res.greeks, H.OptionGreeks) and "OptionGreeks(delta=" in repr(res.greeks)


# ---- TestPythonWrappers ------------------------------------------------------------------------------------

def test_python_wrapper():
    from pde_b200.models import HestonModel, HestonParameters

    m = HestonModel(kappa=2.0, theta=0.04, sigma=0.3, rho=-0.7, v0=0.04)
    p = m.price_option(strike=100, maturity=1.0, spot=100, rate=0.05, dividend=0.02)
    assert isinstance(p, float) and _close(p, 8.853337653490936)
    arr = m.price_options([90, 100, 110], 1.0, spot=100, rate=0.05, dividend=0.02)
    assert isinstance(arr, np.ndarray) and arr.shape == (3,)
    assert m.price_option(100, 1.0, 100, 0.05) > 0  # dividend defaults to 0.0
    with pytest.warns(UserWarning, match="Feller condition violated"):
        HestonModel(kappa=1.0, theta=0.02, sigma=0.5, rho=-0.7, v0=0.04)
    with pytest.raises(ValueError):
        HestonModel(kappa=-1.0)
    assert isinstance(HestonModel.from_dict(m.params.to_dict()), HestonModel)
    assert HestonModel.from_params(HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04)).params == m.params
    surf = m.implied_volatility_surface([95, 100, 105], [0.5, 1.0], 100, 0.05, 0.02)
    assert surf.shape == (3, 2) and (surf > 0.05).all()
    assert "HestonModel(κ=2.000" in repr(m)
    g = m.price_option_with_greeks(100, 1.0, 100, 0.05, 0.02)
    assert g.greeks is not None and 0.3 < g.greeks.delta < 0.7


# ---- TestHestonCalibrator ------------------------------------------------------------------------------------

def test_calibrator_price_objective_residuals_match_reference_golden(g_cal):
    from pde_b200.calibration import HestonCalibrator

    cal = HestonCalibrator()
    K, T, mkt, ic = g_cal["K"], g_cal["T"], g_cal["market"], g_cal["is_call"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, x in enumerate(g_cal["xs"]):
            pr = cal._price_options(x, K, T, ic, S0, R, Q)
            assert np.max(np.abs(pr - g_cal["prices"][i]) / (1e-10 * np.abs(g_cal["prices"][i]) + 1e-12)) <= 4.0
            obj = cal._compute_objective(x, K, T, mkt, ic, S0, R, Q)
            assert obj == pytest.approx(float(g_cal["objective"][i]), rel=1e-9)
            res = cal._compute_residuals(x, K, T, mkt, ic, S0, R, Q)
            np.testing.assert_allclose(res, g_cal["residuals"][i], rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError, match=str(g_cal["invalid_msg"])):  # models/heston.py:166 raises, no NaN
        cal._compute_objective(np.array([-1.0, 0.04, 0.3, -0.7, 0.04]), K, T, mkt, ic, S0, R, Q)
    # batched entry points: same numbers, invalid sets -> 1e10 instead of raising
    X = np.vstack([g_cal["xs"], [-1.0, 0.04, 0.3, -0.7, 0.04]])
    loss = cal.objective_batch(X, K, T, mkt, ic, S0, R, Q)
    np.testing.assert_allclose(loss[:-1], g_cal["objective"], rtol=1e-9)
    assert loss[-1] == 1e10
    single = cal._price_options(np.array([2.0, 0.04, 0.3, -0.7, 0.04]), np.array([100.0]), np.array([1.0]),
                                np.array([True]), S0, R, Q)
    assert len(single) == 1 and 2.0 < single[0] < 15.0  # test_calibration.py:162-181


def test_generate_synthetic_data_matches_reference_generator(g_cal):
    from pde_b200.calibration import HestonCalibrator

    np.random.seed(42)
    df = HestonCalibrator.generate_synthetic_data(n_strikes=7, n_maturities=3, noise_std=0.001)
    assert list(df.columns) == ["strike", "maturity", "mid_price", "option_type", "underlying", "is_call"]
    np.testing.assert_array_equal(df["strike"].values, g_cal["K"])
    np.testing.assert_array_equal(df["maturity"].values, g_cal["T"])
    np.testing.assert_allclose(df["mid_price"].values, g_cal["market"], rtol=1e-10, atol=1e-12)


def test_calibrate_end_to_end():
    """test_calibration.py:125-160 (success, rmse < 0.15, loose parameter ranges; degenerate input)."""
    from pde_b200.calibration import CalibrationResult, HestonCalibrator

    np.random.seed(42)
    df = HestonCalibrator.generate_synthetic_data(n_strikes=7, n_maturities=3, noise_std=0.001)
    cal = HestonCalibrator(global_maxiter=25, global_popsize=10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = cal.calibrate(df, S0=100.0, r=0.05, q=0.02)
    assert isinstance(res, CalibrationResult) and res.success
    assert res.rmse < 0.15
    assert 0.1 <= res.params.kappa <= 10.0 and -0.99 <= res.params.rho <= 0.99
    assert set(res.fit_quality) >= {"rmse", "r_squared", "relative_rmse", "max_abs_error", "mean_abs_error",
                                    "n_options", "feller_satisfied", "feller_value"}
    assert set(res.convergence) >= {"global_converged", "local_converged", "global_nit", "local_nfev",
                                    "calibration_time_ms"}
    tiny = df.iloc[[10]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r1 = HestonCalibrator(global_maxiter=3, global_popsize=4).calibrate(tiny, S0=100.0, r=0.05, q=0.02)
    assert r1 is not None
    fft_cal = HestonCalibrator(global_maxiter=10, global_popsize=8, mode="fft")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r2 = fft_cal.calibrate(df, S0=100.0, r=0.05, q=0.02)
    assert r2.success


def test_calibrate_uses_db_and_cached_fallback():
    """heston_calibrator.py:360-370, :700-733 with a duck-typed store."""
    from pde_b200.calibration import CalibrationError, HestonCalibrator

    class DB:
        def __init__(self):
            self.rows = []

        def store_model_parameters(self, **kw):
            self.rows.append(kw)

        def get_latest_model_parameters(self, model_type, underlying, maturity):
            return {"parameters": {"kappa": 2.0, "theta": 0.04, "sigma": 0.3, "rho": -0.7, "v0": 0.04},
                    "fit_quality": {"rmse": 0.01}, "converged": True, "time": "t0"}

    np.random.seed(1)
    df = HestonCalibrator.generate_synthetic_data(n_strikes=5, n_maturities=2, noise_std=0.001)
    db = DB()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        HestonCalibrator(db=db, global_maxiter=3, global_popsize=4).calibrate(df, 100.0, 0.05, 0.02)
    assert db.rows and db.rows[0]["model_type"] == "heston" and db.rows[0]["underlying"] == "SYNTHETIC"
    # local_method='lm' cannot take bounds -> SciPy raises -> cached parameters are returned
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = HestonCalibrator(db=db, global_maxiter=2, global_popsize=4, local_method="lm").calibrate(df, 100.0, 0.05, 0.02)
    assert res.convergence == {"cached": True} and res.warnings == ["Using cached parameters"]
    with pytest.raises(CalibrationError):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            HestonCalibrator(global_maxiter=2, global_popsize=4, local_method="lm").calibrate(df, 100.0, 0.05, 0.02)
    with pytest.raises(ValueError, match="Missing required column"):
        HestonCalibrator().calibrate(df.drop(columns=["mid_price"]), 100.0, 0.05, 0.02)
