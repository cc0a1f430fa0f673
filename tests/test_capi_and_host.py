"""CPU: the C-ABI library loads and exports every symbol include/heston_b200.h declares, fails
loudly without a GPU, and the host-side mirrors of the reference's Python classes behave like
the reference's (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "heston_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pde_b200 import _lib

    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in pde_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.hb_version() == 100
    assert isinstance(lib.hb_last_error(), bytes)


def test_library_is_sm100a_only_and_in_tree():
    from pde_b200 import _lib

    assert os.path.dirname(_lib.LIB_PATH).endswith(os.path.join("pde_b200", "csrc"))
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pde_b200 import BatchPricer
    from pde_b200._lib import HestonB200Error
    from pde_b200.cpp import quant_cpp

    with pytest.raises(HestonB200Error, match="no CUDA device"):
        BatchPricer("fft")
    m = quant_cpp.heston.HestonModel(quant_cpp.heston.HestonParameters())
    with pytest.raises(HestonB200Error):
        m.price_option(100.0, 1.0, 100.0, 0.05, 0.02, True)
    # the product never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pde_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liborc" not in src, f


def test_parameter_validation_messages_match_reference(g_misc):
    """HestonParameters::validate strings (heston.hpp:81-100) through hb_model_validate."""
    from pde_b200.cpp import quant_cpp

    msgs = dict(zip(g_misc["err_keys"].tolist(), g_misc["err_msgs"].tolist()))
    H = quant_cpp.heston
    for key, bad in [("kappa", (-1, .04, .3, -.7, .04)), ("theta", (2, 0, .3, -.7, .04)),
                     ("sigma", (2, .04, -0.5, -.7, .04)), ("rho", (2, .04, .3, 1.0, .04)),
                     ("v0", (2, .04, .3, -.7, -0.01))]:
        with pytest.raises(ValueError) as e:
            H.HestonParameters(*bad).validate()
        assert str(e.value) == msgs[key]
        with pytest.raises(ValueError):
            H.HestonModel(H.HestonParameters(*bad))  # tests/python/test_cpp_bindings.py:62-66
    p = H.HestonParameters()
    assert (p.kappa, p.theta, p.sigma, p.rho, p.v0) == (2.0, 0.04, 0.3, -0.7, 0.04)  # :28-35
    p = H.HestonParameters(3.0, 0.05, 0.4, -0.5, 0.06)
    assert (p.kappa, p.theta, p.sigma, p.rho, p.v0) == (3.0, 0.05, 0.4, -0.5, 0.06)  # :37-45
    assert H.HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04).is_feller_satisfied()  # :47-56
    assert not H.HestonParameters(1.0, 0.02, 0.5, -0.7, 0.04).is_feller_satisfied()
    assert repr(H.HestonParameters()) == ("HestonParameters(kappa=2.000000, theta=0.040000, sigma=0.300000, "
                                          "rho=-0.700000, v0=0.040000, feller=OK)")
    assert quant_cpp.__version__ == "0.1.0"
    for name in ("OptionGreeks", "PricingResult", "HestonParameters", "HestonModel"):
        assert hasattr(quant_cpp.heston, name)


def test_calibrator_host_classes():
    from pde_b200.calibration import CalibrationResult, HestonCalibrator, HestonParameters

    with pytest.raises(ValueError, match="kappa must be positive"):
        HestonParameters(-1, 0.04, 0.3, -0.7, 0.04)
    with pytest.raises(ValueError, match="rho must be in"):
        HestonParameters(1, 0.04, 0.3, 1.0, 0.04)
    p = HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04)
    assert p.is_feller_satisfied and p.feller_condition_value == pytest.approx(0.07)
    assert np.array_equal(p.to_array(), [2.0, 0.04, 0.3, -0.7, 0.04])
    assert HestonParameters.from_array(p.to_array()) == p
    assert HestonParameters.from_dict(p.to_dict()) == p
    assert p.to_dict()["feller_satisfied"] is True
    from datetime import datetime

    r = CalibrationResult(p, {"rmse": 0.1}, {"local_converged": True}, datetime.now())
    assert r.success and r.rmse == 0.1 and r.to_dict()["success"]
    assert not CalibrationResult(p, {}, {}, datetime.now()).success
    assert CalibrationResult(p, {}, {"cached": True}, datetime.now()).success
    assert HestonCalibrator.DEFAULT_BOUNDS == {"kappa": (0.1, 10.0), "theta": (0.01, 1.0), "sigma": (0.01, 2.0),
                                               "rho": (-0.99, 0.99), "v0": (0.01, 1.0)}
    cal = HestonCalibrator()
    import pandas as pd

    with pytest.raises(ValueError, match="Missing required column"):
        cal._validate_market_data(pd.DataFrame({"strike": [1.0], "maturity": [1.0]}))
    with pytest.raises(ValueError, match="price <= 0"):
        cal._validate_market_data(pd.DataFrame({"strike": [1.0], "maturity": [1.0], "mid_price": [0.0]}))
    with pytest.raises(ValueError, match="maturity <= 0"):
        cal._validate_market_data(pd.DataFrame({"strike": [1.0], "maturity": [0.0], "mid_price": [1.0]}))
    w = cal._validate_parameters(HestonParameters(9.0, 0.01, 1.8, -0.97, 0.6))
    assert len(w) == 5 and "Feller" in w[0]


def test_models_wrapper_validation_without_gpu():
    from pde_b200.models import HestonParameters

    with pytest.raises(ValueError, match="kappa must be positive, got -1"):
        HestonParameters(-1, 0.04, 0.3, -0.7, 0.04).validate()
    p = HestonParameters(2.0, 0.04, 0.3, -0.7, 0.04)
    assert p.is_valid() and p.is_feller_satisfied() and p.to_dict()["v0"] == 0.04


def test_population_helpers_cpu():
    import torch

    from pde_b200.calibration.population import _unpack_normal_equations, sobol_population

    lb, ub = np.array([0.1, 0.01, 0.01, -0.99, 0.01]), np.array([10.0, 1.0, 2.0, 0.99, 1.0])
    X = sobol_population(256, lb, ub, seed=42)
    assert X.shape == (256, 5) and (X >= lb).all() and (X <= ub).all()
    assert np.array_equal(sobol_population(128, lb, ub, seed=42, skip=128), X[128:])  # shardable by fast-forward
    rng = np.random.default_rng(0)
    J, r = rng.normal(size=(3, 40, 5)), rng.normal(size=(3, 40))
    iu = np.triu_indices(5)
    neq = np.zeros((3, 22))
    for i in range(3):
        neq[i, 1], neq[i, 2:7], neq[i, 7:] = r[i] @ r[i], J[i].T @ r[i], (J[i].T @ J[i])[iu]
    rr, g, A = _unpack_normal_equations(torch.tensor(neq))
    for i in range(3):
        np.testing.assert_allclose(A[i].numpy(), J[i].T @ J[i], rtol=1e-13)
        np.testing.assert_allclose(g[i].numpy(), J[i].T @ r[i], rtol=1e-13)
