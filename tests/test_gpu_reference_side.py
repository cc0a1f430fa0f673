"""The drop-in boundary exercised from the REFERENCE's side (VERDICT r01 "missing #1"):

* the reference's own, unmodified Python layer (baseline/_ref, installed by baseline/install_ref.py with
  `pip install --target`) runs over the injected ``pde_b200.cpp.quant_cpp`` and
  reproduces the golden vectors it produced over its own C++ extension;
* the reference's own acceptance tests for the binding boundary (tests/python/test_cpp_bindings.py:25-163
  ``TestHestonBindings``, ``TestSABRBindings``, the Heston/SABR ``TestPythonWrappers``; tests/python/calibration/
  test_calibration.py ``TestHestonParameters`` / ``TestHestonCalibrator``) run unmodified against it;
* the pybind11 binding INTEGRATION.md section 2 proposes is compiled (tests/ext) and agrees bit for bit with the
  ctypes host path.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_inject  # noqa: E402

S0, R, Q = 100.0, 0.05, 0.02


def _run(args, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def _need_reference():
    if not ref_inject.available():
        pytest.skip("baseline/_ref missing: run `python baseline/install_ref.py` in the build container")


@pytest.mark.gpu
def test_reference_python_layer_runs_unmodified_over_injected_module():
    _need_reference()
    r = _run([os.path.join(HERE, "ref_side_check.py")])
    assert r.returncode == 0 and "REFERENCE-SIDE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_reference_own_binding_tests_pass_against_injected_module():
    """tests/python/test_cpp_bindings.py of the reference, unmodified: TestHestonBindings (:25-163), TestSABRBindings
    and the Heston / SABR wrapper tests.  (OU and PDE-solver classes are outside the tier: SURVEY.md section 2.)"""
    _need_reference()
    f = os.path.join(ref_inject.REF_TESTS, "test_cpp_bindings.py")
    r = _run([os.path.join(HERE, "ref_inject.py"), f, "-q", "-p", "no:cacheprovider", "-k",
              "TestHestonBindings or TestSABRBindings or test_heston_wrapper or test_sabr_wrapper"])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_reference_own_calibrator_tests_pass_against_injected_module():
    """tests/python/calibration/test_calibration.py of the reference, unmodified: TestHestonParameters and
    TestHestonCalibrator (:95-181) -- a whole DE + TRF calibration driven by the reference's own SciPy closures,
    one price_option call at a time through the injected module."""
    _need_reference()
    f = os.path.join(ref_inject.REF_TESTS, "test_calibration.py")
    r = _run([os.path.join(HERE, "ref_inject.py"), f, "-q", "-p", "no:cacheprovider", "-k",
              "TestHestonParameters or TestHestonCalibrator"], timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


def _ext():
    sys.path.insert(0, os.path.join(HERE, "ext"))
    try:
        import quant_cpp_b200
    except ImportError:
        import importlib.util

        spec = importlib.util.spec_from_file_location("build_ext", os.path.join(HERE, "ext", "build_ext.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
        import quant_cpp_b200
    return quant_cpp_b200


def test_integration_binding_compiles_and_fails_loudly_without_a_device():
    """INTEGRATION.md section 2, compiled: the class exists with the documented methods; without a CUDA device the
    constructor raises the library's error (no CPU fallback)."""
    q = _ext()
    for name in ("set_surface", "objective", "prices", "normal_equations", "implied_vols", "greeks"):
        assert hasattr(q.heston.B200Plan, name)
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            q.heston.B200Plan("fft")


@pytest.mark.gpu
def test_integration_binding_matches_ctypes_host_path():
    from pde_b200 import BatchPricer

    q = _ext()
    K = np.tile(np.linspace(85, 115, 11), 5)
    T = np.repeat(np.linspace(0.2, 1.2, 5), 11)
    ic = (np.arange(K.size) % 4 != 0)
    X = np.array([[2.0, 0.04, 0.3, -0.7, 0.04], [1.5, 0.09, 0.5, -0.3, 0.06], [4.0, 0.15, 0.8, -0.9, 0.5],
                  [-1.0, 0.04, 0.3, -0.7, 0.04]])
    for mode in ("fft", "refgrid"):
        ours = BatchPricer(mode).set_surface(K, T, ic, None, S0=S0, r=R, q=Q)
        mk = np.maximum(ours.price_host(X[:1])[0], 0.01) * 1.01
        ours.set_surface(K, T, ic, mk, S0=S0, r=R, q=Q)
        plan = q.heston.B200Plan(mode)
        plan.set_surface(K, T, ic.astype(np.uint8), mk, S0, R, Q)
        assert np.array_equal(plan.prices(X), ours.price_host(X), equal_nan=True)
        assert np.array_equal(plan.objective(X), ours.objective_host(X))
        assert np.array_equal(plan.normal_equations(X), ours.normal_equations_host(X), equal_nan=True)
        assert np.array_equal(plan.implied_vols(X[:3]), ours.implied_vol_host(X[:3]), equal_nan=True)
        assert plan.greeks(X[:2]).shape == (2, K.size, 5)
        # the calibrator-side closures of INTEGRATION.md section 2
        loss = plan.objective(np.ascontiguousarray(X[:3]))
        p = plan.prices(X[None, 0])[0]
        want = 1e10 if (np.isnan(p).any() or (p <= 0).any()) else np.sum(((p - mk) / mk) ** 2)  # :507-511
        assert loss[0] == pytest.approx(want, rel=1e-12)
        res = (np.maximum(p, 1e-10) - mk) / mk  # the residual closure of INTEGRATION.md section 2
        assert plan.normal_equations(X[None, 0])[0, 1] == pytest.approx(res @ res, rel=1e-12)
    # error mapping of the snippet's chk(): call-order errors -> RuntimeError with the library's message
    bare = q.heston.B200Plan("fft")
    with pytest.raises(RuntimeError, match="hb_surface_set"):
        bare.objective(X)
