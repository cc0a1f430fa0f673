"""CPU: the SABR restatement (oracle/sabr_oracle.c) against golden vectors generated from the reference
(tests/golden/make_golden_sabr.py): flavour "cpp" vs the reference's sabr.cpp compiled unmodified, flavour
"py" vs the reference's Python SABRCalibrator.sabr_implied_vol."""
import numpy as np
import pytest



@pytest.fixture(scope="module")
def g(g_sabr):
    return g_sabr


@pytest.fixture(scope="module")
def so():
    from oracle.oracle import SabrOracle

    return SabrOracle()


def _same(got, want, rtol):
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=0.0)


def test_cpp_flavour_matches_compiled_reference(g, so):
    for bi, beta in enumerate(g["betas"]):
        for ti, T in enumerate(g["Ts"]):
            got = so.vols("cpp", beta, float(g["F"]), T, g["K"], g["params"])
            _same(got, g["vols_cpp"][bi, ti], 1e-15)  # same libm, same operation order
    # the reference throws for nu < 0, |rho| >= 1, alpha <= 0, K <= 0: NaN in both
    bad = np.array([[0.3, -0.3, -0.1], [0.3, 1.0, 0.5], [0.0, 0.0, 0.5]])
    assert np.isnan(so.vols("cpp", 0.5, 100.0, 1.0, [100.0, -5.0], bad)).all()


def test_cpp_flavour_matches_live_reference(so, reference):
    rng = np.random.default_rng(3)
    params = np.column_stack([rng.uniform(0.001, 2, 200), rng.uniform(-0.99, 0.99, 200), rng.uniform(0.0, 3, 200)])
    K = np.concatenate([np.linspace(40, 250, 60), [100.0, 100.00001]])
    for beta in (0.0, 0.3, 0.5, 1.0):
        for T in (0.0, 0.1, 5.0):
            _same(so.vols("cpp", beta, 100.0, T, K, params), reference.sabr_vols(beta, 100.0, T, K, params), 1e-15)


def test_py_flavour_matches_reference_python(g, so):
    for bi, beta in enumerate(g["betas"]):
        for ti, T in enumerate(g["Ts"]):
            got = so.vols("py", beta, float(g["F"]), T, g["K"], g["params"])
            _same(got, g["vols_py"][bi, ti], 2e-14)  # numpy scalar log/sqrt vs glibc: last-bit differences


def test_objective_restatement(g, so):
    K, mv = g["smile_K"], g["smile_vol"]
    w = np.ones(K.size) / K.size  # sabr_calibrator.py:291-293
    x = np.vstack([g["cal_single"][:3], [0.3, -0.3, 0.5]])
    got = so.objective(0.5, 100.0, 0.25, K, mv, w, x)
    vols = so.vols("py", 0.5, 100.0, 0.25, K, x)
    np.testing.assert_allclose(got, ((vols - mv) ** 2 * w).sum(axis=1), rtol=1e-13)
    # at the reference's calibrated parameters the objective is rmse^2 (uniform weights)
    assert got[0] == pytest.approx(g["cal_single"][3] ** 2, rel=1e-9)
    assert got[0] <= got[1]
