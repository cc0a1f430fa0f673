"""CPU: pin the oracle (oracle/heston_oracle.c) against the reference's outputs.

Golden vectors come from the reference's own heston.cpp / Python calibrator
(tests/golden/make_golden.py).  The reference's tests hold no numeric pins, so
these fixtures plus SURVEY.md Appendix D's known answers are the pin.
"""
import numpy as np
import pytest

from oracle.oracle import MODE_FFT, MODE_REFGRID

S0, R, Q = 100.0, 0.05, 0.02
DEFAULT = [2.0, 0.04, 0.3, -0.7, 0.04]


def test_known_answers_appendix_d(oracle):
    # SURVEY.md Appendix D (compiled reference, this container)
    assert oracle.price_refgrid(DEFAULT, 100, 1.0, S0, R, Q) == pytest.approx(8.853337653490936, rel=1e-14)
    assert oracle.price_refgrid(DEFAULT, 100, 1.0, S0, R, Q, False) == pytest.approx(5.9564127728868215, rel=1e-14)
    assert oracle.price_refgrid(DEFAULT, 200, 1.0, S0, R, Q) == pytest.approx(0.0035939693107424884, rel=1e-10)
    assert oracle.price_refgrid(DEFAULT, 120, 0.1, S0, R, Q) == 0.0
    phi = oracle.cf(DEFAULT, 1 - 1.75j, 1.0, S0, R, Q)
    assert phi.real == pytest.approx(-108.69976080392448, rel=1e-14)
    assert phi.imag == pytest.approx(-3357.5708889312546, rel=1e-14)
    corner = [0.1, 0.01, 2.0, -0.99, 0.01]
    assert oracle.price_refgrid(corner, 100, 1.0, S0, R, Q) == pytest.approx(4.532349033806166, rel=1e-13)


def test_cf_matches_reference_golden(oracle, g_cf):
    got_fft = oracle.cf_grid(g_cf["params"], g_cf["T"], g_cf["v_fft"], float(g_cf["ui"]), S0, R, Q)
    got_rg = oracle.cf_grid(g_cf["params"], g_cf["T"], g_cf["v_rg"], float(g_cf["ui"]), S0, R, Q)
    # same libgcc/glibc routines in the same order: bit-identical on x86-64
    assert np.array_equal(got_fft, g_cf["cf_fft"])
    assert np.array_equal(got_rg, g_cf["cf_rg"])


def test_cf_general_u_and_T0(oracle, g_cf):
    for i, p in enumerate(g_cf["params"][:8]):
        for u, want in zip(g_cf["u_gen"], g_cf["cf_gen"][i]):
            assert oracle.cf(p, u, 0.7, S0, R, Q) == want
    for u, want in zip(g_cf["u_gen"], g_cf["cf_T0"]):
        assert oracle.cf(DEFAULT, u, 0.0, S0, R, Q) == want


def test_cf_properties_from_reference_tests(oracle):
    # tests/cpp/test_heston.cpp:98-130: phi(0)=1, phi at T=0
    for T in (0.1, 0.5, 1.0, 2.0):
        assert abs(oracle.cf(DEFAULT, 0j, T, S0, R, Q) - 1.0) < 1e-10
    z = oracle.cf(DEFAULT, 1.0 + 0j, 0.0, S0, R, Q)
    assert z == pytest.approx(np.exp(1j * np.log(S0)), abs=1e-10)


def test_refgrid_prices_match_reference_golden(oracle, g_prices):
    K, T8 = g_prices["K50"], g_prices["T8"]
    Kf, Tf = np.tile(K, len(T8)), np.repeat(T8, len(K))
    got = oracle.price_batch(MODE_REFGRID, g_prices["params"], Kf, Tf, True, S0, R, Q)
    assert np.array_equal(got.reshape(-1, len(T8), len(K)), g_prices["surf_sets"])
    T32 = g_prices["T32"]
    got = oracle.price_batch(MODE_REFGRID, [DEFAULT], np.tile(K, 32), np.repeat(T32, 50), True, S0, R, Q)
    assert np.array_equal(got.reshape(32, 50), g_prices["surf_default"])
    got = oracle.price_batch(MODE_REFGRID, [DEFAULT], Kf, Tf, False, S0, R, Q)
    assert np.array_equal(got.reshape(len(T8), 50), g_prices["puts_default"])
    for i, T in enumerate((0.1, 1.0)):
        got = oracle.price_batch(MODE_REFGRID, [DEFAULT], g_prices["K_wide"], T, True, S0, R, Q)[0]
        assert np.array_equal(got, g_prices["wide"][i])


def test_t0_intrinsic(oracle, g_misc):
    got = [oracle.price_refgrid(DEFAULT, K, 0.0, S0, R, Q, c) for K in (90.0, 110.0) for c in (True, False)]
    assert np.array_equal(got, g_misc["t0"])
    assert got == [10.0, 0.0, 0.0, 10.0]


def test_calibrator_prices_objective_residuals(oracle, g_cal):
    K, T, mkt, ic = g_cal["K"], g_cal["T"], g_cal["market"], g_cal["is_call"]
    pr = oracle.price_batch(MODE_REFGRID, g_cal["xs"], K, T, ic, S0, R, Q)
    assert np.array_equal(pr, g_cal["prices"])
    obj = oracle.objective_batch(MODE_REFGRID, g_cal["xs"], K, T, ic, mkt, S0, R, Q)
    np.testing.assert_allclose(obj, g_cal["objective"], rtol=1e-13)
    assert (g_cal["objective"] == 1e10).sum() >= 3  # the sentinel path is exercised
    for i in range(len(g_cal["xs"])):
        res = oracle.residuals_from_prices(pr[i], mkt)
        np.testing.assert_allclose(res, g_cal["residuals"][i], rtol=1e-14, atol=0)
    mixed = oracle.price_batch(MODE_REFGRID, [DEFAULT], K, T, g_cal["is_call_mixed"], S0, R, Q)[0]
    assert np.array_equal(mixed, g_cal["prices_mixed"])


def test_fd_jacobian_matches_scipy_on_reference(oracle, g_cal):
    K, T, mkt, ic = g_cal["K"], g_cal["T"], g_cal["market"], g_cal["is_call"]
    for i, x in enumerate(g_cal["xs"]):
        r0, J = oracle.jacobian(MODE_REFGRID, x, g_cal["lb"], g_cal["ub"], K, T, ic, mkt, S0, R, Q)
        np.testing.assert_allclose(r0, g_cal["residuals"][i], rtol=1e-14)
        # identical prices -> identical differences; only the division order can differ
        np.testing.assert_allclose(J, g_cal["jacobian"][i], rtol=1e-12, atol=1e-300)


def test_fd_step_rule_matches_scipy(oracle):
    from scipy.optimize._numdiff import _adjust_scheme_to_bounds, _compute_absolute_step

    lb = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
    ub = np.array([10.0, 1.0, 2.0, 0.99, 1.0])
    rng = np.random.default_rng(3)
    xs = [lb, ub, lb + 1e-9, ub - 1e-9, np.array([2.0, 0.04, 0.3, -0.7, 0.04]), np.array([10, 1, 0.01, 0, 1.0])]
    xs += [lb + (ub - lb) * rng.random(5) for _ in range(50)]
    for x in xs:
        x = np.asarray(x, dtype=float)
        h = _compute_absolute_step(None, x, np.zeros(1), "2-point")
        h, _ = _adjust_scheme_to_bounds(x, h, 1, "1-sided", lb, ub)
        assert np.array_equal(oracle.fd_steps(x, lb, ub), h)
    # degenerate box narrower than the step: neither side fits
    lb2, ub2 = np.full(5, 1.0), np.full(5, 1.0 + 1e-9)
    x = np.full(5, 1.0 + 2e-10)
    h = _compute_absolute_step(None, x, np.zeros(1), "2-point")
    h, _ = _adjust_scheme_to_bounds(x, h, 1, "1-sided", lb2, ub2)
    assert np.array_equal(oracle.fd_steps(x, lb2, ub2), h)


def test_normal_equations_consistent(oracle, g_cal):
    K, T, mkt, ic = g_cal["K"], g_cal["T"], g_cal["market"], g_cal["is_call"]
    ne = oracle.normal_eq_batch(MODE_REFGRID, g_cal["xs"], g_cal["lb"], g_cal["ub"], K, T, ic, mkt, S0, R, Q)
    iu = np.triu_indices(5)
    for i in range(len(g_cal["xs"])):
        J, r0 = g_cal["jacobian"][i], g_cal["residuals"][i]
        assert ne[i, 0] == pytest.approx(g_cal["objective"][i], rel=1e-13)
        assert ne[i, 1] == pytest.approx(r0 @ r0, rel=1e-12)
        np.testing.assert_allclose(ne[i, 2:7], J.T @ r0, rtol=1e-9)
        np.testing.assert_allclose(ne[i, 7:], (J.T @ J)[iu], rtol=1e-9)


# ---- FFT mode: no reference implementation exists; pinned to the reference's CF + an independent transform ----

def test_fft_mode_matches_independent_direct_dft_golden(oracle, g_fft_direct):
    """tests/golden/fft_direct.npz (make_golden_fft_direct.py): the compiled reference's CF values pushed
    through the written Carr-Madan specification by O(N) direct sums at the two bracketing bins in 50-digit
    arithmetic -- no FFT, no code shared with heston_oracle.c.  Pins Simpson weights, transform sign and
    normalisation, grid origin, bin / fraction of each strike, e^{-alpha k}/pi e^{-rT} scale, clamp and parity."""
    g = g_fft_direct
    s0, r, q = g["spot_rate_div"]
    worst = 0.0
    for i in range(g["price"].size):
        got = oracle.price_batch(MODE_FFT, [g["params"][i]], [g["K"][i]], [g["T"][i]], [bool(g["is_call"][i])], s0, r, q,
                                 N=int(g["N"][i]), eta=float(g["eta"][i]), alpha=float(g["alpha"]))[0, 0]
        want = g["price"][i]
        worst = max(worst, abs(got - want) / (1e-12 * abs(want) + 1e-14))
    assert worst <= 1.0, worst
    assert set(np.unique(g["N"])) == {512, 4096, 16384} and (g["price"] == 0.0).any() and (~g["is_call"]).any()


def test_fft_mode_selfcheck_and_numpy_fft(oracle, g_fft):
    K, T8 = g_fft["K50"], g_fft["T8"]
    got = oracle.price_batch(MODE_FFT, g_fft["params"], np.tile(K, len(T8)), np.repeat(T8, len(K)), True, S0, R, Q,
                             N=int(g_fft["N"]), eta=float(g_fft["eta"]), alpha=float(g_fft["alpha"]))
    got = got.reshape(-1, len(T8), len(K))
    np.testing.assert_allclose(got, g_fft["fft_prices"], rtol=1e-12, atol=1e-13)
    # radix-2 restatement vs numpy.fft on the reference's CF values
    np.testing.assert_allclose(got[:4], g_fft["numpy_fft_prices"], rtol=1e-11, atol=1e-12)


def test_fft_mode_converges_to_semi_analytic_price(oracle):
    """The Carr-Madan spec vs direct numerical integration of the same damped integrand
    (scipy.integrate.quad on the reference-pinned CF): ~3e-3 at eta=0.25."""
    from scipy.integrate import quad

    alpha = 0.75
    for K, T in [(90.0, 0.5), (100.0, 1.0), (110.0, 1.0)]:
        k = np.log(K)

        def f(v):
            phi = oracle.cf(DEFAULT, complex(v, -(alpha + 1)), T, S0, R, Q)
            return (np.exp(-1j * v * k) * phi / (alpha * alpha + alpha - v * v + 1j * (2 * alpha + 1) * v)).real

        integral = quad(f, 0, 400, limit=2000, epsabs=1e-11, epsrel=1e-11)[0]
        exact = np.exp(-alpha * k) / np.pi * np.exp(-R * T) * integral
        got = oracle.price_batch(MODE_FFT, [DEFAULT], [K], T, True, S0, R, Q)[0, 0]
        assert abs(got - exact) < 5e-3
        fine = oracle.price_batch(MODE_FFT, [DEFAULT], [K], T, True, S0, R, Q, N=16384, eta=0.0625)[0, 0]
        assert abs(fine - exact) < 1e-3  # linear interpolation error at lambda = 2pi/(N eta)


def test_fft_out_of_grid_strike_is_nan(oracle):
    # ln K beyond b = pi/eta -> NaN (then 1e10 through the objective)
    pr = oracle.price_batch(MODE_FFT, [DEFAULT], [1e-7, 100.0, 1e7], 1.0, True, S0, R, Q)[0]
    assert np.isnan(pr[0]) and np.isfinite(pr[1]) and np.isnan(pr[2])


# ---- live reference (prebuilt oracle/_ref/libheston_ref.so) -----------------------------------------

def test_oracle_bitwise_vs_live_reference(oracle, reference):
    rng = np.random.default_rng(11)
    lb = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
    ub = np.array([10.0, 1.0, 2.0, 0.99, 1.0])
    v = 0.25 * np.arange(0, 4096, 5)
    K = np.linspace(60, 160, 9)
    for _ in range(25):
        p = lb + (ub - lb) * rng.random(5)
        T = float(rng.uniform(0.02, 3.0))
        assert np.array_equal(reference.cf_grid(p, v, -1.75, T, S0, R, Q), oracle.cf_grid(p, [T], v, -1.75, S0, R, Q)[0, 0])
        a = reference.price_options(p, K, T, S0, R, Q)
        b = oracle.price_batch(MODE_REFGRID, [p], K, T, True, S0, R, Q)[0]
        assert np.array_equal(a, b)


def test_reference_error_strings(reference, g_misc):
    msgs = dict(zip(g_misc["err_keys"].tolist(), g_misc["err_msgs"].tolist()))
    assert msgs["kappa"] == "Heston: kappa must be positive, got -1.000000"
    with pytest.raises(ValueError, match="kappa must be positive"):
        reference.validate([-1.0, 0.04, 0.3, -0.7, 0.04])
    assert msgs["strike"] == "Strike must be positive"
    assert msgs["spot"] == "Spot must be positive"
    assert msgs["maturity"] == "Maturity must be non-negative"
