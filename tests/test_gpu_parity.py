"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference's golden
vectors.  Run on the B200 box: ``python -m pytest tests -m gpu``.

Tolerance (BASELINE.json north_star): every price within 1e-10 relative of the reference, or
1e-12 absolute for deep out-of-the-money options -> ``|got - want| <= 1e-10 |want| + 1e-12``.
One documented exception: for sigma < 0.02 (the lower edge of the calibrator's box, where
kappa*theta/sigma^2 reaches 1e3..1e5) the REFERENCE's own double-precision result is only
good to ~1.4x that tolerance (measured against an 80-bit evaluation in
tests/test_conditioning.py), so agreement with it is asserted at 4x there.
"""
import numpy as np
import pytest

from oracle.oracle import MODE_FFT, MODE_REFGRID

pytestmark = pytest.mark.gpu

S0, R, Q = 100.0, 0.05, 0.02
DEFAULT = np.array([2.0, 0.04, 0.3, -0.7, 0.04])
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    return torch


def _dev(torch, a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")


def price_violation(got, want):
    """max of |got-want| / (1e-10 |want| + 1e-12); NaN pattern must agree."""
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    ok = ~np.isnan(want)
    if not ok.any():
        return 0.0
    return float(np.max(np.abs(got[ok] - want[ok]) / (1e-10 * np.abs(want[ok]) + 1e-12)))


def assert_prices(got, want, params):
    params = np.atleast_2d(params)
    for i, p in enumerate(params):
        limit = 4.0 if p[2] < 0.02 else 1.0
        v = price_violation(got[i], want[i])
        assert v <= limit, f"set {i} {p}: {v:.3g} x tolerance (limit {limit})"


def sobol_sets(n, seed=42):
    from scipy.stats import qmc

    return LB + (UB - LB) * qmc.Sobol(d=5, seed=seed).random(n)


# ---- K1: characteristic function -------------------------------------------------------------------

def test_cf_matches_reference_golden(torch_cuda, g_cf):
    from pde_b200 import characteristic_function

    torch = torch_cuda
    params = g_cf["params"]
    for v, want in ((g_cf["v_fft"], g_cf["cf_fft"]), (g_cf["v_rg"], g_cf["cf_rg"])):
        u = v + 1j * float(g_cf["ui"])
        got = characteristic_function(_dev(torch, params), _dev(torch, g_cf["T"]), torch.tensor(u, device="cuda:0"),
                                      S0=S0, r=R, q=Q).cpu().numpy()
        # relative to |phi| with a floor at the slice maximum * 1e-300 (underflowed tails compare as 0)
        scale = np.maximum(np.abs(want), 1e-290)
        rel = np.abs(got - want) / scale
        well = params[:, 2] >= 0.02
        assert rel[well].max() < 1e-10, rel[well].max()
        # sigma = 0.01 corners: the reference's phi itself carries ~2e-10 (tests/test_conditioning.py)
        assert rel[~well].max() < 2e-9, rel[~well].max()


def test_cf_general_u_and_T0(torch_cuda, g_cf):
    from pde_b200 import characteristic_function

    torch = torch_cuda
    u = torch.tensor(g_cf["u_gen"], device="cuda:0")
    got = characteristic_function(_dev(torch, g_cf["params"][:8]), _dev(torch, [0.7]), u, S0=S0, r=R, q=Q).cpu().numpy()
    want = g_cf["cf_gen"]
    assert (np.abs(got[:, 0, :] - want) / np.abs(want)).max() < 1e-10
    got0 = characteristic_function(_dev(torch, [DEFAULT]), _dev(torch, [0.0]), u, S0=S0, r=R, q=Q).cpu().numpy()
    assert (np.abs(got0[0, 0] - g_cf["cf_T0"]) / np.abs(g_cf["cf_T0"])).max() < 1e-12
    # tests/cpp/test_heston.cpp:98-112: phi(0) = 1
    one = characteristic_function(_dev(torch, [DEFAULT]), _dev(torch, [0.1, 0.5, 1.0, 2.0]),
                                  torch.zeros(1, dtype=torch.complex128, device="cuda:0"), S0=S0, r=R, q=Q).cpu().numpy()
    assert np.abs(one - 1.0).max() < 1e-10


def test_cf_vs_oracle_sobol(torch_cuda, oracle):
    from pde_b200 import characteristic_function

    torch = torch_cuda
    params = sobol_sets(256)
    Ts = np.array([0.1, 0.55, 1.0])
    v = 0.25 * np.arange(0, 4096, 7)
    got = characteristic_function(_dev(torch, params), _dev(torch, Ts), torch.tensor(v - 1.75j, device="cuda:0"),
                                  S0=S0, r=R, q=Q).cpu().numpy()
    want = oracle.cf_grid(params, Ts, v, -1.75, S0, R, Q)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-290)
    well = params[:, 2] >= 0.05
    assert rel[well].max() < 1e-10, rel[well].max()
    assert rel.max() < 2e-9


# ---- K2: the Stockham FFT alone ------------------------------------------------------------------------

@pytest.mark.parametrize("n", [4096, 512])
@pytest.mark.parametrize("n_slices", [1, 3, 149, 500])
def test_fft_batch_matches_numpy(torch_cuda, n, n_slices):
    from pde_b200 import fft_batch

    torch = torch_cuda
    rng = np.random.default_rng(n + n_slices)
    x = rng.normal(size=(n_slices, n)) + 1j * rng.normal(size=(n_slices, n))
    x[0, :] = 0
    x[0, 1] = 1.0  # a pure tone: exact answer exp(-2 pi i m / n)
    got = fft_batch(torch.tensor(x, device="cuda:0")).cpu().numpy()
    want = np.fft.fft(x, axis=1)
    scale = np.abs(want).max(axis=1, keepdims=True)
    assert (np.abs(got - want) / scale).max() < 5e-15
    # linearity: FFT(a x + b y) = a FFT(x) + b FFT(y) on device results
    if n_slices >= 3:
        z = 0.3 * x[1] - 1.7j * x[2]
        gz = fft_batch(torch.tensor(z[None, :], device="cuda:0")).cpu().numpy()[0]
        assert np.abs(gz - (0.3 * got[1] - 1.7j * got[2])).max() / np.abs(gz).max() < 1e-14


# ---- fft mode: fused K1+K2+K3 against the oracle --------------------------------------------------------

def test_fft_prices_match_oracle(torch_cuda, oracle, g_fft):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    params, K, T8 = g_fft["params"], g_fft["K50"], g_fft["T8"]
    Kf, Tf = np.tile(K, len(T8)), np.repeat(T8, len(K))
    pr = BatchPricer("fft").set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q)
    got = pr.price(_dev(torch, params)).cpu().numpy()
    want = oracle.price_batch(MODE_FFT, params, Kf, Tf, True, S0, R, Q)
    assert_prices(got, want, params)
    np.testing.assert_allclose(want.reshape(len(params), len(T8), len(K)), g_fft["fft_prices"], rtol=1e-12, atol=1e-13)
    # puts and a mixed call/put book
    for ic in (False, np.arange(Kf.size) % 3 == 0):
        pr.set_surface(Kf, Tf, ic, None, S0=S0, r=R, q=Q)
        got = pr.price(_dev(torch, params[:9])).cpu().numpy()
        assert_prices(got, oracle.price_batch(MODE_FFT, params[:9], Kf, Tf, ic, S0, R, Q), params[:9])


def test_fft_prices_match_independent_direct_dft_golden(torch_cuda, g_fft_direct):
    """The CUDA path against tests/golden/fft_direct.npz: prices built from the compiled reference's CF values
    by direct O(N) sums in 50-digit arithmetic (make_golden_fft_direct.py) -- independent of the oracle's FFT."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    g = g_fft_direct
    s0, r, q = (float(x) for x in g["spot_rate_div"])
    pricers = {}
    # one surface per (N, eta, T, parameter set): the strikes of that golden row group
    keys = sorted({(int(n), float(e), float(t), tuple(p)) for n, e, t, p in zip(g["N"], g["eta"], g["T"], g["params"])})
    for (n, eta, t, p) in keys:
        sel = np.flatnonzero((g["N"] == n) & (g["eta"] == eta) & (g["T"] == t) & (g["params"] == np.array(p)).all(axis=1))
        pr = pricers.setdefault((n, eta), BatchPricer("fft", n_grid=n, eta=eta, alpha=float(g["alpha"])))
        pr.set_surface(g["K"][sel], t, g["is_call"][sel], None, S0=s0, r=r, q=q)
        got = pr.price(_dev(torch, [p])).cpu().numpy()[0]
        v = price_violation(got, g["price"][sel])
        assert v <= (4.0 if p[2] < 0.02 else 1.0), (n, eta, t, p, v)
        assert np.array_equal(got == 0.0, g["price"][sel] == 0.0)  # clamped options stay clamped


def test_fft_prices_sobol_full_surface(torch_cuda, oracle):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    params = sobol_sets(200)  # non-split path (P >= 2 x SMs is not required: exercised below)
    Kf, Tf = np.tile(np.linspace(80, 120, 50), 32), np.repeat(np.linspace(0.1, 1.0, 32), 50)
    pr = BatchPricer("fft").set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q)
    got = pr.price(_dev(torch, params)).cpu().numpy()
    want = oracle.price_batch(MODE_FFT, params, Kf, Tf, True, S0, R, Q)
    assert_prices(got, want, params)


def test_fft_n512_and_ragged_surface(torch_cuda, oracle):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    rng = np.random.default_rng(5)
    # ragged: 1..17 strikes per maturity, unsorted option order, 7 maturities (not a multiple of 3)
    Ks, Ts = [], []
    for T in (0.08, 0.3, 0.31, 0.9, 1.7, 2.5, 3.0):
        n = int(rng.integers(1, 18))
        Ks += list(rng.uniform(40, 250, n))
        Ts += [T] * n
    perm = rng.permutation(len(Ks))
    Kf, Tf = np.array(Ks)[perm], np.array(Ts)[perm]
    ic = rng.random(Kf.size) < 0.5
    params = np.vstack([DEFAULT, sobol_sets(15, seed=3)])
    for n_grid, eta in ((4096, 0.25), (512, 0.5), (4096, 0.1)):
        pr = BatchPricer("fft", n_grid=n_grid, eta=eta).set_surface(Kf, Tf, ic, None, S0=S0, r=R, q=Q)
        got = pr.price(_dev(torch, params)).cpu().numpy()
        want = oracle.price_batch(MODE_FFT, params, Kf, Tf, ic, S0, R, Q, N=n_grid, eta=eta)
        assert_prices(got, want, params)


def test_edge_cases(torch_cuda, oracle):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    X = _dev(torch, [DEFAULT, [-1.0, 0.04, 0.3, -0.7, 0.04], [2.0, 0.04, 0.3, 1.0, 0.04]])
    for mode, omode in (("fft", MODE_FFT), ("refgrid", MODE_REFGRID)):
        pr = BatchPricer(mode)
        # T = 0 (intrinsic), invalid strike / maturity (NaN), off-grid strike (NaN in fft mode), a normal option
        K = np.array([90.0, 110.0, -5.0, 100.0, 1e-7, 100.0])
        T = np.array([0.0, 0.0, 1.0, -1.0, 1.0, 1.0])
        ic = np.array([1, 0, 1, 1, 1, 0], dtype=bool)
        mk = np.array([10.0, 10.0, 1.0, 1.0, 1.0, 6.0])
        pr.set_surface(K, T, ic, mk, S0=S0, r=R, q=Q)
        got = pr.price(X).cpu().numpy()
        want = oracle.price_batch(omode, X.cpu().numpy(), K, T, ic, S0, R, Q)
        assert np.isnan(got[1:]).all()  # invalid parameter sets: all-NaN rows
        assert price_violation(got[0], want[0]) <= 1.0
        assert got[0, 0] == 10.0 and got[0, 1] == 10.0 and np.isnan(got[0, 2]) and np.isnan(got[0, 3])
        loss = pr.objective(X).cpu().numpy()
        assert (loss == 1e10).all()  # NaN prices / invalid sets -> sentinel (heston_calibrator.py:507-508)
        # empty surface and empty batch
        pr.set_surface(np.array([]), np.array([]), True, np.array([]), S0=S0, r=R, q=Q)
        assert pr.price(X).shape == (3, 0)
        assert pr.objective(X[:1]).cpu().numpy()[0] == 0.0
        pr.set_surface(K, T, ic, mk, S0=S0, r=R, q=Q)
        assert pr.objective(X[:0]).shape == (0,)
        with pytest.raises(TypeError):
            pr.objective(X.cpu())
        pr.close()


def test_split_and_persistent_paths_agree_bitwise(torch_cuda):
    """P < 2 x SMs runs one CTA per (set, group) + a finalize kernel; larger P runs one persistent
    CTA per set with in-kernel finalize.  Same arithmetic -> identical bits."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    Kf, Tf = np.tile(np.linspace(80, 120, 13), 8), np.repeat(np.linspace(0.1, 1.5, 8), 13)
    rng = np.random.default_rng(0)
    # The persistent path also skips stage B/F of perturbed classes on decayed tails (kernels.cuh,
    # track_tail); the split path cannot (one job per group), so this comparison is skip vs no-skip.
    corners = np.array([np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)])
    params = np.vstack([sobol_sets(64), corners])
    n_small = len(params)
    big = np.vstack([params, sobol_sets(512, seed=9)])
    for mode in ("fft", "refgrid"):
        pr = BatchPricer(mode)
        mk = pr.set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q).price(_dev(torch, [DEFAULT])).cpu().numpy()[0]
        mk = np.maximum(mk * (1 + 0.01 * rng.normal(size=mk.size)), 0.01)
        pr.set_surface(Kf, Tf, True, mk, S0=S0, r=R, q=Q)
        for fn in (pr.price, pr.objective, pr.normal_equations):
            a = fn(_dev(torch, params)).cpu().numpy()
            b = fn(_dev(torch, big)).cpu().numpy()[:n_small]
            assert np.array_equal(a, b, equal_nan=True), (mode, fn.__name__)


# ---- refgrid mode: against the reference's own outputs ---------------------------------------------------------

def test_refgrid_prices_match_reference_golden(torch_cuda, g_prices):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    K, T8, T32 = g_prices["K50"], g_prices["T8"], g_prices["T32"]
    params = g_prices["params"]
    pr = BatchPricer("refgrid").set_surface(np.tile(K, 8), np.repeat(T8, 50), True, None, S0=S0, r=R, q=Q)
    got = pr.price(_dev(torch, params)).cpu().numpy().reshape(len(params), 8, 50)
    assert_prices(got.reshape(len(params), -1), g_prices["surf_sets"].reshape(len(params), -1), params)
    pr.set_surface(np.tile(K, 32), np.repeat(T32, 50), True, None, S0=S0, r=R, q=Q)
    got = pr.price(_dev(torch, [DEFAULT])).cpu().numpy()
    assert price_violation(got[0], g_prices["surf_default"].ravel()) <= 1.0
    pr.set_surface(np.tile(K, 8), np.repeat(T8, 50), False, None, S0=S0, r=R, q=Q)
    got = pr.price(_dev(torch, [DEFAULT])).cpu().numpy()
    assert price_violation(got[0], g_prices["puts_default"].ravel()) <= 1.0
    # the non-monotone, clamped tail of the reference is reproduced
    for i, T in enumerate((0.1, 1.0)):
        pr.set_surface(g_prices["K_wide"], T, True, None, S0=S0, r=R, q=Q)
        got = pr.price(_dev(torch, [DEFAULT])).cpu().numpy()[0]
        assert price_violation(got, g_prices["wide"][i]) <= 1.0
        assert np.array_equal(got == 0.0, g_prices["wide"][i] == 0.0)


def test_calibrator_objective_residuals_jacobian_match_reference(torch_cuda, g_cal):
    """Values produced by the reference's Python calibrator + SciPy (tests/golden/make_golden.py)."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    K, T, mkt, ic, xs = g_cal["K"], g_cal["T"], g_cal["market"], g_cal["is_call"], g_cal["xs"]
    pr = BatchPricer("refgrid").set_surface(K, T, ic, mkt, S0=S0, r=R, q=Q).set_bounds(g_cal["lb"], g_cal["ub"])
    X = _dev(torch, xs)
    prices = pr.price(X).cpu().numpy()
    assert_prices(prices, g_cal["prices"], xs)
    loss = pr.objective(X).cpu().numpy()
    np.testing.assert_allclose(loss, g_cal["objective"], rtol=1e-9)  # loss 1e-9 rel
    assert np.array_equal(loss == 1e10, g_cal["objective"] == 1e10)
    res, jac = pr.jacobian(X)
    res, jac = res.cpu().numpy(), jac.cpu().numpy()
    np.testing.assert_allclose(res, g_cal["residuals"], rtol=1e-9, atol=1e-12)
    # J = (r(x+h) - r(x))/dx with dx ~ 1.5e-8 max(1,|x|): a price difference at the parity
    # tolerance moves J by 2 tol/(|dx| market) (SURVEY.md section 7 "FD-Jacobian parity")
    dx = 1.4901161193847656e-08 * np.maximum(1.0, np.abs(xs))  # [P,5]
    tol_price = 1e-10 * np.abs(g_cal["prices"]) + 1e-12  # [P,n]
    atol = 2.0 * tol_price[:, :, None] / (dx[:, None, :] * mkt[None, :, None])
    assert (np.abs(jac - g_cal["jacobian"]) <= atol + 1e-9 * np.abs(g_cal["jacobian"])).all()
    mixed = pr.set_surface(K, T, g_cal["is_call_mixed"], None, S0=S0, r=R, q=Q).price(_dev(torch, [DEFAULT]))
    assert price_violation(mixed.cpu().numpy()[0], g_cal["prices_mixed"]) <= 1.0


@pytest.mark.parametrize("mode,omode", [("fft", MODE_FFT), ("refgrid", MODE_REFGRID)])
def test_normal_equations_match_oracle(torch_cuda, oracle, g_cal, mode, omode):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    rng = np.random.default_rng(1)
    Kf, Tf = np.tile(np.linspace(85, 115, 11), 5), np.repeat(np.linspace(0.2, 1.2, 5), 11)
    truth = np.array([1.8, 0.06, 0.45, -0.6, 0.05])
    mk = oracle.price_batch(omode, [truth], Kf, Tf, True, S0, R, Q)[0] * (1 + 0.002 * rng.normal(size=Kf.size))
    xs = np.vstack([truth, truth * 1.1, sobol_sets(6, seed=5), LB, UB, [10.0, 1.0, 2.0, -0.99, 1.0]])
    pr = BatchPricer(mode).set_surface(Kf, Tf, True, mk, S0=S0, r=R, q=Q)
    got = pr.normal_equations(_dev(torch, xs)).cpu().numpy()
    want = oracle.normal_eq_batch(omode, xs, LB, UB, Kf, Tf, True, mk, S0, R, Q)
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=1e-9)  # loss
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=1e-9)  # ||r||^2
    # J^T r and J^T J inherit the finite-difference amplification: a price
    # discrepancy dp moves J_ic by 2 dp/(|dx_c| market_i).  Bound the blocks from first principles
    # with dp = 2e-12 |price| + 2e-14, i.e. 50x tighter than the contractual price tolerance.
    iu = np.triu_indices(5)
    prices = oracle.price_batch(omode, xs, Kf, Tf, True, S0, R, Q)
    for i, x in enumerate(xs):
        if not np.isfinite(want[i, 1:]).all():
            assert np.array_equal(np.isfinite(got[i]), np.isfinite(want[i]))
            continue
        r0, J = oracle.jacobian(omode, x, LB, UB, Kf, Tf, True, mk, S0, R, Q)
        dx = np.abs(oracle.fd_steps(x, LB, UB))
        dJ = 2.0 * (2e-12 * np.abs(prices[i]) + 2e-14)[:, None] / (dx[None, :] * mk[:, None])  # [n,5]
        dr = (2e-12 * np.abs(prices[i]) + 2e-14) / mk
        b_jtr = dJ.T @ np.abs(r0) + np.abs(J).T @ dr
        b_jtj = (np.abs(J).T @ dJ + dJ.T @ np.abs(J))[iu]
        assert (np.abs(got[i, 2:7] - want[i, 2:7]) <= b_jtr + 1e-9 * np.abs(want[i, 2:7])).all(), (i, x)
        assert (np.abs(got[i, 7:] - want[i, 7:]) <= b_jtj + 1e-9 * np.abs(want[i, 7:])).all(), (i, x)


# ---- significance cut (hb_plan_set_truncation) ------------------------------------------------------------------------

def test_significance_cut_is_bounded_by_its_own_error_budget(torch_cuda, oracle):
    """Default plans drop grid points whose |phi| is below e^cut (cut chosen so that ALL dropped points together
    cannot move a price by more than 2^-80); exact mode (0) only skips true exp underflow.  The two must agree
    far inside the parity tolerance, on both launch paths, including the slow-decay corner (nothing dropped)
    and the fast-decay one (almost everything dropped)."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    Kf, Tf = np.tile(np.linspace(80, 120, 50), 32), np.repeat(np.linspace(0.1, 1.0, 32), 50)
    corners = np.array([np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)])
    small = np.vstack([DEFAULT, corners, sobol_sets(31)])          # 64 sets: split path
    big = np.vstack([small, sobol_sets(512, seed=3)])              # persistent path
    cut = BatchPricer("fft").set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q)
    exact = BatchPricer("fft").set_truncation(0.0).set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q)
    assert exact.log_cut == -746.0 and -64.0 < cut.log_cut < -40.0
    # the cut the plan reports is the one its bound asks for: e^cut * sum|tab| * max scale = 2^-80
    v = 0.25 * np.arange(4096)
    w = (0.25 / 3) * np.where(np.arange(4096) == 0, 1.0, np.where(np.arange(4096) % 2 == 1, 4.0, 2.0))
    wsum = np.sum(w / np.abs(0.75 ** 2 + 0.75 - v * v + 1j * 2.5 * v))
    lam, b = 2 * np.pi / (4096 * 0.25), np.pi / 0.25
    km = -b + lam * np.floor((np.log(Kf) + b) / lam)
    scale = np.max(np.exp(-R * Tf) * np.exp(-0.75 * km) / np.pi)
    assert abs(cut.log_cut - np.log(2.0 ** -80 / (wsum * scale))) < 1e-9
    for xs in (small, big):
        a = cut.price(_dev(torch, xs)).cpu().numpy()
        e = exact.price(_dev(torch, xs)).cpu().numpy()
        fin = np.isfinite(e)
        assert np.array_equal(np.isfinite(a), fin)
        # budget 2^-80 absolute; what is left is rounding: default plans run the direct-sum kernel (live prefix, plain
        # sums in chunk order), exact plans the radix-8 transform -- two summation orders of the same terms.  Measured
        # 1e-12 relative at worst (sigma = 2, rho = -0.99 corner, a deep OTM price of 3e-3): 1 % of the parity budget.
        assert np.max(np.abs(a[fin] - e[fin]) / (1e-10 * np.abs(e[fin]) + 1e-12)) <= 0.05
    want = oracle.price_batch(MODE_FFT, small, Kf, Tf, True, S0, R, Q)
    assert_prices(cut.price(_dev(torch, small)).cpu().numpy(), want, small)
    assert_prices(exact.price(_dev(torch, small)).cpu().numpy(), want, small)
    with pytest.raises(ValueError):
        cut.set_truncation(1e-6)


# ---- the benched path itself: C3 surface, persistent kernel, work elision live ----------------------------------

def c3_benched_path_sets():
    """>= 2 x SMs sets so the persistent launch path runs (dead masks / tail skip / asymptotic and series stage B
    live): 512 Sobol points of the calibrator's box, its 32 corners, and a slow-decay cluster (sigma near 2,
    rho near -0.99, small kappa theta: SURVEY.md App. D "no decay") where nothing can be elided."""
    from scipy.stats import qmc

    corners = np.array([np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)])
    lo = np.array([0.1, 0.01, 1.5, -0.99, 0.01])
    hi = np.array([0.5, 0.05, 2.0, -0.90, 0.05])
    slow = lo + (hi - lo) * qmc.Sobol(d=5, seed=11).random(64)
    return np.vstack([sobol_sets(512), corners, slow])


_C3_WANT = {}


@pytest.mark.parametrize("route", ["routed", "all_direct", "all_transform"])
def test_benched_path_normal_equations_match_oracle_on_c3_surface(torch_cuda, oracle, monkeypatch, route):
    """BASELINE.json config 3's own shape (32 maturities x 50 strikes, N = 4096, default bounds) on the launch
    path bench.py times against oracle.normal_eq_batch: all 22 columns, and hb_jacobian's residuals / Jacobian on
    the same sets.  `routed` is the product default (prefix scan, short-prefix sets on direct_job_kernel, the
    slow-decay cluster and the long corners on fft_job_kernel); `all_direct` forces every set -- including the
    ones whose live prefix is the whole grid, i.e. maturities cut over many waves -- through the direct-sum kernel,
    `all_transform` through the transform kernel (persistent path, tail skip / asymptotic and series stage B live)."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    if route == "all_direct":
        monkeypatch.setenv("HB_DIRECT_THR", "0")
    elif route == "all_transform":
        monkeypatch.setenv("HB_DIRECT", "0")
    oracle.use_all_cores()
    xs = c3_benched_path_sets()
    assert len(xs) >= 2 * torch.cuda.get_device_properties(0).multi_processor_count  # persistent path
    Kf, Tf = np.tile(np.linspace(80, 120, 50), 32), np.repeat(np.linspace(0.1, 1.0, 32), 50)
    mk = oracle.price_batch(MODE_FFT, [DEFAULT], Kf, Tf, True, S0, R, Q)[0]
    mk = np.maximum(mk * (1 + 0.001 * np.random.default_rng(42).normal(size=mk.size)), 0.01)
    pr = BatchPricer("fft").set_surface(Kf, Tf, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
    X = _dev(torch, xs)
    got = pr.normal_equations(X).cpu().numpy()
    res, jac = pr.jacobian(X)
    res, jac = res.cpu().numpy(), jac.cpu().numpy()
    prices = pr.price(X).cpu().numpy()
    if route != "all_transform":  # the launch was routed as intended
        pr.profile(True)
        pr.normal_equations(X)
        rt = pr.profile_read()
        pr.profile(False)
        assert rt["direct_ms"] > 0.0
        if route == "routed":
            assert rt["sets_direct"] + rt["sets_transform"] == len(xs)
            assert 64 <= rt["sets_transform"] < 200 and rt["transform_ms"] > 0.0 and rt["scan_ms"] > 0.0  # slow cluster + corners
        else:  # no scan, no second kernel
            assert rt["sets_direct"] == -1 and rt["transform_ms"] == 0.0 and rt["scan_ms"] == 0.0
    if "want" not in _C3_WANT:
        _C3_WANT["want"] = oracle.normal_eq_batch(MODE_FFT, xs, LB, UB, Kf, Tf, True, mk, S0, R, Q)
    want = _C3_WANT["want"]

    fin = np.isfinite(want).all(axis=1) & (want[:, 0] != 1e10)
    assert fin.sum() >= 400  # most of the population is an ordinary candidate
    # sentinel / non-finite pattern (moment explosion -> inf, clamped prices -> 1e10) must be the oracle's
    assert np.array_equal(got[:, 0] == 1e10, want[:, 0] == 1e10)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    np.testing.assert_allclose(got[fin, 0], want[fin, 0], rtol=1e-9)  # loss, SURVEY.md 8c budget
    np.testing.assert_allclose(got[fin, 1], want[fin, 1], rtol=1e-9)  # ||r||^2
    # J^T r / J^T J: bounded from first principles as in test_normal_equations_match_oracle, price discrepancy
    # dp = 2e-12 |p| + 2e-14 (50x tighter than the contractual tolerance), |J| and |r| from the device Jacobian
    # Small sigma (kappa theta / sigma^2 up to 1e5) is the documented conditioning corner where the REFERENCE's own
    # double-precision CF loses digits (tests/test_conditioning.py): there dp is the contractual tolerance itself
    # (1e-10 |p| + 1e-12, x4 below sigma = 0.02).  Measured on B200: worst 0.39 of that; everywhere else 0.52 of
    # the 50x tighter bound.
    iu = np.triu_indices(5)
    worst = [0.0, 0.0]
    for i in np.flatnonzero(fin):
        small_sigma = xs[i][2] < 0.05
        scale = (4.0 if xs[i][2] < 0.02 else 1.0) * 50.0 if small_sigma else 1.0
        dx = np.abs(oracle.fd_steps(xs[i], LB, UB))
        dp = scale * (2e-12 * np.abs(prices[i]) + 2e-14)
        dJ = 2.0 * dp[:, None] / (dx[None, :] * mk[:, None])
        dr = dp / mk
        aJ, ar = np.abs(jac[i]), np.abs(res[i])
        b_jtr = dJ.T @ ar + aJ.T @ dr + 1e-9 * np.abs(want[i, 2:7])
        b_jtj = (aJ.T @ dJ + dJ.T @ aJ)[iu] + 1e-9 * np.abs(want[i, 7:])
        w = max((np.abs(got[i, 2:7] - want[i, 2:7]) / b_jtr).max(), (np.abs(got[i, 7:] - want[i, 7:]) / b_jtj).max())
        worst[int(small_sigma)] = max(worst[int(small_sigma)], w)
        assert w <= 1.0, (i, xs[i], w)
    # hb_jacobian on the same path: residuals within the price tolerance, Jacobian within its FD amplification,
    # against the oracle's own residuals / SciPy-rule Jacobian on a sample that includes corners and slow-decay sets
    sample = np.concatenate([np.arange(0, 512, 37), np.arange(512, 544, 5), np.arange(544, 608, 13)])
    for i in sample:
        if not fin[i]:
            continue
        r0, J = oracle.jacobian(MODE_FFT, xs[i], LB, UB, Kf, Tf, True, mk, S0, R, Q)
        limit = 4.0 if xs[i][2] < 0.02 else 1.0
        tol_p = 1e-10 * np.abs(prices[i]) + 1e-12
        assert (np.abs(res[i] - r0) <= limit * tol_p / mk).all(), (i, xs[i])
        dx = np.abs(oracle.fd_steps(xs[i], LB, UB))
        assert (np.abs(jac[i] - J) <= limit * 2.0 * tol_p[:, None] / (dx[None, :] * mk[:, None]) + 1e-9 * np.abs(J)).all(), (i, xs[i])
    # and the device's normal equations are its own Jacobian contracted (same kernel, different epilogue)
    for i in sample:
        if fin[i]:
            np.testing.assert_allclose(got[i, 2:7], jac[i].T @ res[i], rtol=1e-9, atol=1e-9 * np.abs(got[i, 2:7]).max())
            np.testing.assert_allclose(got[i, 7:], (jac[i].T @ jac[i])[iu], rtol=1e-9, atol=1e-9 * np.abs(got[i, 7:]).max())


# ---- full-size properties (BASELINE.json config 3: 65,536 sets x 32 maturities, N = 4096) -------------------------

def test_full_size_objective_properties(torch_cuda, oracle):
    from pde_b200 import BatchPricer

    torch = torch_cuda
    P = 65536
    params = sobol_sets(P)
    Kf, Tf = np.tile(np.linspace(80, 120, 50), 32), np.repeat(np.linspace(0.1, 1.0, 32), 50)
    mk = oracle.price_batch(MODE_FFT, [DEFAULT], Kf, Tf, True, S0, R, Q)[0]
    rng = np.random.default_rng(42)
    mk = np.maximum(mk * (1 + 0.001 * rng.normal(size=mk.size)), 0.01)
    pr = BatchPricer("fft").set_surface(Kf, Tf, True, mk, S0=S0, r=R, q=Q)
    X = _dev(torch, params)
    loss = pr.objective(X)
    # idempotence / determinism
    assert torch.equal(loss, pr.objective(X))
    # permutation equivariance: shuffled candidates give the shuffled losses, bit for bit
    perm = torch.randperm(P, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(1))
    assert torch.equal(pr.objective(X[perm]), loss[perm])
    # duplicated candidates agree; a sample agrees with the oracle at 1e-9
    idx = np.concatenate([np.arange(16), rng.integers(0, P, 48)])
    want = oracle.objective_batch(MODE_FFT, params[idx], Kf, Tf, True, mk, S0, R, Q)
    got = loss.cpu().numpy()[idx]
    assert np.array_equal(got == 1e10, want == 1e10)
    np.testing.assert_allclose(got, want, rtol=1e-9)  # loss 1e-9 relative
    # non-finite losses (moment explosion of E[S^1.75] -> inf prices -> inf loss in the reference too,
    # heston_calibrator.py:507-511 only screens NaN and <= 0) must be the oracle's non-finite losses
    all_loss = loss.cpu().numpy()
    odd = np.flatnonzero(~np.isfinite(all_loss))[:24]
    if odd.size:
        want_odd = oracle.objective_batch(MODE_FFT, params[odd], Kf, Tf, True, mk, S0, R, Q)
        assert np.array_equal(all_loss[odd], want_odd, equal_nan=True), (params[odd], all_loss[odd], want_odd)


def test_full_size_normal_equations_properties(torch_cuda):
    """BASELINE.json config 3 at full size (65,536 sets x 32 maturities x 50 strikes, FD blocks): properties
    that do not need the oracle.  The six-variant and the one-variant job kernels are different instantiations with
    different CTA shapes, wave capacities and work elision (and a set may be routed to the direct-sum kernel by one and
    to the transform kernel by the other only if their live-prefix tables differ): the loss column of the normal
    equations must still be the objective's loss bit for bit."""
    from pde_b200 import BatchPricer

    torch = torch_cuda
    P = 65536
    X = _dev(torch, sobol_sets(P))
    Kf, Tf = np.tile(np.linspace(80, 120, 50), 32), np.repeat(np.linspace(0.1, 1.0, 32), 50)
    pr = BatchPricer("fft").set_surface(Kf, Tf, True, None, S0=S0, r=R, q=Q)
    mk = pr.price(_dev(torch, [DEFAULT])).cpu().numpy()[0]
    mk = np.maximum(mk * (1 + 0.001 * np.random.default_rng(42).normal(size=mk.size)), 0.01)
    pr.set_surface(Kf, Tf, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
    neq = pr.normal_equations(X)
    loss = pr.objective(X)
    a, b = neq[:, 0].cpu().numpy(), loss.cpu().numpy()
    assert np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(a == 1e10, b == 1e10)
    ok = np.isfinite(a)
    # Bit for bit wherever both kernels see the same live-prefix table.  The six-variant table is the maximum over the
    # four classes, the one-variant table that of the base class: where a perturbed class's bound crosses a block
    # boundary the base slice is summed in other chunks (or the set is routed to the other kernel) and the loss may
    # differ by rounding -- measured on this population: nowhere.
    rel = np.abs(a[ok] - b[ok]) / np.abs(b[ok])
    assert (rel == 0.0).mean() >= 0.999 and rel.max() <= 1e-12, (float((rel == 0.0).mean()), float(rel.max()))
    # determinism and permutation equivariance of all 22 columns
    perm = torch.randperm(P, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(2))
    assert torch.equal(pr.normal_equations(X[perm]).view(torch.int64), neq[perm].view(torch.int64))  # NaN-safe
    # ||r||^2 >= 0, diag(J^T J) >= 0, Cauchy-Schwarz |(J^T J)_ab| <= sqrt((J^T J)_aa (J^T J)_bb) wherever finite
    n = neq.cpu().numpy()
    fin = np.isfinite(n).all(axis=1)
    assert fin.mean() > 0.9
    n = n[fin]
    iu = np.triu_indices(5)
    A = np.zeros((len(n), 5, 5))
    A[:, iu[0], iu[1]] = n[:, 7:]
    d = np.sqrt(np.einsum("pii->pi", A))
    assert (n[:, 1] >= 0).all() and (np.einsum("pii->pi", A) >= 0).all()
    assert (np.abs(A[:, iu[0], iu[1]]) <= d[:, iu[0]] * d[:, iu[1]] * (1 + 1e-12) + 1e-300).all()
    # |J^T r| <= ||J_a|| ||r||
    assert (np.abs(n[:, 2:7]) <= d * np.sqrt(n[:, 1])[:, None] * (1 + 1e-12) + 1e-300).all()
