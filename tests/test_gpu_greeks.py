"""GPU: batched finite-difference Greeks against the compiled reference's
HestonModel::price_option_with_greeks -- the committed goldens (tests/golden/ref_misc.npz) and, where the
prebuilt reference library travelled with the snapshot, direct calls on ragged surfaces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
S0, R, Q = 100.0, 0.05, 0.02
DEFAULT = np.array([2.0, 0.04, 0.3, -0.7, 0.04])
# finite differences of prices that agree to ~1e-13 relative: first differences amplify by 1/bump,
# the second difference (gamma, bump 0.1) by 1/bump^2
TOL1 = dict(rtol=1e-6, atol=1e-8)
TOL2 = dict(rtol=1e-4, atol=1e-8)


def _check(got, want):  # columns: delta, gamma, vega, theta, rho
    np.testing.assert_allclose(got[..., [0, 2, 3, 4]], want[..., [0, 2, 3, 4]], **TOL1)
    np.testing.assert_allclose(got[..., 1], want[..., 1], **TOL2)


def test_batched_greeks_match_reference_golden(g_misc):
    import torch

    from pde_b200 import BatchPricer

    K = np.repeat([90.0, 100.0, 110.0], 2)
    ic = np.tile([True, False], 3)
    pr = BatchPricer("refgrid").set_surface(K, 1.0, ic, None, S0=S0, r=R, q=Q)
    X = np.vstack([DEFAULT, [-1.0, 0.04, 0.3, -0.7, 0.04], [2.0, 0.04, 0.3, -0.7, 0.0005]])
    g = pr.greeks(torch.tensor(X, device="cuda:0")).cpu().numpy()
    assert g.shape == (3, 6, 5)
    _check(g[0], g_misc["greeks"][:, 1:])  # golden rows: price, delta, gamma, vega, theta, rho
    assert np.isnan(g[1]).all()  # invalid parameter set
    # v0 - 1e-3 <= 0: the reference's constructor throws for the down-bumped model -> NaN vega, rest fine
    assert np.isnan(g[2, :, 2]).all() and np.isfinite(g[2][:, [0, 1, 3, 4]]).all()
    np.testing.assert_array_equal(g, pr.greeks_host(X))


def test_batched_greeks_ragged_surface_vs_reference(reference):
    from pde_b200 import BatchPricer

    K = np.array([80.0, 95.0, 100.0, 100.0, 105.0, 120.0, 100.0, 100.0])
    T = np.array([0.25, 0.25, 0.5, 0.5, 2.0, 2.0, 0.002, 0.0])  # 0.002 <= 1/365: theta = 0; T = 0: intrinsic
    ic = np.array([True, False, True, False, True, False, True, True])
    X = np.array([DEFAULT, [1.5, 0.09, 0.5, -0.3, 0.06], [4.0, 0.15, 0.8, -0.9, 0.5]])
    pr = BatchPricer("refgrid").set_surface(K, T, ic, None, S0=S0, r=R, q=Q)
    got = pr.greeks_host(X)
    want = np.array([[reference.greeks(x, k, t, S0, R, Q, bool(c))[1:] for k, t, c in zip(K, T, ic)] for x in X])
    _check(got, want)
    assert (got[:, 6:, 3] == 0.0).all()
    # the surface can be replaced: sibling plans follow
    pr.set_surface(K[:3], T[:3], ic[:3], None, S0=S0 * 1.1, r=R, q=Q)
    got = pr.greeks_host(X[:1])
    want = np.array([[reference.greeks(X[0], k, t, S0 * 1.1, R, Q, bool(c))[1:] for k, t, c in zip(K[:3], T[:3], ic[:3])]])
    _check(got, want)


def test_greeks_fft_mode_consistent_with_own_prices():
    """fft mode: the same recipe applied to fft-mode prices (no reference arithmetic exists for it)."""
    from pde_b200 import BatchPricer

    K = np.array([90.0, 100.0, 110.0])
    pr = BatchPricer("fft").set_surface(K, 1.0, True, None, S0=S0, r=R, q=Q)
    g = pr.greeks_host(DEFAULT[None, :])[0]
    es = S0 * 0.001
    up = BatchPricer("fft").set_surface(K, 1.0, True, None, S0=S0 + es, r=R, q=Q).price_host(DEFAULT[None, :])[0]
    dn = BatchPricer("fft").set_surface(K, 1.0, True, None, S0=S0 - es, r=R, q=Q).price_host(DEFAULT[None, :])[0]
    np.testing.assert_array_equal(g[:, 0], (up - dn) / (2.0 * es))
    assert (g[:, 0] > 0).all() and (g[:, 0] < 1).all() and (g[:, 1] > 0).all()
