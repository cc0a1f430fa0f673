"""GPU: fine grids (BASELINE.json config 5: N = 16384, 128 maturities x 200 strikes).

A 16384-point complex128 slice is 256 KiB and does not fit one CTA's shared memory; the kernel
transforms it as R = N/4096 decimated on-chip sub-transforms and accumulates
X[m] = sum_ph W_N^{ph m} Y_ph[m mod 4096] at the quoted bins only.  Checked against
the oracle's plain radix-2 FFT over the full grid.
"""
import numpy as np
import pytest

from oracle.oracle import MODE_FFT

pytestmark = pytest.mark.gpu

S0, R, Q = 100.0, 0.05, 0.02
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])


def _sets(n, seed):
    from scipy.stats import qmc

    return np.vstack([[2.0, 0.04, 0.3, -0.7, 0.04], LB + (UB - LB) * qmc.Sobol(d=5, seed=seed).random(n)])


def _viol(got, want):
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    return float(np.max(np.abs(got[ok] - want[ok]) / (1e-10 * np.abs(want[ok]) + 1e-12)))


@pytest.mark.parametrize("n_grid,eta", [(16384, 0.25), (16384, 0.0625), (8192, 0.125), (32768, 0.05)])
def test_fine_grid_prices_match_oracle(oracle, monkeypatch, n_grid, eta):
    import torch

    from pde_b200 import BatchPricer

    K = np.tile(np.linspace(60, 160, 200), 6)
    T = np.repeat(np.array([1 / 12, 0.25, 0.5, 1.0, 1.5, 2.0]), 200)
    ic = np.arange(K.size) % 4 != 0
    params = _sets(7, seed=n_grid)
    params = params[params[:, 2] >= 0.02]  # keep the sigma ~ 0.01 corner for test_gpu_parity (4x tolerance there)
    pr = BatchPricer("fft", n_grid=n_grid, eta=eta).set_surface(K, T, ic, None, S0=S0, r=R, q=Q)
    want = oracle.price_batch(MODE_FFT, params, K, T, ic, S0, R, Q, N=n_grid, eta=eta)
    # default dispatch (the one-variant direct-sum kernel where a maturity has <= 256 conjugate pairs: every case but
    # the first) and the transform kernel with decimation in time forced (HB_DIRECT=0)
    for force_transform in (False, True):
        if force_transform:
            monkeypatch.setenv("HB_DIRECT", "0")
        got = pr.price(torch.tensor(params, device="cuda:0")).cpu().numpy()
        for i in range(len(params)):
            assert _viol(got[i], want[i]) <= 1.0, (force_transform, i, params[i], _viol(got[i], want[i]))


def test_config5_shape_objective_and_normal_equations(oracle):
    """128 maturities x 200 strikes, N = 16384: loss and normal-equation blocks vs the oracle at a
    handful of candidates; both launch paths (split / persistent) agree bit for bit."""
    import torch

    from pde_b200 import BatchPricer

    n_grid, eta = 16384, 0.25
    K = np.tile(np.linspace(80, 120, 200), 128)
    T = np.repeat(np.linspace(1 / 12, 2.0, 128), 200)
    truth = np.array([1.5, 0.06, 0.5, -0.6, 0.05])
    pr = BatchPricer("fft", n_grid=n_grid, eta=eta).set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    mk = pr.price(torch.tensor(truth[None, :], device="cuda:0")).cpu().numpy()[0]
    want_mk = oracle.price_batch(MODE_FFT, [truth], K, T, True, S0, R, Q, N=n_grid, eta=eta)[0]
    assert _viol(mk, want_mk) <= 1.0
    mk = np.maximum(want_mk * (1 + 0.001 * np.random.default_rng(0).normal(size=mk.size)), 0.01)
    pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q)
    xs = np.vstack([truth, truth * 1.05, [3.0, 0.1, 0.8, -0.3, 0.2]])
    X = torch.tensor(xs, device="cuda:0")
    loss = pr.objective(X).cpu().numpy()
    want = oracle.objective_batch(MODE_FFT, xs, K, T, True, mk, S0, R, Q, N=n_grid, eta=eta)
    np.testing.assert_allclose(loss, want, rtol=1e-8)
    neq = pr.normal_equations(X).cpu().numpy()
    np.testing.assert_allclose(neq[:, 0], want, rtol=1e-8)
    big = torch.tensor(np.vstack([xs, np.tile(truth * 0.9, (400, 1))]), device="cuda:0")  # persistent path
    assert np.array_equal(pr.objective(big).cpu().numpy()[:3], loss)
