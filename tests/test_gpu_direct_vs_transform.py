"""Differential test of the two Carr-Madan job kernels on random surfaces: the direct-sum kernel
(live prefix + conjugate-pair sums; HB_DIRECT_THR=0 sends every set to it), the transform kernel (HB_DIRECT=0) and the
routed default must agree far inside the parity tolerance on prices, losses and residuals, and within the finite-difference
amplification of that tolerance on the Jacobian -- ragged maturities, calls and puts, expired / invalid options, strikes
off the grid, several grids and dampings, invalid and corner parameter sets, small (split) and large (persistent) batches."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

S0, R, Q = 100.0, 0.05, 0.02
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])


def _surface(rng, n_grid, eta):
    n_mat = int(rng.integers(1, 12))
    Ks, Ts, calls = [], [], []
    lam, b = 2 * np.pi / (n_grid * eta), np.pi / eta
    k_hi = min(np.log(400.0), -b + lam * (n_grid - 2))
    for _ in range(n_mat):
        T = float(rng.choice([0.02, 0.1, 0.25, 0.5, 1.0, 2.0, 3.0])) * float(rng.uniform(0.8, 1.2))
        n = int(rng.integers(1, 40))
        centre, width = rng.uniform(70, 140), rng.uniform(0.05, 0.6)
        K = np.exp(np.clip(np.log(centre) + width * rng.uniform(-1, 1, n), np.log(20.0), k_hi))
        Ks.append(K)
        Ts.append(np.full(n, T))
        calls.append(rng.random(n) < 0.6)
    K, T, ic = np.concatenate(Ks), np.concatenate(Ts), np.concatenate(calls)
    if rng.random() < 0.5:  # an expired option and an invalid strike (intrinsic / NaN rows)
        K = np.concatenate([K, [90.0, -5.0]])
        T = np.concatenate([T, [0.0, 0.5]])
        ic = np.concatenate([ic, [True, False]])
    perm = rng.permutation(K.size)
    return K[perm], T[perm], ic[perm]


def _params(rng, n):
    from scipy.stats import qmc

    x = LB + (UB - LB) * qmc.Sobol(d=5, seed=int(rng.integers(1 << 30))).random(n)
    corners = np.array([np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in rng.integers(0, 32, 6)])
    bad = np.array([[-1.0, 0.04, 0.3, -0.7, 0.04], [2.0, 0.04, 0.3, 1.5, 0.04]])
    return np.vstack([x, corners, bad])


@pytest.mark.parametrize("seed,n_grid,eta,alpha,n_sets", [
    (0, 4096, 0.25, 0.75, 40), (1, 4096, 0.25, 0.75, 700), (2, 4096, 0.125, 1.25, 60), (3, 512, 0.5, 0.75, 50),
    (4, 8192, 0.25, 0.75, 30), (5, 4096, 0.25, 0.75, 350), (6, 16384, 0.0625, 0.75, 20), (7, 4096, 0.5, 1.0, 90),
])
def test_direct_and_transform_kernels_agree(monkeypatch, seed, n_grid, eta, alpha, n_sets):
    import torch

    from pde_b200 import BatchPricer

    rng = np.random.default_rng(seed)
    K, T, ic = _surface(rng, n_grid, eta)
    xs = _params(rng, n_sets)
    X = torch.tensor(xs, device="cuda:0")
    pr = BatchPricer("fft", n_grid=n_grid, eta=eta, alpha=alpha).set_surface(K, T, ic, None, S0=S0, r=R, q=Q)
    truth = np.array([[2.0, 0.04, 0.3, -0.7, 0.04]])
    mk = pr.price(torch.tensor(truth, device="cuda:0")).cpu().numpy()[0]
    mk = np.where(np.isfinite(mk), np.maximum(mk * (1 + 0.01 * rng.normal(size=mk.size)), 0.01), 1.0)
    pr.set_surface(K, T, ic, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)

    def run():
        p = pr.price(X).cpu().numpy()
        l = pr.objective(X).cpu().numpy()
        r, j = (t.cpu().numpy() for t in pr.jacobian(X))
        neq = pr.normal_equations(X).cpu().numpy()
        return p, l, r, j, neq

    monkeypatch.setenv("HB_DIRECT", "0")
    ref = run()  # transform kernel
    monkeypatch.delenv("HB_DIRECT")
    for thr in (None, "0"):  # routed default; every set on the direct-sum kernel
        if thr is not None:
            monkeypatch.setenv("HB_DIRECT_THR", thr)
        got = run()
        p0, l0, r0, j0, n0 = ref
        p1, l1, r1, j1, n1 = got
        assert np.array_equal(np.isnan(p0), np.isnan(p1))
        ok = np.isfinite(p0)
        tol_p = 1e-10 * np.abs(p0) + 1e-12
        assert (np.abs(p1 - p0)[ok] <= 0.05 * tol_p[ok]).all(), (thr, float(np.max((np.abs(p1 - p0) / tol_p)[ok])))
        assert np.array_equal(l0 == 1e10, l1 == 1e10) and np.array_equal(np.isfinite(l0), np.isfinite(l1))
        fin = np.isfinite(l0) & (l0 != 1e10)
        np.testing.assert_allclose(l1[fin], l0[fin], rtol=1e-10)
        assert np.array_equal(np.isfinite(r0), np.isfinite(r1))
        okr = np.isfinite(r0)
        tol_r = (0.05 * tol_p / mk[None, :])
        assert (np.abs(r1 - r0)[okr] <= tol_r[okr] + 1e-15).all()
        # Jacobian: a price discrepancy dp moves an entry by 2 dp / (|dx| market); dx >= 1.49e-8 max(1, |x|) (SciPy rule)
        dx = 1.4901161193847656e-08 * np.maximum(1.0, np.abs(xs))  # [P, 5] (bound-limited steps are larger: looser bound not needed)
        okj = np.isfinite(j0) & np.isfinite(j1)
        assert np.array_equal(np.isfinite(j0), np.isfinite(j1))
        bound = 2.0 * (0.05 * tol_p / mk[None, :])[:, :, None] / dx[:, None, :] + 1e-9 * np.abs(j0)
        assert (np.abs(j1 - j0)[okj] <= bound[okj]).all(), (thr, float(np.max((np.abs(j1 - j0) / bound)[okj])))
        np.testing.assert_allclose(n1[fin, :2], n0[fin, :2], rtol=1e-10)
    pr.close()


def test_maturity_entirely_off_the_grid(monkeypatch):
    """A maturity none of whose strikes lies on the log-strike grid has no bins at all (a wave piece without chunks in
    the direct-sum kernel): its options are NaN, the others are priced as usual, on both kernels and both launch paths."""
    import torch

    from pde_b200 import BatchPricer

    n_grid, eta = 512, 0.5  # grid ends at ln K = pi/eta = 6.28: K = 535
    K = np.array([90.0, 100.0, 110.0, 600.0, 700.0, 900.0, 95.0, 105.0])
    T = np.array([0.5, 0.5, 0.5, 1.0, 1.0, 1.0, 2.0, 2.0])
    pr = BatchPricer("fft", n_grid=n_grid, eta=eta).set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    for n_sets in (5, 700):
        xs = _params(np.random.default_rng(n_sets), n_sets)
        X = torch.tensor(xs, device="cuda:0")
        got = pr.price(X).cpu().numpy()
        monkeypatch.setenv("HB_DIRECT", "0")
        ref = pr.price(X).cpu().numpy()
        monkeypatch.delenv("HB_DIRECT")
        valid = (xs[:, 0] > 0) & (np.abs(xs[:, 3]) < 1)
        assert np.isnan(got[:, 3:6]).all() and np.isnan(got[~valid]).all()
        assert np.isfinite(got[valid][:, [0, 1, 2, 6, 7]]).all()
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = np.isfinite(ref)
        assert (np.abs(got - ref)[ok] <= 0.05 * (1e-10 * np.abs(ref[ok]) + 1e-12)).all()
    pr.close()


@pytest.mark.parametrize("n_sets", [1, 2, 295, 296, 297, 591, 592, 593, 1183, 1185])
def test_batch_sizes_around_the_launch_path_boundaries(monkeypatch, n_sets):
    """Batch sizes on both sides of the split / persistent boundaries of the six-variant (2 x SMs x 2 CTAs = 592 on a
    148-SM part) and one-variant (1184) direct-sum kernels and of the transform kernel (296): every row must equal the
    row the same set gets in a large batch, bit for bit, and the two kernels must agree."""
    import torch

    from pde_b200 import BatchPricer

    rng = np.random.default_rng(7)
    K, T = np.tile(np.linspace(85, 115, 9), 5), np.repeat(np.array([0.1, 0.3, 0.6, 1.0, 1.7]), 9)
    pr = BatchPricer("fft").set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    mk = pr.price(torch.tensor(np.array([[2.0, 0.04, 0.3, -0.7, 0.04]]), device="cuda:0")).cpu().numpy()[0]
    pr.set_surface(K, T, True, np.maximum(mk * (1 + 0.01 * rng.normal(size=mk.size)), 0.01), S0=S0, r=R, q=Q)
    big = _params(np.random.default_rng(11), 1400)
    Xb = torch.tensor(big, device="cuda:0")
    ref_loss, ref_neq = pr.objective(Xb).cpu().numpy(), pr.normal_equations(Xb).cpu().numpy()
    X = Xb[:n_sets].contiguous()
    loss, neq = pr.objective(X).cpu().numpy(), pr.normal_equations(X).cpu().numpy()
    assert np.array_equal(loss, ref_loss[:n_sets], equal_nan=True)
    assert np.array_equal(neq, ref_neq[:n_sets], equal_nan=True)
    monkeypatch.setenv("HB_DIRECT", "0")
    loss_t = pr.objective(X).cpu().numpy()
    fin = np.isfinite(loss) & (loss != 1e10)
    assert np.array_equal(loss == 1e10, loss_t == 1e10)
    np.testing.assert_allclose(loss_t[fin], loss[fin], rtol=1e-10)
    pr.close()


def test_bound_limited_finite_difference_steps(monkeypatch):
    """A degenerate box (a parameter on a bound of a narrow box) makes SciPy's step rule take steps as large as the box:
    theta' / v0' then differ from the base by far more than 1e-8 and the class-0 expansion must fall back to the full
    exponentials; both kernels against each other, and the Jacobian against a plain forward difference of prices."""
    import torch

    from pde_b200 import BatchPricer

    K, T = np.tile(np.linspace(85, 115, 7), 3), np.repeat(np.array([0.25, 0.75, 1.5]), 7)
    lb = np.array([1.0, 0.04, 0.3, -0.7, 0.04])
    ub = np.array([3.0, 0.04 + 1e-9, 0.3 + 1e-3, -0.7 + 1e-12, 0.05])  # theta, rho: narrower than the default step
    xs = np.array([[2.0, 0.04, 0.3, -0.7, 0.04], [3.0, 0.04 + 1e-9, 0.3 + 1e-3, -0.7 + 1e-12, 0.05],
                   [1.0, 0.04 + 5e-10, 0.3005, -0.7, 0.045]])
    X = torch.tensor(xs, device="cuda:0")
    pr = BatchPricer("fft").set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    mk = pr.price(X[:1]).cpu().numpy()[0] * 1.01
    pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(lb, ub)
    res, jac = (t.cpu().numpy() for t in pr.jacobian(X))
    monkeypatch.setenv("HB_DIRECT", "0")
    res_t, jac_t = (t.cpu().numpy() for t in pr.jacobian(X))
    monkeypatch.delenv("HB_DIRECT")
    p0 = pr.price(X).cpu().numpy()
    tol_r = 0.05 * (1e-10 * np.abs(p0) + 1e-12) / mk[None, :]
    assert (np.abs(res - res_t) <= tol_r).all()
    assert np.isfinite(jac).all() and np.isfinite(jac_t).all()
    scale = np.abs(jac_t).max(axis=1, keepdims=True) + 1e-300
    # large steps: the finite difference is well conditioned, the kernels agree to rounding relative to the column scale;
    # tiny steps (1e-12 on rho): bounded by the amplified price tolerance
    bound = 2.0 * tol_r[:, :, None] / 1e-12 + 1e-9 * scale
    assert (np.abs(jac - jac_t) <= bound).all()
    pr.close()
