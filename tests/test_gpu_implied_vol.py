"""GPU: batched implied-volatility epilogue against the compiled reference's
HestonModel::implied_volatility (tests/golden/ref_misc.npz) and against the scalar drop-in."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
S0, R, Q = 100.0, 0.05, 0.02
DEFAULT = np.array([2.0, 0.04, 0.3, -0.7, 0.04])


def test_batched_implied_vol_matches_reference(g_misc):
    import torch

    from pde_b200 import BatchPricer
    from pde_b200.cpp import quant_cpp
    from pde_b200.models import HestonModel

    K = np.tile([90.0, 100.0, 110.0], 2)
    T = np.repeat([0.25, 1.0], 3)
    pr = BatchPricer("refgrid").set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    X = np.vstack([DEFAULT, [1.5, 0.09, 0.5, -0.3, 0.06], [-1.0, 0.04, 0.3, -0.7, 0.04]])
    iv = pr.implied_vol(torch.tensor(X, device="cuda:0")).cpu().numpy()
    np.testing.assert_allclose(iv[0].reshape(2, 3), g_misc["iv"], rtol=1e-7)  # Newton stops at |diff| < 1e-8
    assert np.isnan(iv[2]).all()  # invalid parameter set
    np.testing.assert_array_equal(iv, pr.implied_vol_host(X))
    # same recipe as the scalar drop-in (heston.cpp:311-349), puts included, T = 0 -> 0
    Kp = np.array([95.0, 100.0, 105.0, 100.0])
    Tp = np.array([0.5, 0.5, 2.0, 0.0])
    ic = np.array([False, True, False, True])
    pr.set_surface(Kp, Tp, ic, None, S0=S0, r=R, q=Q)
    got = pr.implied_vol_host(X[:2])
    for i, x in enumerate(X[:2]):
        m = quant_cpp.heston.HestonModel(quant_cpp.heston.HestonParameters(*x))
        want = [m.implied_volatility(k, t, S0, R, Q, bool(c)) for k, t, c in zip(Kp, Tp, ic)]
        np.testing.assert_allclose(got[i], want, rtol=1e-9, atol=1e-12)
    assert got[0, 3] == 0.0
    # the wrapper's surface (strikes x maturities), one launch instead of a double loop
    surf = HestonModel(*DEFAULT).implied_volatility_surface([90.0, 100.0, 110.0], [0.25, 1.0], S0, R, Q)
    np.testing.assert_allclose(surf, g_misc["iv"].T, rtol=1e-7)
