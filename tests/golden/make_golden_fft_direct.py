"""Independent pin of the `fft` mode's transform / interpolation stage -> tests/golden/fft_direct.npz.

    python tests/golden/make_golden_fft_direct.py        (build container only: needs oracle/_ref)

The reference has no FFT pricer, so `oracle/heston_oracle.c`'s ORC_MODE_FFT could only be
checked against itself for everything after the characteristic function.  This generator shares NO code
with it:

  * phi_j comes from the reference's own HestonModel::characteristic_function, compiled unmodified
    (oracle/_ref/libheston_ref.so, src/cpp/models/heston.cpp:74-92), on the damped grid
    u_j = eta j - (alpha + 1) i;
  * everything after it is evaluated here in 50-digit arithmetic (mpmath) straight from the written
    specification -- docs/models/heston-model.md:89-106 (N, eta, "FFT ... interpolate to desired
    strikes"), src/cpp/models/heston.cpp:109-149 (psi's denominator, e^{-alpha k}/pi e^{-rT} scaling, clamp,
    put by parity), SURVEY.md Appendix B for the step order -- with NO fast transform: the two grid values
    bracketing a strike are the plain O(N) sums
        X_m = sum_j e^{i b v_j} psi_j w_j e^{-2 pi i j m / N},        m = floor((ln K + b)/lambda), m + 1
    (e^{i b v_j} is exponentiated as written, not replaced by (-1)^j; the twiddle is exp of the exactly
    reduced angle), then linear interpolation in ln K, max(., 0), parity, rounded once to double.

What this pins: Simpson weights, the sign and normalisation of the transform, the grid origin b = pi/eta,
lambda = 2 pi/(N eta), the bin/fraction of each strike, the scale e^{-alpha k_m}/pi, the discount, the clamp
and the parity formula -- for N = 4096 and 16384 (eta 0.25 / 0.0625) and N = 512, calls and puts,
default / Sobol / box-corner parameter sets (slow-decay and fast-decay corners included).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import Reference, build  # noqa: E402

S0, R_, Q_ = 100.0, 0.05, 0.02
ALPHA = 0.75
DEFAULT = [2.0, 0.04, 0.3, -0.7, 0.04]
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])  # heston_calibrator.py:201-207
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])


def direct_prices(ref, p, T, N, eta, strikes, is_call):
    """Prices of `strikes` for one (parameter set, maturity) by the written specification, 50 digits."""
    import mpmath as mp

    mp.mp.dps = 50
    f = lambda x: mp.mpf(float(x))
    alpha, r, q, s0 = f(ALPHA), f(R_), f(Q_), f(S0)
    eta_, Tm = f(eta), f(T)
    v = eta * np.arange(N)  # the doubles the pricers feed to the CF
    phi = ref.cf_grid(p, v, -(ALPHA + 1.0), T, S0, R_, Q_)  # reference's double-precision CF
    lam = 2 * mp.pi / (N * eta_)
    b = mp.pi / eta_
    disc = mp.e ** (-r * Tm)
    x = []
    for j in range(N):
        vj = f(v[j])
        den = mp.mpc(alpha * alpha + alpha - vj * vj, (2 * alpha + 1) * vj)  # heston.cpp:117
        psi = mp.mpc(f(phi[j].real), f(phi[j].imag)) / den
        w = (eta_ / 3) * (3 + (-1) ** (j + 1) - (1 if j == 0 else 0))  # Simpson, Carr-Madan (1999)
        x.append(mp.e ** (mp.mpc(0, 1) * b * vj) * psi * w)
    tw = [mp.e ** (mp.mpc(0, -2) * mp.pi * k / N) for k in range(N)]
    out = []
    for K, call in zip(strikes, is_call):
        k = mp.log(f(K))
        m = int(mp.floor((k + b) / lam))
        assert 0 <= m < N - 1
        C = []
        for mm in (m, m + 1):
            X = mp.mpc(0)
            for j in range(N):
                X += x[j] * tw[(j * mm) % N]
            km = -b + lam * mm
            C.append(mp.e ** (-alpha * km) / mp.pi * disc * X.real)  # heston.cpp:139
        km = -b + lam * m
        c = C[0] + (C[1] - C[0]) * (k - km) / lam
        c = max(c, mp.mpf(0))  # heston.cpp:142
        if not call:
            c = max(c - s0 * mp.e ** (-q * Tm) + f(K) * disc, mp.mpf(0))  # heston.cpp:148-149
        out.append(float(c))
    return out


def cases():
    from scipy.stats import qmc

    sob = LB + (UB - LB) * qmc.Sobol(d=5, seed=42).random(4)
    slow = [0.1, 0.01, 2.0, -0.99, 0.01]   # no decay up to v = 1023.75
    fast = [10.0, 1.0, 0.01, 0.99, 1.0]    # phi underflows by v = 100
    mid1 = [1.5, 0.09, 0.5, -0.3, 0.06]
    mid2 = [4.0, 0.15, 0.8, -0.9, 0.5]
    K5 = [80.0, 95.3, 100.0, 111.7, 120.0]
    rows = []
    # (params, T, N, eta, strikes, is_call)
    for p in [DEFAULT, slow, fast, mid1, mid2, *sob]:
        for T in (0.1, 1.0):
            rows.append((p, T, 4096, 0.25, K5, [True, True, False, True, False]))
    rows.append((DEFAULT, 0.55, 4096, 0.25, [60.0, 150.0, 200.0, 99.999], [True, True, True, False]))
    rows.append((DEFAULT, 0.5, 512, 0.5, [85.0, 100.0, 118.0], [True, False, True]))
    for p in [DEFAULT, mid2]:
        rows.append((p, 0.5, 16384, 0.25, [90.0, 104.2], [True, False]))
        rows.append((p, 0.25, 16384, 0.0625, [90.0, 104.2], [True, False]))
    return rows


def main():
    build(ref=True)
    ref = Reference()
    P, T, N, E, K, C, V = [], [], [], [], [], [], []
    for (p, t, n, eta, strikes, calls) in cases():
        pr = direct_prices(ref, p, t, n, eta, strikes, calls)
        for k, c, val in zip(strikes, calls, pr):
            P.append(p), T.append(t), N.append(n), E.append(eta), K.append(k), C.append(c), V.append(val)
        print(np.round(p, 4), t, n, eta, pr, flush=True)
    np.savez(os.path.join(HERE, "fft_direct.npz"), params=np.array(P, dtype=np.float64), T=np.array(T), N=np.array(N),
             eta=np.array(E), K=np.array(K), is_call=np.array(C, dtype=bool), price=np.array(V),
             spot_rate_div=np.array([S0, R_, Q_]), alpha=np.array(ALPHA))
    print("wrote fft_direct.npz:", len(V), "prices")


if __name__ == "__main__":
    main()
