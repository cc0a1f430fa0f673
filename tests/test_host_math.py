"""CPU checks of the product's device math (pde_b200/csrc/fp64_math.cuh, heston_math.cuh).

The headers are compiled as plain C++ by tests/host_math_check.cpp (MUFU seeds emulated in
float precision) so the FORMULAS -- minimax tables, Newton steps, argument reductions, the
cancellation-free little-trap restatement -- are verified without a GPU.  This harness is test
infrastructure only; nothing under pde_b200/ builds or loads it, and the GPU tests
(tests/test_gpu_parity.py) check the same code compiled by nvcc.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S0, R, Q = 100.0, 0.05, 0.02
LB = np.array([0.1, 0.01, 0.01, -0.99, 0.01])
UB = np.array([10.0, 1.0, 2.0, 0.99, 1.0])
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhm.so")
    flags = ["-mfma"] if "fma" in open("/proc/cpuinfo").read() else []
    subprocess.run(["g++", "-O2", *flags, "-fPIC", "-shared", os.path.join(ROOT, "tests", "host_math_check.cpp"),
                    "-o", out], check=True)
    return C.CDLL(out)


def _ulps(got, exact):
    """|got - exact| in units of 2^-52 |exact| (exact: list of mpmath numbers)."""
    import mpmath as mp

    worst = 0.0
    for g, e in zip(got, exact):
        if e == 0:
            assert g == 0.0
            continue
        worst = max(worst, float(abs(mp.mpf(float(g)) - e) / abs(e) * 2 ** 52))
    return worst


def test_elementary_functions_within_2ulp(hm):
    import mpmath as mp

    mp.mp.dps = 40
    rng = np.random.default_rng(0)
    n = 3000
    f = lambda t: mp.mpf(float(t))
    # exp: normal results; subnormal/underflow/overflow checked separately
    x = np.concatenate([rng.uniform(-700, 709, n), rng.uniform(-1, 1, n), [0.0, 1e-300, -1e-300]])
    y = np.empty_like(x)
    hm.hm_exp(x.size, _p(x), _p(y))
    assert _ulps(y, [mp.exp(f(t)) for t in x]) < 1.5
    x = np.array([-1e4, -1500.0, -746.0, -740.0, 710.0, 1500.0, np.inf, -np.inf])
    y = np.empty_like(x)
    hm.hm_exp(x.size, _p(x), _p(y))
    assert y[0] == 0 and y[1] == 0 and y[2] == 0 and y[7] == 0
    assert abs(y[3] - np.exp(-740.0)) <= 2 * 4.94e-324 and np.isinf(y[4]) and np.isinf(y[5]) and np.isinf(y[6])
    # sincos: the phases of this path (|x| up to ~1e6), quadrant boundaries included
    x = np.concatenate([rng.uniform(-6000, 6000, n), rng.uniform(-1.5e6, 1.5e6, n), rng.uniform(-1, 1, n),
                        np.arange(-200, 200) * (np.pi / 2), [0.0, 1e-300]])
    s, c = np.empty_like(x), np.empty_like(x)
    hm.hm_sincos(x.size, _p(x), _p(s), _p(c))
    exact_s, exact_c = [mp.sin(f(t)) for t in x], [mp.cos(f(t)) for t in x]
    big = np.abs(np.array([float(e) for e in exact_s])) > 1e-3  # relative error only away from zeros
    assert _ulps(s[big], [e for e, b in zip(exact_s, big) if b]) < 2.0
    bigc = np.abs(np.array([float(e) for e in exact_c])) > 1e-3
    assert _ulps(c[bigc], [e for e, b in zip(exact_c, bigc) if b]) < 2.0
    # near the zeros of sin/cos: absolute error against the exact value
    assert max(abs(float(mp.mpf(float(g)) - e)) for g, e in zip(s, exact_s)) < 3e-16
    assert max(abs(float(mp.mpf(float(g)) - e)) for g, e in zip(c, exact_c)) < 3e-16
    # far outside the reducible range: finite, bounded output (the caller's magnitude is 0 there)
    x = np.array([1e9, -3e12, 1e20, np.inf])
    s, c = np.empty_like(x), np.empty_like(x)
    hm.hm_sincos(x.size, _p(x), _p(s), _p(c))
    assert np.isfinite(s).all() and np.isfinite(c).all() and np.abs(s).max() < 2 and np.abs(c).max() < 2
    # log1p
    x = np.concatenate([rng.uniform(-0.999, 10, n), 10 ** rng.uniform(-20, 0, n) * rng.choice([-1, 1], n),
                        10 ** rng.uniform(0, 30, n), [-1 + 1e-15, 1e-300, 0.0]])
    y = np.empty_like(x)
    hm.hm_log1p(x.size, _p(x), _p(y))
    assert _ulps(y, [mp.log1p(f(t)) for t in x]) < 1.5
    x = np.array([-1.0, -2.0, np.inf])
    y = np.empty_like(x)
    hm.hm_log1p(x.size, _p(x), _p(y))
    assert y[0] == -np.inf and np.isnan(y[1]) and y[2] == np.inf
    # atan2: all quadrants, tiny and huge ratios, the 1 + delta shape used by stage B
    yy = np.concatenate([rng.normal(size=n), 10 ** rng.uniform(-12, 0, n) * rng.choice([-1, 1], n),
                         rng.normal(size=200) * 1e8, [0.0, 1.0, -1.0, 0.0]])
    xx = np.concatenate([rng.normal(size=n), 1 + rng.normal(size=n) * 1e-3, rng.normal(size=200), [1.0, 0.0, 0.0, -1.0]])
    a = np.empty_like(xx)
    hm.hm_atan2(xx.size, _p(yy), _p(xx), _p(a))
    assert _ulps(a, [mp.atan2(f(p), f(q)) for p, q in zip(yy, xx)]) < 2.0
    # division / reciprocal / sqrt: compare with IEEE operations
    aa = rng.normal(size=n) * 10 ** rng.uniform(-30, 30, n)
    bb = rng.normal(size=n) * 10 ** rng.uniform(-30, 30, n)
    q = np.empty_like(aa)
    hm.hm_div(n, _p(aa), _p(bb), _p(q))
    assert np.max(np.abs(q - aa / bb) / np.abs(aa / bb)) <= 2 ** -52
    hm.hm_rcp(n, _p(bb), _p(q))
    assert np.max(np.abs(q * bb - 1)) <= 2 ** -51
    pos = np.abs(aa)
    sq, rs = np.empty_like(pos), np.empty_like(pos)
    hm.hm_sqrt(n, _p(pos), _p(sq), _p(rs))
    assert np.max(np.abs(sq - np.sqrt(pos)) / np.sqrt(pos)) <= 2 ** -52
    assert np.max(np.abs(rs * np.sqrt(pos) - 1)) <= 2 ** -50


def _hm_cf(hm, p, v, ui, T):
    p = np.ascontiguousarray(p, dtype=float)
    v = np.ascontiguousarray(v, dtype=float)
    out = np.empty((v.size, 2))
    hm.hm_cf_grid(_p(p), C.c_int(v.size), _p(v), C.c_double(ui), C.c_double(T), C.c_double(S0), C.c_double(R),
                  C.c_double(Q), _p(out))
    return out[:, 0] + 1j * out[:, 1]


def test_cf_formulas_match_reference_golden(hm, g_cf):
    """The real-arithmetic, cancellation-free restatement against the reference's CF values."""
    params, Ts = g_cf["params"], g_cf["T"]
    worst_well, worst_ill = 0.0, 0.0
    for i, p in enumerate(params):
        for m, T in enumerate(Ts):
            for v, want in ((g_cf["v_fft"], g_cf["cf_fft"][i, m]), (g_cf["v_rg"], g_cf["cf_rg"][i, m])):
                got = _hm_cf(hm, p, v, float(g_cf["ui"]), T)
                rel = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-290))
                if p[2] >= 0.02:
                    worst_well = max(worst_well, rel)
                else:
                    worst_ill = max(worst_ill, rel)
    assert worst_well < 1e-10, worst_well
    assert worst_ill < 2e-9, worst_ill  # sigma = 0.01 corners: the reference's own noise (test_conditioning.py)


def test_cf_formulas_are_closer_to_exact_than_the_reference(hm, oracle, g_cf):
    """Against the 80-bit yardstick the product formulas carry ~1e-13, including the small-sigma
    corners where the reference's subtraction xi - d and log(ratio) lose up to 6 digits."""
    v = 0.25 * np.arange(0, 4096, 3)
    worst_ours, worst_ref = 0.0, 0.0
    for p in g_cf["params"]:
        for T in (0.1, 1.0):
            exact = oracle.cf_ld_grid(p, T, v, -1.75, S0, R, Q)
            ours = _hm_cf(hm, p, v, -1.75, T)
            ref = oracle.cf_grid(p, [T], v, -1.75, S0, R, Q)[0, 0]
            ok = np.abs(exact) > 1e-280
            worst_ours = max(worst_ours, np.max(np.abs(ours - exact)[ok] / np.abs(exact)[ok]))
            worst_ref = max(worst_ref, np.max(np.abs(ref - exact)[ok] / np.abs(exact)[ok]))
    assert worst_ours < 5e-12, worst_ours
    assert worst_ref > 10 * worst_ours  # documents why parity is asserted at 4x tolerance for sigma < 0.02


def test_cf_general_u(hm, g_cf):
    for i, p in enumerate(g_cf["params"][:8]):
        for u, want in zip(g_cf["u_gen"], g_cf["cf_gen"][i]):
            got = _hm_cf(hm, p, [u.real], u.imag, 0.7)[0]
            assert abs(got - want) / abs(want) < 1e-10


def test_tail_skip_margin(hm):
    """kernels.cuh skips stage B/F of the kappa'/sigma'/rho' slices where the BASE set's log|phi| is
    below -800.  That is exact iff the perturbed exponent stays below -745.14 there.  Over Sobol sets
    and all box corners the finite-difference perturbation (1.5e-8 relative) moves the exponent by a
    relative 3e-6 at most (condition number < 200), four orders of magnitude inside the margin."""
    import sys

    sys.path.insert(0, ROOT)
    from bench import fd_variants
    from scipy.stats import qmc

    corners = [np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)]
    X = np.vstack([LB + (UB - LB) * qmc.Sobol(d=5, seed=7).random(96), corners])
    v = 0.25 * np.arange(4096.0)

    def exponent(p, T):
        p = np.ascontiguousarray(p, dtype=float)
        out = np.empty(v.size)
        hm.hm_cf_exponent(_p(p), C.c_int(v.size), _p(v), C.c_double(-1.75), C.c_double(T), C.c_double(S0),
                          C.c_double(R), C.c_double(Q), _p(out))
        return out

    worst_rel, flagged = 0.0, 0
    for x in X:
        V = fd_variants(x[None, :])
        for T in (0.1, 1.0, 2.0):
            e0 = exponent(V[0], T)
            dead = e0 < -800.0
            flagged += int(dead.sum())
            for k in (1, 3, 4):  # kappa', sigma', rho'
                ek = exponent(V[k], T)
                if dead.any():
                    assert (ek[dead] < -746.0).all()
                    worst_rel = max(worst_rel, float(np.max(np.abs(ek - e0)[dead] / np.abs(e0[dead]))))
    assert flagged > 0.2 * len(X) * 3 * v.size  # the regime is common: > 20 % of all grid points
    assert worst_rel < 1e-4  # vs the 6.7 % (54/800) it would take to break the skip


def test_tail_bound_is_rigorous_and_useful(hm):
    """heston_math.cuh tail_dead(): decides from stage A alone (no cexp, no clog) that Re(exponent) < -746,
    i.e. that stage F would return exactly 0.  Implication must hold on every grid point of every set
    (Sobol + all box corners, FD-perturbed variants included); coverage shows the bound is worth having."""
    import sys

    sys.path.insert(0, ROOT)
    from bench import fd_variants
    from scipy.stats import qmc

    corners = [np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)]
    X = np.vstack([LB + (UB - LB) * qmc.Sobol(d=5, seed=11).random(128), corners])
    v = 0.25 * np.arange(4096.0)
    n_dead = n_caught = 0
    for x in X:
        for p in fd_variants(x[None, :])[[0, 1, 3, 4]]:  # base, kappa', sigma', rho'
            p = np.ascontiguousarray(p, dtype=float)
            for T in (0.1, 0.55, 1.0, 2.0):
                er = np.empty(v.size)
                dead = np.zeros(v.size, dtype=np.int32)
                hm.hm_tail_bound(_p(p), C.c_int(v.size), _p(v), C.c_double(-1.75), C.c_double(T), C.c_double(S0),
                                 C.c_double(R), C.c_double(Q), _p(er), dead.ctypes.data_as(C.POINTER(C.c_int)))
                flagged = dead.astype(bool)
                assert (er[flagged] < -749.0).all(), (p, T, er[flagged].max())  # 3 of the 4 units of slack unused
                n_dead += int((er < -746.0).sum())
                n_caught += int(flagged.sum())
    assert n_dead > 0 and n_caught > 0.8 * n_dead, (n_caught, n_dead)


def test_asymptotic_stage_b_is_the_full_one_beyond_dT_45(hm):
    """heston_math.cuh stage_b_asym(): for Re(d) T > 45 the exponent of phi from the two-FMA form must equal the
    one from the full stage B (cexp + clog) to rounding -- wherever phi has not underflowed, the difference has
    to stay far below one ulp of exp (1.1e-16 relative in phi = 1.1e-16 absolute in the exponent)."""
    from scipy.stats import qmc

    corners = [np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)]
    X = np.vstack([LB + (UB - LB) * qmc.Sobol(d=5, seed=13).random(128), corners])
    v = 0.25 * np.arange(4096.0)
    n_fast = n_alive = 0
    worst = 0.0
    for p in X:
        p = np.ascontiguousarray(p, dtype=float)
        for T in (0.1, 0.55, 1.0, 2.0):
            o = [np.empty(v.size) for _ in range(5)]
            hm.hm_stage_b_asym(_p(p), C.c_int(v.size), _p(v), C.c_double(-1.75), C.c_double(T), C.c_double(S0),
                               C.c_double(R), C.c_double(Q), *[_p(x) for x in o])
            er_f, ei_f, er_a, ei_a, dT = o
            alive = er_f > -746.0
            fast = (dT > 45.0) & alive
            n_alive += int(alive.sum())
            n_fast += int(fast.sum())
            if fast.any():
                # rounding of the two evaluations: a few ulp of the exponent's terms (|er| up to 746, kts up to 1e5)
                scale = 1.0 + np.abs(er_f[fast]) + np.abs(ei_f[fast])
                worst = max(worst, float(np.max(np.maximum(np.abs(er_a - er_f)[fast], np.abs(ei_a - ei_f)[fast]) / scale)))
    assert n_fast > 0.5 * n_alive  # most of the live integrand is in the asymptotic regime
    assert worst < 2e-15, worst


def test_intermediate_stage_b_matches_the_full_one(hm):
    """heston_math.cuh stage_b_mid() (series for log(1 - g e) and 1/(1 - g e); prepared for the next round, not yet
    called by the kernels): wherever its premise |g e| <= 2^-17 holds and phi has not underflowed, the exponent of
    phi must equal the one from the full stage B to rounding."""
    from scipy.stats import qmc

    corners = [np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)]
    X = np.vstack([LB + (UB - LB) * qmc.Sobol(d=5, seed=17).random(128), corners])
    v = 0.25 * np.arange(4096.0)
    n_mid = 0
    worst = 0.0
    for p in X:
        p = np.ascontiguousarray(p, dtype=float)
        for T in (0.1, 0.55, 1.0, 2.0):
            o = [np.empty(v.size) for _ in range(4)]
            flag = np.zeros(v.size, dtype=np.int32)
            hm.hm_stage_b_mid(_p(p), C.c_int(v.size), _p(v), C.c_double(-1.75), C.c_double(T), C.c_double(S0),
                              C.c_double(R), C.c_double(Q), *[_p(x) for x in o], flag.ctypes.data_as(C.POINTER(C.c_int)))
            er_f, ei_f, er_m, ei_m = o
            use = flag.astype(bool) & (er_f > -746.0)
            n_mid += int(use.sum())
            if use.any():
                scale = 1.0 + np.abs(er_f[use]) + np.abs(ei_f[use])
                worst = max(worst, float(np.max(np.maximum(np.abs(er_m - er_f)[use], np.abs(ei_m - ei_f)[use]) / scale)))
    assert n_mid > 100000
    assert worst < 2e-15, worst



def test_live_prefix_bound_is_rigorous_and_tight(hm):
    """prefix_bound.cuh: the direct-sum kernel evaluates the points j < J of a slice only.  J comes from a bound on
    log|phi| over blocks of the grid that needs no characteristic function; every grid point j >= J must lie below
    the significance cut (Sobol sets, all box corners, FD-perturbed classes, two cuts, N = 4096 and 16384), and the
    bound's prefix must not be much longer than the true one (it is paid for in stage B / F work)."""
    import sys

    sys.path.insert(0, ROOT)
    from bench import fd_variants
    from scipy.stats import qmc

    corners = [np.where([(m >> b) & 1 for b in range(5)], UB, LB) for m in range(32)]
    sob = LB + (UB - LB) * qmc.Sobol(d=5, seed=3).random(128)
    Ts = np.ascontiguousarray(np.linspace(0.1, 1.0, 8).tolist() + [0.02, 2.0, 5.0])
    ip = C.POINTER(C.c_int)
    # (the corners with sigma = 0.01 are the bound's weak spot: never asymptotic, |Im d| ~ Re d, prefix = whole grid --
    # such sets are routed to the transform kernel; tightness is asserted on the interior of the box)
    for N, cut, X, tight, eta, alpha in ((4096, -51.3, sob, 1.2, 0.25, 0.75), (4096, -51.3, np.array(corners), None, 0.25, 0.75),
                                         (4096, -36.0, sob[:32], 1.2, 0.25, 0.75), (4096, -120.0, sob[:32], 1.2, 0.25, 0.75),
                                         (16384, -52.0, sob[:24], 1.2, 0.25, 0.75), (4096, -45.0, sob[:32], 1.25, 0.125, 1.25),
                                         (512, -48.0, sob[:32], None, 0.5, 1.0)):
        v = eta * np.arange(float(N))
        tot_true = tot_bound = 0
        for x in X:
            for p in fd_variants(x[None, :])[[0, 1, 3, 4]]:  # base, kappa', sigma', rho'
                p = np.ascontiguousarray(p, dtype=float)
                J = np.zeros(Ts.size, dtype=np.int32)
                nb = hm.hm_prefix_J(_p(p), C.c_int(Ts.size), _p(Ts), C.c_int(N), C.c_double(eta), C.c_double(alpha),
                                    C.c_double(S0), C.c_double(R), C.c_double(Q), C.c_double(cut), J.ctypes.data_as(ip))
                assert 0 < nb <= 160
                for t, T in enumerate(Ts):
                    er = np.empty(v.size)
                    hm.hm_cf_exponent(_p(p), C.c_int(v.size), _p(v), C.c_double(-(alpha + 1.0)), C.c_double(T), C.c_double(S0),
                                      C.c_double(R), C.c_double(Q), _p(er))
                    assert 1 <= J[t] <= N
                    tail = er[J[t]:]
                    assert not (tail >= cut).any(), (N, cut, p, T, J[t], np.nanmax(tail))
                    live = np.nonzero(~(er < cut))[0]  # NaN counts as live
                    tot_true += int(live.max()) + 1 if live.size else 1
                    tot_bound += int(J[t])
        assert tight is None or tot_bound < tight * tot_true, (N, cut, tot_bound, tot_true)
