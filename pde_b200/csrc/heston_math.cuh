// heston_math.cuh -- FP64 real-arithmetic Heston characteristic function for sm_100a.
//
// Restates /root/reference src/cpp/models/heston.cpp:37-92
// (compute_cf_intermediates + characteristic_function, the "little trap" form)
// without std::complex, split into three stages so that work shared between the
// slices of one calibration job is done once (SURVEY.md Appendix A "reuse"):
//
//   stage A  (kappa,sigma,rho ; u)        -> d, g, xi-d, g/(1-g)      T-independent
//   stage B  (stage A ; T)                -> bracket B, Dq            theta/v0-independent
//   stage F  (stage B ; theta, v0, T, ...) -> phi                      one cexp
//
//   C = (kappa*theta/sigma^2) * B,  B  = (xi-d) T - 2 log((1-g e)/(1-g))      heston.cpp:63-65
//   D = Dq / sigma^2,               Dq = (xi-d) (1-e)/(1-g e)                 heston.cpp:69
//   phi = exp(C + D v0 + i u (ln S0 + (r-q) T))                                heston.cpp:87-91
//
// Numerics.  The reference evaluates xi-d by subtraction and log(ratio) of a
// ratio close to 1; both cancel when sigma is small (kappa*theta/sigma^2 up to
// 1e5 inside the calibrator's bounds) and the reference's own double result is
// only good to ~1e-10 there (tests/test_conditioning.py).  Here
//   xi-d  = -sigma^2 (iu+u^2)/(xi+d)      (or xi+d from xi-d when Re xi < 0)
//   log((1-g e)/(1-g)) = log1p-form of 1 + delta, delta = g (1-e)/(1-g)
// so this side carries ~1e-15 relative error everywhere; branch choices
// (principal csqrt with Re d >= 0, principal arg of the ratio) are the
// reference's.
//
// Everything is `double`; the file also compiles as plain C++ (g++) so the
// formulas can be checked on the CPU by tests/test_host_math.py -- that is a
// test harness, never a product path.
#pragma once
#include "fp64_math.cuh"

namespace hb {

struct cplx {
    double re, im;
};

// Explicit FMA placement: the contraction the compiler would pick may differ from one call site or template
// instantiation to the next, and the launch paths are compared bit for bit.
HB_HD cplx cmul(cplx a, cplx b) { return {fma_(a.re, b.re, -(a.im * b.im)), fma_(a.re, b.im, a.im * b.re)}; }

// Parameter-set constants shared by every grid point of a (kappa,sigma,rho) class.
struct ClassConst {
    double kappa, sigma2, rs;  // rs = rho*sigma
};

// T-independent stage, one per (class, u).
struct StageA {
    cplx d;    // principal sqrt, Re d >= 0                 heston.cpp:52
    cplx g;    // (xi-d)/(xi+d)                             heston.cpp:56
    cplx num;  // xi - d
    cplx q1;   // g/(1-g) = (xi-d)/(2d)
};

HB_HD StageA stage_a(const ClassConst& c, double ur, double ui) {
    StageA a;
    // xi = kappa - rho sigma i u ; i u = (-ui, ur)
    const double xr = c.kappa + c.rs * ui;
    const double xi = -c.rs * ur;
    // w = i u + u^2
    const double wr = ur * ur - ui * ui - ui;
    const double wi = 2.0 * ur * ui + ur;
    const double swr = c.sigma2 * wr, swi = c.sigma2 * wi;
    const double zr = xr * xr - xi * xi + swr;
    const double zi = 2.0 * xr * xi + swi;
    // d = csqrt(z), principal branch (glibc csqrt: Re >= 0, Im carries sign of zi)
    double m, rm_;
    sqrt_rsqrt(zr * zr + zi * zi, &m, &rm_);  // m = |z| = |d|^2
    const double h = 0.5 * (m + fabs(zr));
    double big, rh;
    sqrt_rsqrt(h, &big, &rh);            // sqrt(h), 1/sqrt(h)
    const double small = 0.5 * zi * rh;  // zi / (2 sqrt(h)), signed
    if (zr >= 0.0) {
        a.d.re = big;
        a.d.im = small;
    } else {
        a.d.re = fabs(small);
        a.d.im = copysign(big, zi);
    }
    // xi -/+ d without cancellation: (xi-d)(xi+d) = xi^2 - d^2 = -sigma^2 w
    cplx num, den;
    if (xr >= 0.0) {
        den = {xr + a.d.re, xi + a.d.im};
        const double r = rcp_nr(den.re * den.re + den.im * den.im);
        const cplx inv = {den.re * r, -den.im * r};
        num = cmul({-swr, -swi}, inv);
        a.g = cmul(num, inv);
    } else {
        num = {xr - a.d.re, xi - a.d.im};
        const double r = rcp_nr(num.re * num.re + num.im * num.im);
        const cplx inv = {num.re * r, -num.im * r};
        den = cmul({-swr, -swi}, inv);
        const double r2 = rcp_nr(den.re * den.re + den.im * den.im);
        a.g = cmul(num, {den.re * r2, -den.im * r2});
    }
    a.num = num;
    // q1 = num / (2 d) = num * conj(d) / (2 m)
    const double rm = 0.5 * rm_;  // 1/(2 m): sqrt_rsqrt(m^2) returned 1/m
    a.q1 = cmul(num, {a.d.re * rm, -a.d.im * rm});
    return a;
}

// Stage A for a point of the damped grid together with the Carr-Madan weight
//   tab = wgt / (alpha^2 + alpha - v^2 + i (2 alpha + 1) v)                         heston.cpp:117
// whose reciprocal's Newton steps are interleaved with the first square root's (two chains).
HB_HD StageA stage_a_tab(const ClassConst& c, double v, double ui, double alpha, double wgt, cplx* tab) {
    StageA a;
    const double xr = c.kappa + c.rs * ui;
    const double xi = -c.rs * v;
    const double wr = v * v - ui * ui - ui;
    const double wi = 2.0 * v * ui + v;
    const double ta = alpha * alpha + alpha - v * v;
    const double tb = (2.0 * alpha + 1.0) * v;
    const double swr = c.sigma2 * wr, swi = c.sigma2 * wi;
    const double zr = xr * xr - xi * xi + swr;
    const double zi = 2.0 * xr * xi + swi;
    const double m2 = fma_(zr, zr, zi * zi);
    const double tn = fma_(ta, ta, tb * tb);
    // m = sqrt(m2) = |z| and 1/m (coupled Newton)   ||   1/tn
    const double y = rsqrt_seed(m2);
    double yt = rcp_seed(tn);
    double g = m2 * y, h = 0.5 * y;
    double et = fma_(-tn, yt, 1.0);
    double r = fma_(-g, h, 0.5);
    yt = fma_(yt, et, yt);
    g = fma_(g, r, g);
    h = fma_(h, r, h);
    et = fma_(-tn, yt, 1.0);
    r = fma_(-g, h, 0.5);
    yt = fma_(yt, et, yt);
    g = fma_(g, r, g);
    h = fma_(h, r, h);
    const double dd = fma_(-g, g, m2);
    const double m = fma_(dd, h, g);
    const double rm = h;  // 1/(2 m)
    yt *= wgt;
    tab->re = ta * yt;
    tab->im = -tb * yt;
    // d = csqrt(z), principal branch
    const double hh = 0.5 * (m + fabs(zr));
    double big, rh;
    sqrt_rsqrt(hh, &big, &rh);
    const double small = 0.5 * zi * rh;
    a.d.re = (zr >= 0.0) ? big : fabs(small);
    a.d.im = (zr >= 0.0) ? small : copysign(big, zi);
    cplx num, den;
    if (xr >= 0.0) {
        den = {xr + a.d.re, xi + a.d.im};
        const double rr = rcp_nr(den.re * den.re + den.im * den.im);
        const cplx inv = {den.re * rr, -den.im * rr};
        num = cmul({-swr, -swi}, inv);
        a.g = cmul(num, inv);
    } else {
        num = {xr - a.d.re, xi - a.d.im};
        const double rr = rcp_nr(num.re * num.re + num.im * num.im);
        const cplx inv = {num.re * rr, -num.im * rr};
        den = cmul({-swr, -swi}, inv);
        const double r2 = rcp_nr(den.re * den.re + den.im * den.im);
        a.g = cmul(num, {den.re * r2, -den.im * r2});
    }
    a.num = num;
    a.q1 = cmul(num, {a.d.re * rm, -a.d.im * rm});
    return a;
}

// theta/v0-independent stage, one per (stage A, T).
struct StageB {
    cplx B;   // (xi-d) T - 2 log((1 - g e)/(1 - g))
    cplx Dq;  // (xi-d) (1-e)/(1-g e)
};

// ---- asymptotic stage B -------------------------------------------------------------------------------
// Once Re(d) T > 45, e = exp(-d T) is below 2.9e-20 in modulus and drops out of stage B in double
// precision: 1 - e = 1, 1 - g e = 1 (to far less than an ulp; what is left of it moves the exponent of phi
// by at most (4 kts + 3 v0s |num|) |e| < 1e-16 wherever phi has not underflowed), so
//   B  = (xi - d) T - 2 log(1/(1 - g)) = (xi - d) T - L0,   L0 = 2 clog(1 + g/(1-g))   T-independent
//   Dq = xi - d
// L0 is formed once per (class, grid point) by the same clog1p routine stage_b uses, fed with e = 0, and
// stage B costs two FMAs instead of a cexp and a clog.  On the Carr-Madan grid (N = 4096, eta = 0.25) 43 %
// of all (point, maturity) pairs are in this regime, 40 % have underflowed altogether, and 17 % need the
// full stage B.  The rule depends on the slice alone (its own d and T), so a slice still evaluates to the
// same bits whichever group or launch path it is priced in.
#ifndef HB_ASYM_DT
#define HB_ASYM_DT 45.0
#endif
constexpr double kAsymDT = HB_ASYM_DT;
constexpr double kMidDT = 10.0;  // below this L0 need not exist (decimation path) and |g e| <= 2^-17 is rare anyway

HB_HD cplx stage_b_l0(const StageA& a) {
    double lg, ar, rn;
    clog1p_rcp_nb(a.q1.re, a.q1.im, 1.0, &lg, &ar, &rn);
    return {lg, 2.0 * ar};
}
HB_HD StageB stage_b_asym(const StageA& a, cplx l0, double T) {
    StageB b;
    b.B.re = fma_(a.num.re, T, -l0.re);
    b.B.im = fma_(a.num.im, T, -l0.im);
    b.Dq = a.num;
    return b;
}

// Intermediate regime: with e = exp(-d T)
// known and x = g e small (|x| <= 2^-17, i.e. Re(d) T above ~12.5), log(1 - x) and 1/(1 - x) have three-term
// series whose truncation (|x|^4 / 4 < 1e-21) is far below rounding, so stage B needs the cexp but no clog
// and no reciprocal:
//   log(ratio) = log(1 + q1) + log(1 - x) = L0/2 - s,  s = x + x^2/2 + x^3/3   (argument wrapped to (-pi, pi])
//   Dq = num (1 - e) (1 + x + x^2 + x^3)
// tests/test_host_math.py compares it with the full stage B point by point.
HB_HD bool stage_b_mid_ok(const StageA& a, cplx e) {
    const double gr = a.g.re * e.re - a.g.im * e.im, gi = a.g.re * e.im + a.g.im * e.re;
    return gr * gr + gi * gi <= 5.8e-11;  // |g e|^2 <= 2^-34
}
HB_HD StageB stage_b_mid(const StageA& a, cplx l0, cplx e, double T) {
    StageB b;
    const cplx x = cmul(a.g, e);
    cplx t = {fma_(x.re, 1.0 / 3.0, 0.5), x.im * (1.0 / 3.0)};
    t = cmul(x, t);
    t.re += 1.0;
    const cplx s = cmul(x, t);  // -log(1 - x)
    const double pi = 3.14159265358979323846;
    double arg = fma_(0.5, l0.im, -s.im);  // principal value of the product's argument
    arg = (arg > pi) ? arg - 2.0 * pi : ((arg <= -pi) ? arg + 2.0 * pi : arg);
    b.B.re = fma_(a.num.re, T, -fma_(-2.0, s.re, l0.re));
    b.B.im = fma_(a.num.im, T, -(2.0 * arg));
    cplx u = {1.0 + x.re, x.im};
    u = cmul(x, u);
    u.re += 1.0;
    u = cmul(x, u);
    u.re += 1.0;  // 1 + x + x^2 + x^3
    const cplx ome = {1.0 - e.re, -e.im};
    b.Dq = cmul(a.num, cmul(ome, u));
    return b;
}

// Stage B given e = exp(-d T): everything after the cexp of heston.cpp:59.
HB_HD StageB stage_b_rest(const StageA& a, cplx e, double T) {
    StageB b;
    const cplx ome = {1.0 - e.re, -e.im};  // 1 - e
    // ratio = (1-g e)/(1-g) = 1 + delta, delta = g (1-e)/(1-g)
    const cplx dl = cmul(a.q1, ome);
    const cplx ge = cmul(a.g, e);
    const cplx n = {1.0 - ge.re, -ge.im};  // 1 - g e
    // lg = 2 Re log(ratio), ar = principal arg (as clog), rn = 1/|1 - g e|^2 -- evaluated together
    double lg, ar, rn;
    clog1p_rcp_nb(dl.re, dl.im, n.re * n.re + n.im * n.im, &lg, &ar, &rn);
    b.B.re = fma_(a.num.re, T, -lg);
    b.B.im = fma_(a.num.im, T, -(2.0 * ar));
    // Dq = num (1-e)/(1-g e)
    const cplx Q = cmul(ome, {n.re * rn, -n.im * rn});
    b.Dq = cmul(a.num, Q);
    return b;
}

HB_HD StageB stage_b(const StageA& a, double T) {
    // e = exp(-d T)                                          heston.cpp:59
    cplx e;
    cexp_nb(-a.d.re * T, -a.d.im * T, &e.re, &e.im);
    return stage_b_rest(a, e, T);
}

// Stage B by regime (the fused kernel's entry point): asymptotic beyond Re(d) T = 45, series form where
// |g e| <= 2^-17, the full form otherwise.  The choice depends on the slice alone.
#ifndef HB_MID
#define HB_MID 1
#endif
HB_HD StageB stage_b_auto(const StageA& a, cplx l0, double T, bool allow_mid = true) {
    if (a.d.re * T > kAsymDT) return stage_b_asym(a, l0, T);
    cplx e;
    cexp_nb(-a.d.re * T, -a.d.im * T, &e.re, &e.im);
    if (HB_MID && allow_mid && a.d.re * T > kMidDT && stage_b_mid_ok(a, e)) return stage_b_mid(a, l0, e, T);
    return stage_b_rest(a, e, T);
}

// Per-slice constants of stage F.
struct SliceConst {
    double kts;   // kappa*theta/sigma^2
    double v0s;   // v0/sigma^2
    double lsm;   // ln S0 + (r-q) T
};

// phi = exp(C + D v0 + i u (ln S0 + (r-q)T)),  i u = (-ui, ur)        heston.cpp:87-91
// Below kUnderflow exp gives exactly 0 in double precision.  `cut` >= kUnderflow is the significance cut of the
// Carr-Madan modes (kernels.cuh, GridConst::cut): |phi| = e^er < e^cut contributes less than the plan's
// admissible absolute price error even if every grid point sat at the cut, and is treated as 0 as well.
constexpr double kUnderflow = -746.0;
HB_HD cplx stage_f(const StageB& b, const SliceConst& s, double ur, double ui, double* er_out = nullptr,
                   double cut = kUnderflow) {
    // explicit FMA order: the fused kernel restates these two lines inline (cexp_w<3> path) and must round alike
    const double er = fma_(s.kts, b.B.re, fma_(s.v0s, b.Dq.re, -(ui * s.lsm)));
    const double ei = fma_(s.kts, b.B.im, fma_(s.v0s, b.Dq.im, ur * s.lsm));
    if (er_out) *er_out = er;  // log|phi|, used by the caller's decayed-tail bookkeeping
    // exp(er) underflows to exactly 0 below -745.14: the result is (+-0, +-0) whatever the phase, so the
    // cexp is skipped -- bit-identical output (sign of zero aside).  On the calibrator's box ~40 % of the
    // grid points of a slice are in this regime (the integrand has decayed); consecutive lanes hold
    // consecutive grid points, so whole warps take the short path.  NaN compares false and falls through.
    if (er < cut) return {0.0, 0.0};
    cplx phi;
    cexp_nb(er, ei, &phi.re, &phi.im);
    return phi;
}

// ---- decayed tail: a rigorous bound that needs no stage B ------------------------------------------
// stage_f returns exactly (0, 0) when er = Re(exponent) < -746 (exp underflows).  On the Carr-Madan grid
// that is ~40 % of all (point, maturity) pairs, and stage B (a cexp and a clog) is what tells.  The bound
// below decides most of them from stage A alone.  With E = |e| = exp(-Re d T), and |g| <= 2:
//   Re B  = Re(num) T - 2 log|1 - g e| + 2 log|1 - g|  <=  Re(num) T + l1g + 4E/(1-2E)
//           (|log|1-x|| <= |x|/(1-|x|), |x| = |g| E <= 2E;   l1g = 2 log|1 - g|)
//   Re Dq = Re(num) + Re(num (g-1) e/(1 - g e))        <=  Re(num) + |num| 3E/(1-2E)
// so for kappa theta/sigma^2 >= 0, v0/sigma^2 >= 0 and Re d T > 5 (E < e^-5)
//   er <= kts (Re(num) T + l1g + c1) + v0s (Re(num) + c2 |num|_1) - ui lsm =: ub,
// and ub < -750 implies that the computed er is below -746 (the 4 units of slack dwarf the rounding of
// B and Dq, ~1e-9 here) -- the slice's value at the point is exactly 0 and stage B / F are not evaluated.
// tests/test_host_math.py checks implication and coverage (89 % of the underflowed pairs) on the box.
HB_HD double tail_l1g(const StageA& a) {
    const double g2 = a.g.re * a.g.re + a.g.im * a.g.im;
    const double l = log1p_nb(fma_(-2.0, a.g.re, g2));  // log|1 - g|^2
    return (g2 <= 4.0) ? l : HUGE_VAL;                   // premise |g| <= 2 (NaN -> inf: never dead)
}
// Point part of the bound (shared by the slices of a group) and slice part.
struct TailPoint {
    double dre, nre, cD;  // Re d, Re(num), Re(num) + c2 |num|_1
};
HB_HD TailPoint tail_point(const StageA& a) {
    const double c2 = 0.02049;  // 3E/(1-2E) = 0.0204899..., E = e^-5, rounded up
    return {a.d.re, a.num.re, fma_(c2, fabs(a.num.re) + fabs(a.num.im), a.num.re)};
}
// Upper bound of er, or +inf when the premises fail (NaN inputs compare false -> +inf: never dead).
HB_HD double tail_ub(const TailPoint& t, double l1g, double T, double kts, double v0s, double lsm, double ui) {
    const double c1 = 0.02732;  // 4E/(1-2E) = 0.0273199..., rounded up
    const double ub = kts * (fma_(t.nre, T, l1g) + c1) + v0s * t.cD - ui * lsm;
    return ((t.dre * T > 5.0) && (kts >= 0.0) && (v0s >= 0.0)) ? ub : HUGE_VAL;
}
// l1g without the logarithm: |g| <= 2 gives |1 - g| <= 3, so 2 log|1 - g| <= 2 log 3.  Looser by a few units of
// kts at most (84 % instead of 89 % of the underflowed pairs), but free: used where stage A is not cached.
HB_HD double tail_l1g_const(const StageA& a) {
    const double g2 = a.g.re * a.g.re + a.g.im * a.g.im;
    return (g2 <= 4.0) ? 2.1973 : HUGE_VAL;  // 2 log 3 = 2.19722..., rounded up
}
HB_HD bool tail_dead(const StageA& a, double l1g, double T, double kts, double v0s, double lsm, double ui) {
    return tail_ub(tail_point(a), l1g, T, kts, v0s, lsm, ui) < -750.0;
}

// Full CF for arbitrary complex u (the bound API characteristic_function(u,T,S0,r,q)).
HB_HD cplx heston_cf(double kappa, double theta, double sigma, double rho, double v0, double ur, double ui,
                     double T, double S0, double r, double q) {
    const double L = log(S0);
    if (T <= 0.0) {  // heston.cpp:77-79: exp(i u ln S0)
        const double mag = exp_nb(-ui * L);
        double sn, cs;
        sincos_nb(ur * L, &sn, &cs);
        return {mag * cs, mag * sn};
    }
    ClassConst c = {kappa, sigma * sigma, rho * sigma};
    StageA a = stage_a(c, ur, ui);
    StageB b = stage_b(a, T);
    SliceConst s = {kappa * theta / c.sigma2, v0 / c.sigma2, L + (r - q) * T};
    return stage_f(b, s, ur, ui);
}

// Carr-Madan damping denominator 1/(alpha^2+alpha-v^2 + i(2 alpha+1) v)   heston.cpp:117
HB_HD cplx cm_inv_denominator(double v, double alpha) {
    const double a = alpha * alpha + alpha - v * v;
    const double b = (2.0 * alpha + 1.0) * v;
    const double r = rcp_nr(a * a + b * b);
    return {a * r, -b * r};
}

}  // namespace hb
