// prefix_bound.cuh -- a rigorous "live prefix" of the damped Carr-Madan integrand.
//
// The Carr-Madan modes treat a grid point whose log|phi| = Re(exponent) lies below the plan's significance cut as
// an exact zero (kernels.cuh, GridConst::cut).  The integrand decays in v, so on the calibrator's box the points
// that survive are a short prefix of the grid (median 216 of 4096 at N = 4096, eta = 0.25).  The direct-sum kernel
// (direct_kernel.cuh) evaluates only the points j < J of a slice; this header supplies a J for which EVERY grid
// point j >= J is provably below the cut, from the class constants alone (no characteristic function is
// evaluated).  Points below J are still tested one by one with the canonical exponent, so the rule "zero iff
// er < cut" is exactly the one the transform kernel applies to all N points.
//
// Notation (heston.cpp:37-70 with u = v - i (alpha+1)):  xi = kp - i rho sigma v,  kp = kappa - rho sigma (alpha+1),
//   d^2 = P(v) - i C v,  P = A0 + B v^2,  A0 = kp^2 - sigma^2 alpha (alpha+1),  B = sigma^2 (1-rho^2) > 0,
//   C = 2 kp rho sigma + sigma^2 (2 alpha + 1),   w(v) = i u + u^2,  |w| >= v^2,  |w| <= v^2 + (2 alpha+1) v + alpha (alpha+1),
//   (xi + d)(xi - d) = -sigma^2 w,   num = xi - d,  g = num/(xi + d),  q1 = num/(2 d),  e = exp(-d T),
//   er = -ui lsm + kts [Re(num) T - 2 log|1 + q1 (1 - e)|] + v0s Re(num (1 - e)/(1 - g e)).
// Facts used, for a block of grid points v in [va, vb], va > 0:
//   (1) Re d(v) = sqrt((|d^2| + P)/2) is non-decreasing in v >= 0:  d/dv (|d^2| + P) = v [2B(|d^2| + P) + C^2]/|d^2| >= 0.
//       Hence Re(num) <= kp - dra,  E = |e| in [Eb, Ea],  with dra = Re d(va), drb = Re d(vb).
//   (2) |xi + d| >= Re(xi + d) >= kp + dra =: sp (used only when sp > 0), so |num| <= sigma^2 wup(vb)/sp =: num_up,
//       |g| <= num_up/sp,  |q1| <= num_up/(2 dra)   (|d| >= Re d).
//   (3) |d|^2 <= D2(v) = |A0| + B v^2 + |C| v,  |xi|^2 = X2(v) = kp^2 + rho^2 sigma^2 v^2, and D2/v^2, X2/v^2 decrease in v:
//       |g| = |num|^2/(sigma^2 |w|) <= (sqrt X2 + sqrt D2)^2/(sigma^2 v^2) =: G2(va),
//       |1 - g| = 2 |d| |num|/(sigma^2 |w|) <= 2 sqrt D2 (sqrt X2 + sqrt D2)/(sigma^2 v^2) =: G1(va).
//   (4) |1 - e| <= (1 - Eb) + Ea min(2, |Im d| T),  |Im d| = |C| v/(2 Re d) <= |C| vb/(2 dra).
//   (5) -2 log|1 + q1 (1-e)| <= -2 log(1 - x), x = |q1|_up |1-e|_up < 1   (small sigma: the two logs of the closed
//       form cancel to first order and must not be bounded separately), or
//       = 2 log|1-g| - 2 log|1 - g e| <= 2 log(min(1 + gg, G1)) - 2 log(1 - gg Ea),  gg = min(|g| bounds)   (large v).
//   (6) Re(num (1-e)/(1-g e)) = Re(num) + Re(num e (g-1)/(1 - g e)) <= (kp - dra) + |num| |1-g| Ea/(1 - gg Ea),
//       |num| |1-g| = 2 |d| |g| <= min(2 sqrt D2(vb) gg, num_up min(1 + gg, G1)).
//   (7) small |d T| (short maturities, small sigma), z = d T/2 with z^2 = d^2 T^2/4 known without a square root:
//       Re B = kp T - 2 log|sinh z/z| - 2 log|xi T/2 + z coth z|,  Dq = -sigma^2 w (T/2)/(xi T/2 + z coth z), bounded from the
//       series of sinh z/z and z coth z in z^2 (prefix_terms below carries the remainders).
// The smaller of the bounds (5)/(7) on Re B and of (6)/(7) on Re Dq is taken.  With kts, v0s >= 0 their combination is an
// upper bound of er on the whole block; tests/test_host_math.py checks it against the exponent at every grid point of
// Sobol sets and all corners of the calibrator's box (several cuts, grids and dampings), and measures how much longer
// than the true live prefix the bound's prefix is: 7 % on the box (5 % for T >= 0.25, 9 % at T = 0.1).
// Compiles as plain C++ for that test (never a product path on the CPU).
#pragma once
#include "heston_math.cuh"

namespace hb {

struct PrefixClass {
    double kp, A0, B, C, sigma2, rs, rs2, alpha;
};

HB_HD PrefixClass prefix_class(const ClassConst& c, double alpha) {
    PrefixClass p;
    p.kp = c.kappa - c.rs * (alpha + 1.0);
    p.sigma2 = c.sigma2;
    p.rs = c.rs;
    p.rs2 = c.rs * c.rs;
    p.A0 = p.kp * p.kp - c.sigma2 * alpha * (alpha + 1.0);
    p.B = c.sigma2 - p.rs2;
    p.C = 2.0 * p.kp * c.rs + c.sigma2 * (2.0 * alpha + 1.0);
    p.alpha = alpha;
    return p;
}

// Re d(v), principal square root of P - i C v (cancellation-free on both signs of P)
// square root by the branch-free routine of fp64_math.cuh (~1 ulp; the bound carries kPrefixMargin of slack)
HB_HD double prefix_sqrt(double x) {
    double s, r;
    sqrt_rsqrt(x, &s, &r);
    return s;
}
HB_HD double prefix_dr(const PrefixClass& p, double v) {
    const double zr = p.A0 + p.B * v * v, zi = p.C * v;
    const double m = prefix_sqrt(zr * zr + zi * zi);
    // zr >= 0: sqrt((m + zr)/2);  zr < 0: |zi| / sqrt(2 (m - zr))  (the same number, without the cancellation)
    double s, r;
    sqrt_rsqrt((zr >= 0.0) ? 0.5 * (m + zr) : 2.0 * (m - zr), &s, &r);
    return (zr >= 0.0) ? s : fabs(zi) * r;
}

// T-independent part of the bound on the block of grid points [va, vb]
struct PrefixBlock {
    double nre;     // kp - dra >= Re(num)
    double dra, drb;
    double q1_up;   // >= |q1|                (inf when the premise sp > 0, dra > 0 fails)
    double di_up;   // >= |Im d|
    double gg;      // >= |g|
    double l1g;     // >= 2 log|1 - g|
    double n1g;     // >= |num| |1 - g|
    // small |d T| form (7)
    double kp, Pa, Pb, D2b;  // P(va), P(vb), D2(vb) >= |d|^2 on the block
    double bx;               // |rho sigma| vb >= |Im xi|
    double cv;               // |C| vb >= |Im d^2|
    double swr, swi;         // sigma^2 (va^2 - alpha(alpha+1)) <= sigma^2 Re w,  sigma^2 (2 alpha + 1) vb >= sigma^2 |Im w|
    double rsa, ca;          // rho sigma va, C va (signed): Im(xi T/2 + z^2/3) = -(rho sigma T/2 + C T^2/12) v
    double swia;             // sigma^2 (2 alpha + 1) va <= sigma^2 |Im w|
};

HB_HD PrefixBlock prefix_block(const PrefixClass& p, double va, double vb) {
    PrefixBlock b;
    const double inf = HUGE_VAL;
    b.dra = prefix_dr(p, va);
    b.drb = prefix_dr(p, vb);
    b.nre = p.kp - b.dra;
    const double sp = p.kp + b.dra;
    const bool ok = (sp > 0.0) && (b.dra > 0.0);
    const double wup = vb * vb + (2.0 * p.alpha + 1.0) * vb + p.alpha * (p.alpha + 1.0);
    const double rsp = ok ? rcp_nr(sp) : 0.0, rdra = ok ? rcp_nr(2.0 * b.dra) : 0.0;
    const double num_up = ok ? p.sigma2 * wup * rsp : inf;
    const double g_up = ok ? num_up * rsp : inf;
    b.q1_up = ok ? num_up * rdra : inf;
    b.di_up = ok ? fabs(p.C) * vb * rdra : inf;
    const double D2a = fabs(p.A0) + p.B * va * va + fabs(p.C) * va;
    const double D2b = fabs(p.A0) + p.B * vb * vb + fabs(p.C) * vb;
    const double X2a = p.kp * p.kp + p.rs2 * va * va;
    const double sd = prefix_sqrt(D2a), sx = prefix_sqrt(X2a);
    const double rs2v2 = rcp_nr(p.sigma2 * va * va);
    const double G2 = (sx + sd) * (sx + sd) * rs2v2;
    const double G1 = 2.0 * sd * (sx + sd) * rs2v2;
    b.gg = fmin(g_up, G2);
    const double one_g = fmin(1.0 + b.gg, G1);
    b.l1g = 2.0 * log1p_nb(one_g - 1.0);
    b.n1g = fmin(2.0 * prefix_sqrt(D2b) * b.gg, num_up * one_g);
    b.kp = p.kp;
    b.Pa = p.A0 + p.B * va * va;
    b.Pb = p.A0 + p.B * vb * vb;
    b.D2b = D2b;
    b.bx = fabs(p.rs) * vb;
    b.cv = fabs(p.C) * vb;
    b.swr = p.sigma2 * (va * va - p.alpha * (p.alpha + 1.0));
    b.swi = p.sigma2 * (2.0 * p.alpha + 1.0) * vb;
    b.rsa = p.rs * va;
    b.ca = p.C * va;
    b.swia = p.sigma2 * (2.0 * p.alpha + 1.0) * va;
    return b;
}

// Upper bounds of Re B and Re Dq over the block at maturity T (+inf where a premise fails; NaN inputs compare false
// and end in +inf or NaN: never "dead").  They do not depend on theta or v0: the variants of a class share them.
struct PrefixTerms {
    double reB, reDq;
};
HB_HD PrefixTerms prefix_terms(const PrefixBlock& b, double T) {
    const double inf = HUGE_VAL;
    const double Ea = exp_nb(-b.dra * T), Eb = exp_nb(-b.drb * T);
    const double ome = (1.0 - Eb) + Ea * fmin(2.0, b.di_up * T);
    const double x = b.q1_up * ome;
    const double gE = b.gg * Ea;
    const bool gok = gE < 0.5;
    // (5): the smaller of the two forms; one log1p serves both (its argument picked first)
    const double L2a = gok ? b.l1g : inf;             // + (-2 log1p(-gE))
    double Lt;
    if (x < 0.5 && gok) {
        const double l1 = -2.0 * log1p_nb(-x), l2 = L2a - 2.0 * log1p_nb(-gE);
        Lt = fmin(l1, l2);
    } else if (x < 0.5) {
        Lt = -2.0 * log1p_nb(-x);
    } else if (gok) {
        Lt = L2a - 2.0 * log1p_nb(-gE);
    } else {
        Lt = inf;
    }
    PrefixTerms t;
    t.reB = b.nre * T + Lt;
    t.reDq = gok ? b.nre + b.n1g * Ea * rcp_nr(1.0 - gE) : inf;
    // (7) small |d T| (short maturities, small sigma: e is not small and the forms above lose the cancellation between
    // Re(num) and the rest).  With z = d T/2 (z^2 = d^2 T^2/4 is known without a square root):
    //   Re B = kp T - 2 log|sinh z / z| - 2 log|xi T/2 + z coth z|,      Dq = -sigma^2 w (T/2)/(xi T/2 + z coth z),
    //   sinh z/z = 1 + z^2/6 + r1,  z coth z = 1 + z^2/3 + r2,  |r1| <= (|z|^4/120)/(1 - |z|^2/42),  |r2| <= (|z|^4/45)/(1 - |z|^2/pi^2),
    // with Re z^2 = P T^2/4 non-decreasing in v and |1 + x|^2 >= (1 + Re x)^2 + (|Im x| - |r|)^2.
    const double z2 = 0.25 * b.D2b * T * T;  // >= |z|^2 on the block
    if (z2 <= 2.0) {
        const double T2 = T * T, c12 = 1.0 / 12.0, c24 = 1.0 / 24.0;
        double q1r, q2r;  // 1 / (120 (1 - z2/42)),  1 / (45 (1 - z2/pi^2))
        rcp2_nr(120.0 * (1.0 - z2 * (1.0 / 42.0)), 45.0 * (1.0 - z2 * 0.10132118364233778), &q1r, &q2r);
        const double r1 = z2 * z2 * q1r;
        const double r2 = z2 * z2 * q2r;
        const double s1 = 1.0 + b.Pa * T2 * c24 - r1;
        const double s2 = 1.0 + 0.5 * b.kp * T + b.Pa * T2 * c12 - r2;
        if (s1 > 0.0 && s2 > 0.0) {
            // imaginary parts: |Im(z^2/6)| >= |C| va T^2/24 - r1;  Im(xi T/2 + z^2/3) = -cim v, cim = rho sigma T/2 + C T^2/12
            const double i1 = fmax(0.0, fabs(b.ca) * T2 * c24 - r1);
            const double cima = 0.5 * b.rsa * T + b.ca * T2 * c12;  // cim va (signed)
            const double i2 = fmax(0.0, fabs(cima) - r2);
            // log(m1) + log(m2) = log1p(m1 m2 - 1): the moduli are O(1) here (z2 <= 2)
            const double m1 = s1 * s1 + i1 * i1, m2 = s2 * s2 + i2 * i2;
            t.reB = fmin(t.reB, b.kp * T - log1p_nb(m1 * m2 - 1.0));
            const double s2hi = 1.0 + 0.5 * b.kp * T + b.Pb * T2 * c12 + r2;
            const double bup = 0.5 * b.bx * T + b.cv * T2 * c12 + r2;
            // Re(w conj(den)) = Re w Re den + Im w Im den, Im w = -(2 alpha + 1) v, Im den = -cim v + Im r2:
            // the second product is >= swia (cim va - r2) when cim > 0, >= -swi bup always
            const double cross = (cima > r2) ? b.swia * (cima - r2) : -b.swi * bup;
            const double numer = b.swr * s2 + cross;
            if (b.swr > 0.0 && numer > 0.0) t.reDq = fmin(t.reDq, -0.5 * T * numer * rcp_nr(s2hi * s2hi + bup * bup));
        }
    }
    return t;
}
// Upper bound of er = log|phi| over the block for a slice (kts, v0s; cst = -ui lsm)
HB_HD double prefix_ub_of(const PrefixTerms& t, double kts, double v0s, double cst) {
    const double ub = cst + kts * t.reB + v0s * t.reDq;
    return ((kts >= 0.0) && (v0s >= 0.0) && (ub == ub)) ? ub : HUGE_VAL;
}
HB_HD double prefix_ub(const PrefixBlock& b, double T, double kts, double v0s, double cst) {
    return prefix_ub_of(prefix_terms(b, T), kts, v0s, cst);
}

// Slack between the bound and the cut: covers the rounding of the bound itself and of the exponent the kernels
// compute (both ~1e-9 at worst); costs a fraction of a grid point of prefix.
constexpr double kPrefixMargin = 0.25;

// Block boundaries over the grid indices 1..N: blk[0] = 1 < blk[1] < ... < blk[nb] = N; block k = [blk[k], blk[k+1]).
// Geometric (ratio ~1.06) above 32 so that the prefix overshoots the first all-dead block start by ~3 % on average.
constexpr int kMaxPrefixBlocks = 160;
inline int prefix_blocks_host(int N, int* blk) {
    int nb = 0;
    blk[0] = 1;
    int cur = 1;
    while (cur < N) {
        int nx = (cur < 32) ? 32 : cur + (cur >> 4 > 1 ? cur >> 4 : 1);
        if (nx > N) nx = N;
        blk[++nb] = nx;
        cur = nx;
        if (nb >= kMaxPrefixBlocks - 1 && cur < N) {  // cannot happen for N <= 2^20; keep the table valid anyway
            blk[++nb] = N;
            break;
        }
    }
    return nb;
}

}  // namespace hb
