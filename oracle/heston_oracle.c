/*
 * heston_oracle.c -- CPU restatement of the reference's Heston pricing /
 * calibration-objective path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (pde_b200/) never links, imports or calls it.
 *
 * Parity pin status
 *   - orc_cf, orc_price_refgrid, orc_price_options_refgrid, the objective /
 *     residual / Jacobian functions in ORC_MODE_REFGRID: PINNED.  They are
 *     checked to ~1e-14 against the reference's own heston.cpp compiled
 *     unmodified into oracle/_ref/libheston_ref.so (see oracle/Makefile and
 *     tests/test_oracle.py) and against tests/golden/*.npz generated from
 *     that library and from the reference's Python calibrator.
 *   - ORC_MODE_FFT (Carr-Madan N-point FFT + log-strike interpolation):
 *     the reference has no FFT pricer, so no reference output
 *     exists for the transform/interpolation stage.  The spec followed is
 *     docs/models/heston-model.md:89-106 (N, eta, order of steps) plus
 *     Carr & Madan (1999) Simpson-weighted FFT, with psi, alpha, scaling and
 *     clamp/parity taken from src/cpp/models/heston.cpp:109-149.  Its CF
 *     stage is the pinned orc_cf; the stage after it is PINNED TO AN
 *     INDEPENDENT EVALUATION OF THAT SPEC: tests/golden/fft_direct.npz holds
 *     105 prices formed from the compiled reference's own CF values by plain
 *     O(N) direct sums at the two bracketing bins in 50-digit arithmetic
 *     (tests/golden/make_golden_fft_direct.py -- no FFT, no code shared with
 *     this file; N = 512 / 4096 / 16384); this restatement agrees with them
 *     to 0.09 x (1e-12 relative + 1e-14 absolute), tests/test_oracle.py.
 *
 * All citations are file:line in /root/reference.
 * Complex arithmetic is C99 <complex.h>; gcc lowers double complex * and /
 * to __muldc3/__divdc3, the same libgcc routines std::complex<double> uses,
 * and cexp/clog/csqrt are the glibc routines behind std::exp/log/sqrt, so
 * operation order is kept identical to the reference on purpose.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MODE_REFGRID 0
#define ORC_MODE_FFT 1

#define ORC_PI 3.14159265358979323846 /* src/cpp/models/heston.cpp:13 */

typedef double complex cplx;

static inline cplx mk(double re, double im) { return CMPLX(re, im); }

/* HestonParameters::is_valid -- src/cpp/models/heston.hpp:72-74 */
int orc_params_valid(const double *p) {
    return p[0] > 0.0 && p[1] > 0.0 && p[2] > 0.0 && fabs(p[3]) < 1.0 && p[4] > 0.0;
}

/* HestonParameters::is_feller_satisfied -- src/cpp/models/heston.hpp:65-67 */
int orc_feller(const double *p) { return 2.0 * p[0] * p[1] >= p[2] * p[2]; }

/*
 * characteristic_function + compute_cf_intermediates
 * src/cpp/models/heston.cpp:37-72 and :74-92.  p = {kappa,theta,sigma,rho,v0}.
 */
void orc_cf(const double *p, double ur, double ui, double T, double S0, double r, double q,
            double *out) {
    const double kappa = p[0], theta = p[1], sigma = p[2], rho = p[3], v0 = p[4];
    const cplx u = mk(ur, ui);
    const cplx i = mk(0.0, 1.0);
    cplx phi;
    if (T <= 0.0) { /* :77-79 */
        phi = cexp(i * u * log(S0));
    } else {
        const cplx sigma2 = mk(sigma * sigma, 0.0);                 /* :46 */
        cplx xi = kappa - rho * sigma * i * u;                      /* :51 */
        cplx d = csqrt(xi * xi + sigma2 * (i * u + u * u));         /* :52 */
        cplx g = (xi - d) / (xi + d);                               /* :56 */
        cplx e = cexp(-d * T);                                      /* :59 */
        cplx term1 = (xi - d) * T;                                  /* :63 */
        cplx term2 = -2.0 * clog((1.0 - g * e) / (1.0 - g));        /* :64 */
        cplx C = (mk(kappa * theta, 0.0) / sigma2) * (term1 + term2); /* :65 */
        cplx D = ((xi - d) / sigma2) * ((1.0 - e) / (1.0 - g * e)); /* :69 */
        cplx drift = (r - q) * i * u * T;                           /* :87 */
        phi = cexp(C + D * v0 + i * u * log(S0) + drift);           /* :91 */
    }
    out[0] = creal(phi);
    out[1] = cimag(phi);
}

/* price_option_integration -- src/cpp/models/heston.cpp:94-151 ("refgrid") */
double orc_price_refgrid(const double *p, double strike, double maturity, double spot, double rate,
                         double dividend, int is_call) {
    if (maturity <= 0.0) { /* :97-100 */
        return is_call ? fmax(spot - strike, 0.0) : fmax(strike - spot, 0.0);
    }
    const double alpha = 0.75; /* heston.hpp:261 */
    const double log_strike = log(strike);
    const double discount = exp(-rate * maturity);
    const int n_points = 1024; /* :126 */
    const double du = 0.01;    /* :127 */
    double integral = 0.0;     /* 0.5*integrand(0.0) == 0, :110,:131 */
    for (int j = 1; j < n_points; ++j) {
        double v = j * du;
        double ph[2];
        orc_cf(p, v, -(alpha + 1.0), maturity, spot, rate, dividend, ph);
        cplx phi = mk(ph[0], ph[1]);
        cplx numerator = cexp(-mk(0.0, 1.0) * v * log_strike);                     /* :116 */
        cplx denominator = mk(alpha * alpha + alpha - v * v, (2.0 * alpha + 1.0) * v); /* :117 */
        cplx res = numerator * phi / denominator;                                  /* :120 */
        integral += creal(res);
    }
    integral *= du;
    double call_price = (exp(-alpha * log_strike) / ORC_PI) * discount * integral; /* :139 */
    call_price = fmax(call_price, 0.0);                                            /* :142 */
    if (is_call) return call_price;
    double put = call_price - spot * exp(-dividend * maturity) + strike * discount; /* :148 */
    return fmax(put, 0.0);
}

/* Argument checks of price_option -- src/cpp/models/heston.cpp:156-164.
 * Returns 0 ok, 1 strike<=0, 2 spot<=0, 3 maturity<0. */
int orc_check_option(double strike, double maturity, double spot) {
    if (strike <= 0.0) return 1;
    if (spot <= 0.0) return 2;
    if (maturity < 0.0) return 3;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Carr-Madan FFT ("fft" mode)                          */

static void fft_radix2(cplx *x, int n) {
    /* plain iterative decimation-in-time, forward sign e^{-2 pi i jk/n} */
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            cplx t = x[i];
            x[i] = x[j];
            x[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1;
        for (int k = 0; k < half; ++k) {
            double ang = -2.0 * ORC_PI * (double)k / (double)len;
            cplx w = mk(cos(ang), sin(ang));
            for (int s = 0; s < n; s += len) {
                cplx a = x[s + k];
                cplx b = x[s + k + half] * w;
                x[s + k] = a + b;
                x[s + k + half] = a - b;
            }
        }
    }
}

/*
 * One slice: call prices C(k_m) on the log-strike grid k_m = -b + lambda m.
 *   v_j = eta j, lambda = 2 pi/(N eta), b = pi/eta
 *   psi_j = e^{-rT} phi(v_j - (alpha+1) i) / (alpha^2+alpha-v_j^2 + i(2 alpha+1) v_j)
 *           (numerator/denominator as heston.cpp:112-120, discount as :139)
 *   x_j = e^{i b v_j} psi_j w_j with e^{i b v_j} = e^{i pi j} = (-1)^j exactly,
 *   w_j = (eta/3)(3 + (-1)^{j+1} - delta_{j0})        (Carr-Madan Simpson rule)
 *   C(k_m) = e^{-alpha k_m}/pi * Re FFT(x)_m
 * grid must hold N doubles; work must hold N cplx.
 */
void orc_fft_slice(const double *p, double T, double S0, double r, double q, int N, double eta,
                   double alpha, double *grid, void *work) {
    cplx *x = (cplx *)work;
    const double lambda = 2.0 * ORC_PI / ((double)N * eta);
    const double b = ORC_PI / eta;
    const double disc = exp(-r * T);
    for (int j = 0; j < N; ++j) {
        double v = eta * (double)j;
        double ph[2];
        orc_cf(p, v, -(alpha + 1.0), T, S0, r, q, ph);
        cplx den = mk(alpha * alpha + alpha - v * v, (2.0 * alpha + 1.0) * v);
        cplx psi = disc * mk(ph[0], ph[1]) / den;
        double w = (eta / 3.0) * (j == 0 ? 1.0 : ((j & 1) ? 4.0 : 2.0));
        double sgn = (j & 1) ? -1.0 : 1.0;
        x[j] = psi * (w * sgn);
    }
    fft_radix2(x, N);
    for (int m = 0; m < N; ++m) {
        double k = -b + lambda * (double)m;
        grid[m] = exp(-alpha * k) / ORC_PI * creal(x[m]);
    }
}

/* Linear interpolation in log-strike + clamp/parity (heston.cpp:142-149).
 * Returns NaN if ln K falls outside [k_0, k_{N-1}). */
double orc_fft_interp(const double *grid, int N, double eta, double strike, double T, double S0,
                      double r, double q, int is_call) {
    const double lambda = 2.0 * ORC_PI / ((double)N * eta);
    const double b = ORC_PI / eta;
    double k = log(strike);
    double pos = (k + b) / lambda;
    double fm = floor(pos);
    if (!(fm >= 0.0) || !(fm < (double)(N - 1))) return NAN;
    int m = (int)fm;
    double km = -b + lambda * (double)m;
    double call = grid[m] + (grid[m + 1] - grid[m]) * (k - km) / lambda;
    call = fmax(call, 0.0);
    if (is_call) return call;
    double put = call - S0 * exp(-q * T) + strike * exp(-r * T);
    return fmax(put, 0.0);
}

/* ------------------------------------------------------------------------ */
/* Flat option lists: prices, objective, residuals, FD Jacobian              */

/*
 * Prices for a flat option list (strike[i], maturity[i], is_call[i]) -- the
 * contract of HestonCalibrator._price_options,
 * src/python/quant_trading/calibration/heston_calibrator.py:538-586:
 * an invalid parameter set or an invalid option yields NaN (the Python
 * wrapper's exception -> NaN at :583-584).
 */
void orc_price_options(int mode, const double *p, int n, const double *strike,
                       const double *maturity, const uint8_t *is_call, double S0, double r,
                       double q, int N, double eta, double alpha, double *out) {
    if (!orc_params_valid(p)) {
        for (int i = 0; i < n; ++i) out[i] = NAN;
        return;
    }
    if (mode == ORC_MODE_REFGRID) {
        for (int i = 0; i < n; ++i) {
            out[i] = orc_check_option(strike[i], maturity[i], S0)
                         ? NAN
                         : orc_price_refgrid(p, strike[i], maturity[i], S0, r, q, is_call[i]);
        }
        return;
    }
    double *grid = (double *)malloc(sizeof(double) * (size_t)N);
    cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)N);
    double lastT = NAN;
    for (int i = 0; i < n; ++i) {
        if (orc_check_option(strike[i], maturity[i], S0)) {
            out[i] = NAN;
            continue;
        }
        double T = maturity[i];
        if (T <= 0.0) {
            out[i] = is_call[i] ? fmax(S0 - strike[i], 0.0) : fmax(strike[i] - S0, 0.0);
            continue;
        }
        if (!(T == lastT)) {
            orc_fft_slice(p, T, S0, r, q, N, eta, alpha, grid, work);
            lastT = T;
        }
        out[i] = orc_fft_interp(grid, N, eta, strike[i], T, S0, r, q, is_call[i]);
    }
    free(grid);
    free(work);
}

/* _compute_objective -- heston_calibrator.py:486-513 */
double orc_objective_from_prices(int n, const double *price, const double *market) {
    for (int i = 0; i < n; ++i) {
        if (isnan(price[i]) || price[i] <= 0.0) return 1e10; /* :507-508 */
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double e = (price[i] - market[i]) / market[i]; /* :511 */
        s += e * e;
    }
    return s;
}

/* _compute_residuals -- heston_calibrator.py:515-536 (np.maximum propagates NaN) */
void orc_residuals_from_prices(int n, const double *price, const double *market, double *res) {
    for (int i = 0; i < n; ++i) {
        double pi = price[i];
        if (!isnan(pi)) pi = fmax(pi, 1e-10); /* :533 */
        res[i] = (pi - market[i]) / market[i]; /* :535 */
    }
}

/*
 * SciPy 2-point finite-difference step with bounds:
 * scipy/optimize/_numdiff.py _compute_absolute_step (rel_step=None) followed
 * by _adjust_scheme_to_bounds(scheme='1-sided', num_steps=1).
 */
void orc_fd_steps(const double *x, const double *lb, const double *ub, double *h) {
    const double rstep = 1.4901161193847656e-08; /* sqrt(DBL_EPSILON) */
    for (int j = 0; j < 5; ++j) {
        double sgn = (x[j] >= 0.0) ? 1.0 : -1.0;
        double hj = rstep * sgn * fmax(1.0, fabs(x[j]));
        double lower = x[j] - lb[j], upper = ub[j] - x[j];
        double xn = x[j] + hj;
        int violated = (xn < lb[j]) || (xn > ub[j]);
        int fitting = fabs(hj) <= fmax(lower, upper);
        if (violated && fitting) hj = -hj;
        if (!fitting) hj = (upper >= lower) ? upper : -lower;
        h[j] = hj;
    }
}

/*
 * Residual vector r0[n] and forward-difference Jacobian J[n][5] (row-major)
 * exactly as scipy.optimize.least_squares(jac='2-point', bounds=...) forms
 * them around HestonCalibrator._compute_residuals
 * (heston_calibrator.py:459-477; _numdiff.py _dense_difference '2-point').
 */
void orc_jacobian(int mode, const double *p, const double *lb, const double *ub, int n,
                  const double *strike, const double *maturity, const uint8_t *is_call,
                  const double *market, double S0, double r, double q, int N, double eta,
                  double alpha, double *r0, double *J) {
    double *price = (double *)malloc(sizeof(double) * (size_t)n);
    double *r1 = (double *)malloc(sizeof(double) * (size_t)n);
    orc_price_options(mode, p, n, strike, maturity, is_call, S0, r, q, N, eta, alpha, price);
    orc_residuals_from_prices(n, price, market, r0);
    double h[5];
    orc_fd_steps(p, lb, ub, h);
    for (int j = 0; j < 5; ++j) {
        double x1[5];
        memcpy(x1, p, sizeof(x1));
        x1[j] = p[j] + h[j];
        double dx = x1[j] - p[j];
        orc_price_options(mode, x1, n, strike, maturity, is_call, S0, r, q, N, eta, alpha, price);
        orc_residuals_from_prices(n, price, market, r1);
        for (int i = 0; i < n; ++i) J[(size_t)i * 5 + j] = (r1[i] - r0[i]) / dx;
    }
    free(price);
    free(r1);
}

/*
 * Normal-equation block for one parameter set: out[22] =
 *   [0]      objective loss (A9 semantics, 1e10 sentinel)
 *   [1]      ||r||^2 with r from _compute_residuals
 *   [2..6]   J^T r
 *   [7..21]  upper triangle of J^T J, row-major (00 01 02 03 04 11 12 ...)
 */
void orc_normal_eq(int mode, const double *p, const double *lb, const double *ub, int n,
                   const double *strike, const double *maturity, const uint8_t *is_call,
                   const double *market, double S0, double r, double q, int N, double eta,
                   double alpha, double *out) {
    double *r0 = (double *)malloc(sizeof(double) * (size_t)n);
    double *J = (double *)malloc(sizeof(double) * (size_t)n * 5);
    double *price = (double *)malloc(sizeof(double) * (size_t)n);
    orc_price_options(mode, p, n, strike, maturity, is_call, S0, r, q, N, eta, alpha, price);
    out[0] = orc_objective_from_prices(n, price, market);
    orc_jacobian(mode, p, lb, ub, n, strike, maturity, is_call, market, S0, r, q, N, eta, alpha,
                 r0, J);
    double rr = 0.0, jtr[5] = {0}, jtj[15] = {0};
    for (int i = 0; i < n; ++i) {
        rr += r0[i] * r0[i];
        int t = 0;
        for (int a = 0; a < 5; ++a) {
            jtr[a] += J[(size_t)i * 5 + a] * r0[i];
            for (int c = a; c < 5; ++c) jtj[t++] += J[(size_t)i * 5 + a] * J[(size_t)i * 5 + c];
        }
    }
    out[1] = rr;
    memcpy(out + 2, jtr, sizeof(jtr));
    memcpy(out + 7, jtj, sizeof(jtj));
    free(r0);
    free(J);
    free(price);
}

/* ------------------------------------------------------------------------ */
/* Batched (OpenMP over parameter sets) -- used for CPU baselines and tests   */

/* params: AoS double[P][5].  loss[P]. */
void orc_objective_batch(int mode, int P, const double *params, int n, const double *strike,
                         const double *maturity, const uint8_t *is_call, const double *market,
                         double S0, double r, double q, int N, double eta, double alpha,
                         double *loss) {
#pragma omp parallel
    {
        double *price = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(dynamic)
        for (int i = 0; i < P; ++i) {
            orc_price_options(mode, params + (size_t)i * 5, n, strike, maturity, is_call, S0, r, q,
                              N, eta, alpha, price);
            loss[i] = orc_objective_from_prices(n, price, market);
        }
        free(price);
    }
}

/* prices[P][n] */
void orc_price_batch(int mode, int P, const double *params, int n, const double *strike,
                     const double *maturity, const uint8_t *is_call, double S0, double r, double q,
                     int N, double eta, double alpha, double *prices) {
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < P; ++i) {
        orc_price_options(mode, params + (size_t)i * 5, n, strike, maturity, is_call, S0, r, q, N,
                          eta, alpha, prices + (size_t)i * n);
    }
}

/* out[P][22] */
void orc_normal_eq_batch(int mode, int P, const double *params, const double *lb,
                         const double *ub, int n, const double *strike, const double *maturity,
                         const uint8_t *is_call, const double *market, double S0, double r,
                         double q, int N, double eta, double alpha, double *out) {
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < P; ++i) {
        orc_normal_eq(mode, params + (size_t)i * 5, lb, ub, n, strike, maturity, is_call, market,
                      S0, r, q, N, eta, alpha, out + (size_t)i * 22);
    }
}

/* CF on a grid for many (param, T) pairs: out[P][M][n_u][2] -- parity of K1. */
void orc_cf_grid(int P, const double *params, int M, const double *T, int n_u, const double *ur,
                 double ui, double S0, double r, double q, double *out) {
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int i = 0; i < P; ++i) {
        for (int m = 0; m < M; ++m) {
            double *o = out + (((size_t)i * M + m) * (size_t)n_u) * 2;
            for (int j = 0; j < n_u; ++j) {
                orc_cf(params + (size_t)i * 5, ur[j], ui, T[m], S0, r, q, o + 2 * (size_t)j);
            }
        }
    }
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/*
 * Accuracy yardstick (NOT a reference output): the same formulas as orc_cf in
 * x87 long double (64-bit mantissa).  Used by tests/test_conditioning.py to
 * measure how far the reference's own double-precision result is from the
 * mathematically exact value where the little-trap form is ill-conditioned
 * (small sigma: kappa*theta/sigma^2 up to 1e5 inside the calibrator's bounds).
 */
void orc_cf_ld(const double *p, double ur, double ui, double T, double S0, double r, double q,
               double *out) {
    typedef long double complex lc;
    const long double kappa = p[0], theta = p[1], sigma = p[2], rho = p[3], v0 = p[4];
    const lc u = (long double)ur + (long double)ui * I;
    const lc i = I;
    const lc s2 = sigma * sigma;
    lc phi;
    if (T <= 0.0) {
        phi = cexpl(i * u * logl((long double)S0));
    } else {
        lc xi = kappa - rho * sigma * i * u;
        lc d = csqrtl(xi * xi + s2 * (i * u + u * u));
        lc g = (xi - d) / (xi + d);
        lc e = cexpl(-d * (long double)T);
        lc Cc = (kappa * theta / s2) * ((xi - d) * (long double)T - 2.0L * clogl((1.0L - g * e) / (1.0L - g)));
        lc D = ((xi - d) / s2) * ((1.0L - e) / (1.0L - g * e));
        phi = cexpl(Cc + D * v0 + i * u * logl((long double)S0) +
                    ((long double)r - (long double)q) * i * u * (long double)T);
    }
    out[0] = (double)creall(phi);
    out[1] = (double)cimagl(phi);
}

void orc_cf_ld_grid(const double *p, int n, const double *ur, double ui, double T, double S0, double r,
                    double q, double *out) {
    for (int j = 0; j < n; ++j) orc_cf_ld(p, ur[j], ui, T, S0, r, q, out + 2 * (size_t)j);
}
