// ref_shim.cpp -- extern "C" handle onto the UNMODIFIED reference Heston model.
//
// TEST INFRASTRUCTURE ONLY (see oracle/heston_oracle.c header).  This file is
// ours; it only #includes the reference header and is linked against the
// reference's own src/cpp/models/heston.cpp, compiled where it lies under
// /root/reference by oracle/Makefile.  Output: oracle/_ref/libheston_ref.so
// (git-ignored, travels to the GPU box with the snapshot).
//
// Each function forwards to one reference entry point:
//   ref_cf            -> HestonModel::characteristic_function (heston.cpp:74-92)
//   ref_price_option  -> HestonModel::price_option            (heston.cpp:153-167)
//   ref_price_options -> HestonModel::price_options           (heston.cpp:220-245, OpenMP)
//   ref_implied_vol   -> HestonModel::implied_volatility      (heston.cpp:311-349)
//   ref_greeks        -> HestonModel::price_option_with_greeks(heston.cpp:169-218)
//   ref_sabr_vols     -> SABRModel::implied_volatility        (sabr.cpp:130-192); NaN where it throws
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <exception>
#include <limits>
#include <string>
#include <vector>

#include "models/heston.hpp"
#include "models/sabr.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

using quant::models::HestonModel;
using quant::models::HestonParameters;

namespace {
thread_local std::string g_err;
inline HestonParameters mkp(const double* p) { return HestonParameters(p[0], p[1], p[2], p[3], p[4]); }
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// 0 ok, 1 std::invalid_argument (message in ref_last_error)
int ref_validate(const double* p) {
    try {
        mkp(p).validate();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int ref_cf(const double* p, double ur, double ui, double T, double S0, double r, double q, double* out) {
    try {
        HestonModel m(mkp(p));
        std::complex<double> z = m.characteristic_function({ur, ui}, T, S0, r, q);
        out[0] = z.real();
        out[1] = z.imag();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int ref_cf_grid(const double* p, int n, const double* ur, double ui, double T, double S0, double r, double q,
                double* out) {
    try {
        HestonModel m(mkp(p));
        for (int j = 0; j < n; ++j) {
            std::complex<double> z = m.characteristic_function({ur[j], ui}, T, S0, r, q);
            out[2 * j] = z.real();
            out[2 * j + 1] = z.imag();
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int ref_price_option(const double* p, double K, double T, double S0, double r, double q, int is_call, double* out) {
    try {
        HestonModel m(mkp(p));
        *out = m.price_option(K, T, S0, r, q, is_call != 0);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        *out = std::numeric_limits<double>::quiet_NaN();
        return 1;
    }
}

// maturities: nT == 1 or nT == n (heston.cpp:228-231)
int ref_price_options(const double* p, int n, const double* K, int nT, const double* T, double S0, double r,
                      double q, int is_call, double* out) {
    try {
        HestonModel m(mkp(p));
        std::vector<double> ks(K, K + n), ts(T, T + nT);
        std::vector<double> pr = m.price_options(ks, ts, S0, r, q, is_call != 0);
        if (n) std::memcpy(out, pr.data(), sizeof(double) * pr.size());
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Many parameter sets over one flat surface through the reference's own
// vectorised entry (OpenMP over options inside price_options).  prices[P][n].
// Invalid sets -> NaN row (what the calibrator's try/except produces).
int ref_price_surface_batch(int P, const double* params, int n, const double* K, const double* T, double S0,
                            double r, double q, double* prices) {
    std::vector<double> ks(K, K + n), ts(T, T + n);
    for (int i = 0; i < P; ++i) {
        try {
            HestonModel m(mkp(params + 5 * (size_t)i));
            std::vector<double> pr = m.price_options(ks, ts, S0, r, q, true);
            std::memcpy(prices + (size_t)i * n, pr.data(), sizeof(double) * (size_t)n);
        } catch (const std::exception&) {
            for (int j = 0; j < n; ++j) prices[(size_t)i * n + j] = std::numeric_limits<double>::quiet_NaN();
        }
    }
    return 0;
}

int ref_implied_vol(const double* p, double K, double T, double S0, double r, double q, int is_call, double* out) {
    try {
        HestonModel m(mkp(p));
        *out = m.implied_volatility(K, T, S0, r, q, is_call != 0);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// out = {price, delta, gamma, vega, theta, rho}
int ref_greeks(const double* p, double K, double T, double S0, double r, double q, int is_call, double* out) {
    try {
        HestonModel m(mkp(p));
        auto res = m.price_option_with_greeks(K, T, S0, r, q, is_call != 0);
        out[0] = res.price;
        out[1] = res.greeks.delta;
        out[2] = res.greeks.gamma;
        out[3] = res.greeks.vega;
        out[4] = res.greeks.theta;
        out[5] = res.greeks.rho;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// torchrun exports OMP_NUM_THREADS=1; the CPU baseline must say how many threads it really used.
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}


// vols[p * n + i] for P rows (alpha, rho, nu) and n strikes through the reference's SABRModel
void ref_sabr_vols(double beta, double F, double T, int n, const double* K, int P, const double* params, double* vols) {
    quant::models::SABRModel m(beta);
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < n; ++i) {
            try {
                vols[(size_t)p * n + i] = m.implied_volatility(K[i], F, T, params[3 * p], params[3 * p + 1], params[3 * p + 2]);
            } catch (const std::exception&) {
                vols[(size_t)p * n + i] = std::numeric_limits<double>::quiet_NaN();
            }
        }
}
}  // extern "C"
