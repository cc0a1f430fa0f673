// Test harness: compiles the product's device math header (pde_b200/csrc/heston_math.cuh)
// as plain C++ so tests/test_host_math.py can check the FORMULAS on the CPU against the
// oracle.  Not a product path: nothing under pde_b200/ builds or loads this file.
#include "../pde_b200/csrc/heston_math.cuh"

extern "C" void hm_cf_grid(const double* p, int n, const double* ur, double ui, double T, double S0, double r,
                           double q, double* out) {
    for (int j = 0; j < n; ++j) {
        hb::cplx z = hb::heston_cf(p[0], p[1], p[2], p[3], p[4], ur[j], ui, T, S0, r, q);
        out[2 * j] = z.re;
        out[2 * j + 1] = z.im;
    }
}
extern "C" void hm_cf(const double* p, double ur, double ui, double T, double S0, double r, double q, double* out) {
    hb::cplx z = hb::heston_cf(p[0], p[1], p[2], p[3], p[4], ur, ui, T, S0, r, q);
    out[0] = z.re;
    out[1] = z.im;
}
