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

// elementary routines of fp64_math.cuh (MUFU seeds emulated in float precision on the host)
extern "C" void hm_exp(int n, const double* x, double* y) { for (int i = 0; i < n; ++i) y[i] = hb::exp_nb(x[i]); }
extern "C" void hm_sincos(int n, const double* x, double* s, double* c) { for (int i = 0; i < n; ++i) hb::sincos_nb(x[i], s + i, c + i); }
extern "C" void hm_log1p(int n, const double* x, double* y) { for (int i = 0; i < n; ++i) y[i] = hb::log1p_nb(x[i]); }
extern "C" void hm_atan2(int n, const double* y, const double* x, double* a) { for (int i = 0; i < n; ++i) a[i] = hb::atan2_nb(y[i], x[i]); }
extern "C" void hm_div(int n, const double* a, const double* b, double* q) { for (int i = 0; i < n; ++i) q[i] = hb::div_nr(a[i], b[i]); }
extern "C" void hm_rcp(int n, const double* a, double* q) { for (int i = 0; i < n; ++i) q[i] = hb::rcp_nr(a[i]); }
extern "C" void hm_sqrt(int n, const double* a, double* s, double* r) { for (int i = 0; i < n; ++i) hb::sqrt_rsqrt(a[i], s + i, r + i); }

// fused, hand-interleaved routines
extern "C" void hm_cexp(int n, const double* x, const double* y, double* re, double* im) {
    for (int i = 0; i < n; ++i) hb::cexp_nb(x[i], y[i], re + i, im + i);
}
extern "C" void hm_clog1p(int n, const double* dr, const double* di, const double* w, double* lg, double* ar, double* rw) {
    for (int i = 0; i < n; ++i) hb::clog1p_rcp_nb(dr[i], di[i], w[i], lg + i, ar + i, rw + i);
}

// log|phi| (the real part of the CF exponent) on a grid: used to study how far the exponent of a
// finite-difference-perturbed parameter set can move away from the base set's (tail-skip margin).
extern "C" void hm_cf_exponent(const double* p, int n, const double* ur, double ui, double T, double S0, double r,
                               double q, double* er) {
    hb::ClassConst c = {p[0], p[2] * p[2], p[3] * p[2]};
    hb::SliceConst s = {p[0] * p[1] / c.sigma2, p[4] / c.sigma2, log(S0) + (r - q) * T};
    for (int j = 0; j < n; ++j) {
        hb::StageA a = hb::stage_a(c, ur[j], ui);
        hb::StageB b = hb::stage_b(a, T);
        hb::stage_f(b, s, ur[j], ui, er + j);
    }
}

// decayed-tail bound of heston_math.cuh next to the exponent stage F computes: dead[j] = tail_dead(...)
extern "C" void hm_tail_bound(const double* p, int n, const double* ur, double ui, double T, double S0, double r,
                              double q, double* er, int* dead) {
    hb::ClassConst c = {p[0], p[2] * p[2], p[3] * p[2]};
    hb::SliceConst s = {p[0] * p[1] / c.sigma2, p[4] / c.sigma2, log(S0) + (r - q) * T};
    for (int j = 0; j < n; ++j) {
        hb::StageA a = hb::stage_a(c, ur[j], ui);
        hb::StageB b = hb::stage_b(a, T);
        hb::stage_f(b, s, ur[j], ui, er + j);
        dead[j] = hb::tail_dead(a, hb::tail_l1g(a), T, s.kts, s.v0s, s.lsm, ui) ? 1 : 0;
    }
}

// asymptotic stage B next to the full one: exponent of phi from both, and Re(d) T
extern "C" void hm_stage_b_asym(const double* p, int n, const double* ur, double ui, double T, double S0, double r,
                                double q, double* er_full, double* ei_full, double* er_asym, double* ei_asym, double* dT) {
    hb::ClassConst c = {p[0], p[2] * p[2], p[3] * p[2]};
    hb::SliceConst s = {p[0] * p[1] / c.sigma2, p[4] / c.sigma2, log(S0) + (r - q) * T};
    for (int j = 0; j < n; ++j) {
        hb::StageA a = hb::stage_a(c, ur[j], ui);
        hb::StageB bf = hb::stage_b(a, T), ba = hb::stage_b_asym(a, hb::stage_b_l0(a), T);
        er_full[j] = s.kts * bf.B.re + s.v0s * bf.Dq.re - ui * s.lsm;
        ei_full[j] = s.kts * bf.B.im + s.v0s * bf.Dq.im + ur[j] * s.lsm;
        er_asym[j] = s.kts * ba.B.re + s.v0s * ba.Dq.re - ui * s.lsm;
        ei_asym[j] = s.kts * ba.B.im + s.v0s * ba.Dq.im + ur[j] * s.lsm;
        dT[j] = a.d.re * T;
    }
}

// intermediate-regime stage B next to the full one (flag = 1 where stage_b_mid_ok holds)
extern "C" void hm_stage_b_mid(const double* p, int n, const double* ur, double ui, double T, double S0, double r,
                               double q, double* er_full, double* ei_full, double* er_mid, double* ei_mid, int* flag) {
    hb::ClassConst c = {p[0], p[2] * p[2], p[3] * p[2]};
    hb::SliceConst s = {p[0] * p[1] / c.sigma2, p[4] / c.sigma2, log(S0) + (r - q) * T};
    for (int j = 0; j < n; ++j) {
        hb::StageA a = hb::stage_a(c, ur[j], ui);
        hb::cplx e;
        hb::cexp_nb(-a.d.re * T, -a.d.im * T, &e.re, &e.im);
        hb::StageB bf = hb::stage_b(a, T), bm = hb::stage_b_mid(a, hb::stage_b_l0(a), e, T);
        er_full[j] = s.kts * bf.B.re + s.v0s * bf.Dq.re - ui * s.lsm;
        ei_full[j] = s.kts * bf.B.im + s.v0s * bf.Dq.im + ur[j] * s.lsm;
        er_mid[j] = s.kts * bm.B.re + s.v0s * bm.Dq.re - ui * s.lsm;
        ei_mid[j] = s.kts * bm.B.im + s.v0s * bm.Dq.im + ur[j] * s.lsm;
        flag[j] = hb::stage_b_mid_ok(a, e) ? 1 : 0;
    }
}


// live-prefix bound of prefix_bound.cuh (direct-sum kernel): J[t] for each maturity T[t] such that every grid point
// j >= J[t] has an exponent below `cut`; one parameter set, one class (its own kts / v0s).
#include <vector>

#include "../pde_b200/csrc/prefix_bound.cuh"
extern "C" int hm_prefix_J(const double* p, int nT, const double* T, int N, double eta, double alpha, double S0, double r,
                           double q, double cut, int* J) {
    hb::ClassConst c = {p[0], p[2] * p[2], p[3] * p[2]};
    const hb::PrefixClass pc = hb::prefix_class(c, alpha);
    std::vector<int> blk(hb::kMaxPrefixBlocks + 1);
    const int nb = hb::prefix_blocks_host(N, blk.data());
    std::vector<hb::PrefixBlock> blocks(nb);
    for (int k = 0; k < nb; ++k) blocks[k] = hb::prefix_block(pc, eta * blk[k], eta * (blk[k + 1] - 1));
    const double kts = p[0] * p[1] / c.sigma2, v0s = p[4] / c.sigma2, ui = -(alpha + 1.0);
    for (int t = 0; t < nT; ++t) {
        const double cst = -ui * (log(S0) + (r - q) * T[t]);
        int last_live = -1;
        for (int k = nb - 1; k >= 0; --k)
            if (!(hb::prefix_ub(blocks[k], T[t], kts, v0s, cst) < cut - hb::kPrefixMargin)) {
                last_live = k;
                break;
            }
        J[t] = blk[last_live + 1];
    }
    return nb;
}
