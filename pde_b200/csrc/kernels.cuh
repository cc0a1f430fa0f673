// kernels.cuh -- sm_100a kernels of the batched Heston pricing / calibration-objective path.
//
//   fft_job_kernel      K1+K2+K3 fused: one CTA per parameter-set job; characteristic
//                       function on the damped grid -> slices in shared memory -> in-place
//                       radix-8 DIF transform pruned to the quoted bins -> log-strike interpolation -> prices / loss /
//                       finite-difference normal equations.  psi never touches HBM.
//   refgrid_job_kernel  K1': the reference's own arithmetic (1023-point quadrature,
//                       heston.cpp:94-151) as CF-once-per-slice + per-strike twiddle sums.
//   cf_kernel           K1 alone: phi(u; T) for arbitrary complex u (heston.cpp:74-92).
//   fft_batch_kernel    K2 alone: bulk-async (TMA) staged batched FFT, HBM -> smem -> HBM.
//   dfma_peak_kernel    FP64-pipe roofline probe.
//
// Job = one base parameter set.  In normal-equation / Jacobian mode the job prices 6
// variants (base + 5 forward-difference perturbations, SciPy step rule)
// and shares work between them:
//   class 0 = {base, theta+h, v0+h}: same (kappa,sigma,rho) => same stage A and stage B;
//             only the final cexp differs -> the three slices of one maturity are one group.
//   classes kappa+h, sigma+h, rho+h: own stage A, shared by the 3 maturities of a group.
// A group = up to 3 slices that are resident in shared memory at once (3 x 64 KiB at
// N = 4096).  Arithmetic of a perturbed slice is identical to evaluating it on its own.
// Work without effect is not executed: stage A is cached per (set, class) in thread-private
// local memory; stage F skips the cexp where |phi| is below the plan's significance cut (exactly 0 below exp's
// underflow; GridConst::cut bounds what all dropped points together can add to a price); the perturbed classes
// skip stage B/F where the base set's integrand lies 54 units of log|phi| below that cut -- a MEASURED margin
// (condition number of the exponent < 200 over the calibrator's default box, tests/test_host_math.py), so that
// skip is switched off for any other box (hb_set_bounds); the first two transform passes skip rows that are exact
// zeros; CTAs pull jobs from a global counter because jobs then differ in cost.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_smem.cuh"
#include "heston_math.cuh"

namespace hb {

// Phase probe (diagnostic builds only, -DHB_PROBE: benchmarks/step_rate.py --probe): lane 0 of every warp adds the
// cycles it spends in each phase of a group to a global table.  Not compiled into the product library.
#ifdef HB_PROBE
__device__ unsigned long long g_probe[16];
__device__ __forceinline__ long long probe_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");  // not to be moved across barriers / memory ops
    return t;
}
#define HB_PROBE_T(var) const long long var = probe_clock()
#define HB_PROBE_ADD(slot, t0, t1) \
    do {                           \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_probe[slot], (unsigned long long)((t1) - (t0))); \
    } while (0)
#else
#define HB_PROBE_T(var)
#define HB_PROBE_ADD(slot, t0, t1)
#endif

enum What { W_PRICE = 0, W_LOSS = 1, W_NEQ = 2, W_JAC = 3 };

// Device view of the option surface, grouped by distinct maturity (CSR).  "sorted" arrays
// are in maturity-group order; opt_orig maps back to the caller's option index.
struct SurfaceDev {
    int n_opt, n_mat, n_intr;
    double spot, rate, dividend, ln_spot;
    const double* mat_T;     // [n_mat] distinct maturities > 0
    const double* mat_disc;  // exp(-r T)
    const double* mat_fwd;   // S0 exp(-q T)
    const int* mat_off;      // [n_mat+1]
    const int* opt_orig;     // [n_sorted]
    const int* opt_bin;      // FFT: lower grid bin m, -1 if ln K is off the grid
    const double* opt_frac;  // FFT: (k - k_m)/lambda
    const double* opt_s0;    // FFT: exp(-alpha k_m)/pi        REFGRID: exp(-alpha ln K)/pi
    const double* opt_s1;    // FFT: exp(-alpha k_{m+1})/pi
    const double* opt_lnk;   // ln K
    const double* opt_kdisc; // K exp(-r T) (put parity, heston.cpp:148)
    const uint8_t* opt_call;
    const unsigned* need_mask;  // [n_mat] FFT: digits m_{L-2} of the quoted bins (outputs kept by the S = 8 pass)
    // FFT, bin-sum epilogue: the DISTINCT grid bins each maturity quotes (ascending) and, per option, the indices of
    // its two bracketing bins in its maturity's list (-1: ln K is off the grid)
    const int* bin_off;  // [n_mat+1]
    const int* bin_m;    // [bin_off[n_mat]] full-grid bin index m
    const int* opt_b0;   // [n_sorted]
    const int* opt_b1;
    const int* intr_orig;    // options not priced through a slice (T == 0 or invalid)
    const double* intr_val;  // intrinsic value or NaN
    const double* mkt_orig;  // [n_opt] market prices in caller order (may be null for W_PRICE)
};

struct GridConst {
    double eta, alpha, ui, w0;  // ui = -(alpha+1), w0 = eta/3
    // Grids longer than one CTA's shared memory (N = R * Nsub, e.g. 16384 = 4 * 4096; a 256 KiB
    // slice does not fit 227 KB) are transformed by decimation in time: phase ph holds the points
    // j = ph + R jj, its Nsub-point FFT Y_ph is taken on chip, and
    //   X[m] = sum_ph W_N^{ph m} Y_ph[m mod Nsub]
    // is accumulated only at the quoted bins.  R = 1 is the plain single-transform path.
    int R, n_full;
    // Significance cut (hb_plan_set_truncation): a grid point with log|phi| < cut is treated as exactly 0.
    // |x_j| = |phi_j| |tab_j| < e^cut |tab_j|, so all dropped points together move X_m by less than
    // e^cut sum_j |tab_j| and a price by less than that times max_o(scale_o) -- the host picks cut so that this
    // bound equals the plan's admissible absolute price error (default 2^-80; 0 = exact: cut = -746, where exp
    // underflows).  cut_dead = cut - 54: margin below which the perturbed classes trust the base set's decay.
    double cut, cut_dead;
    int tail_ok;  // FD bounds inside the box the perturbed-class tail skip was validated on
};

struct Bounds {
    double lb[5], ub[5];
};

constexpr int kMaxGroup = kMaxGroupFft;

struct SubSlice {
    double T, kts, v0s, lsm, disc, fwd;
    int mat, variant;
    int o0, o1;  // options of this maturity: sorted indices [o0, o1)
    int b0, nb;  // its distinct quoted bins: S.bin_m[b0 .. b0 + nb)
};
struct Group {
    ClassConst cc;
    int count;
    int max_opt;                // max over slices of o1 - o0
    unsigned fmask[kMaxGroup];  // need_mask of each slice's maturity
    SubSlice s[kMaxGroup];
};

// Variant v of the base set x: v = 0 base, v = 1..5 perturbs x[v-1].
struct JobState {
    double x[6][5];
    double dx[5];
    double rdx[5];  // 1 / dx
    int valid;
};

// HestonParameters::is_valid, heston.hpp:72-74
__device__ __forceinline__ bool params_valid(const double* x) {
    return x[0] > 0.0 && x[1] > 0.0 && x[2] > 0.0 && fabs(x[3]) < 1.0 && x[4] > 0.0;
}

// SciPy 2-point step with bounds (scipy/optimize/_numdiff.py
// _compute_absolute_step + _adjust_scheme_to_bounds '1-sided').
__device__ __forceinline__ double fd_step(double x, double lb, double ub) {
    const double rstep = 1.4901161193847656e-08;
    double h = rstep * (x >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x));
    const double lower = x - lb, upper = ub - x;
    const double xn = x + h;
    const bool violated = (xn < lb) || (xn > ub);
    const bool fitting = fabs(h) <= fmax(lower, upper);
    if (violated && fitting) h = -h;
    if (!fitting) h = (upper >= lower) ? upper : -lower;
    return h;
}

__device__ __forceinline__ void job_setup(JobState& js, const double* params, int ld, int p, const Bounds& bd,
                                          int V) {
    double x[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) x[c] = params[(size_t)c * ld + p];
    js.valid = params_valid(x);
#pragma unroll
    for (int v = 0; v < 6; ++v)
#pragma unroll
        for (int c = 0; c < 5; ++c) js.x[v][c] = x[c];
    if (V > 1) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const double h = fd_step(x[c], bd.lb[c], bd.ub[c]);
            const double x1 = x[c] + h;
            js.x[c + 1][c] = x1;
            js.dx[c] = x1 - x[c];  // (x0 + h) - x0, as _dense_difference does
            js.rdx[c] = 1.0 / js.dx[c];
        }
    }
}

// Sub-slice list of a class.  In 6-variant mode class 0 enumerates (maturity, {base,
// theta', v0'}) so that a group of 3 shares one maturity; other classes enumerate maturities.
__device__ __forceinline__ void sub_slice_at(int cls, int V, int t, int& variant, int& mat) {
    if (V > 1 && cls == 0) {
        const int r = t % 3;
        mat = t / 3;
        variant = (r == 0) ? 0 : (r == 1 ? 2 : 5);
    } else {
        mat = t;
        variant = cls;
    }
}

__device__ __forceinline__ void fill_group(Group& grp, const JobState& js, const SurfaceDev& S, int cls, int V,
                                           int t0, int n_total, int gmax) {
    const double* xc = js.x[cls];
    grp.cc.kappa = xc[0];
    grp.cc.sigma2 = xc[2] * xc[2];
    grp.cc.rs = xc[3] * xc[2];
    const int cnt = min(gmax, n_total - t0);
    grp.count = cnt;
    int max_opt = 0;
    for (int g = 0; g < cnt; ++g) {
        int variant, mat;
        sub_slice_at(cls, V, t0 + g, variant, mat);
        const double* xv = js.x[variant];
        SubSlice& s = grp.s[g];
        s.variant = variant;
        s.mat = mat;
        s.T = S.mat_T[mat];
        s.kts = xv[0] * xv[1] / grp.cc.sigma2;  // kappa*theta/sigma^2, heston.cpp:65
        s.v0s = xv[4] / grp.cc.sigma2;
        s.lsm = S.ln_spot + (S.rate - S.dividend) * s.T;
        s.disc = S.mat_disc[mat];
        s.fwd = S.mat_fwd[mat];
        s.o0 = S.mat_off[mat];
        s.o1 = S.mat_off[mat + 1];
        max_opt = max(max_opt, s.o1 - s.o0);
        s.b0 = S.bin_off ? S.bin_off[mat] : 0;
        s.nb = S.bin_off ? S.bin_off[mat + 1] - s.b0 : 0;
        grp.fmask[g] = S.need_mask ? S.need_mask[mat] : 0xffu;
    }
    grp.max_opt = max_opt;
}

// ---- block reductions ---------------------------------------------------------------------

// FT <= NT: only the first FT threads carry partial sums (finalize_job: the reduction order then depends on FT alone,
// so kernels with different block sizes agree bit for bit); every thread of the block takes part in the barriers.
template <int NT, int W, int FT = NT>
__device__ __forceinline__ void block_sum(double (&acc)[W], double* red /* [FT/32][W] */, int tid) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0 && warp < FT / 32) {
#pragma unroll
        for (int i = 0; i < W; ++i) red[warp * W + i] = acc[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            double v = (lane < FT / 32) ? red[lane * W + i] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[i] = v;
        }
    }
    __syncthreads();
}

// Threads that carry the sums of finalize_job in the Carr-Madan job kernels: 256 whatever the block size (the
// transform kernel runs 512 threads, the direct-sum kernel may run another shape, and a routed launch prices some
// sets of a batch with one and some with the other -- same reduction order, same bits).
template <int NT>
__host__ __device__ constexpr int kFinalizeT() { return NT >= 256 ? 256 : NT; }

// _compute_residuals, heston_calibrator.py:533-535 (np.maximum propagates NaN)
// One division per option (1 / market) and one per parameter and job (1 / dx) instead of twelve per option: the
// residual (p - m) (1/m) is within 1.5 ulp of the reference's (p - m)/m.
__device__ __forceinline__ double residual_of(double price, double mkt, double rmkt) {
    const double p = (price != price) ? price : fmax(price, 1e-10);
    return (p - mkt) * rmkt;
}

// Turn the price rows of one job into the requested output.  rows[v * n_opt + i].
template <int NT, int FT = NT>
__device__ __forceinline__ void finalize_job(int what, const double* rows, const SurfaceDev& S, const JobState& js,
                                             int p, double* out, double* out2, double* red, int tid) {
    static_assert(FT <= NT && FT % 32 == 0, "FT threads of the block carry the sums");
    const int n = S.n_opt;
    if (what == W_PRICE) return;
    if (what == W_LOSS) {
        double acc[2] = {0.0, 0.0};
        for (int i = tid; i < n && tid < FT; i += FT) {
            const double pr = rows[i], m = S.mkt_orig[i];
            if (pr != pr || pr <= 0.0) {
                acc[1] += 1.0;  // heston_calibrator.py:507-508
            } else {
                const double e = (pr - m) * (1.0 / m);
                acc[0] += e * e;
            }
        }
        block_sum<NT, 2, FT>(acc, red, tid);
        if (tid == 0) out[p] = (acc[1] > 0.0) ? 1e10 : acc[0];
        return;
    }
    if (what == W_JAC) {
        for (int i = tid; i < n; i += NT) {
            const double m = S.mkt_orig[i], rm = 1.0 / m;
            const double r0 = residual_of(rows[i], m, rm);
            out[(size_t)p * n + i] = r0;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const double rc = residual_of(rows[(size_t)(c + 1) * n + i], m, rm);
                out2[((size_t)p * n + i) * 5 + c] = (rc - r0) * js.rdx[c];
            }
        }
        return;
    }
    // W_NEQ: { loss, ||r||^2, J^T r, triu(J^T J) }
    double acc[23];
#pragma unroll
    for (int i = 0; i < 23; ++i) acc[i] = 0.0;
    for (int i = tid; i < n && tid < FT; i += FT) {
        const double pr = rows[i], m = S.mkt_orig[i], rm = 1.0 / m;
        if (pr != pr || pr <= 0.0) {
            acc[22] += 1.0;
        } else {
            const double e = (pr - m) * rm;
            acc[0] += e * e;
        }
        const double r0 = residual_of(pr, m, rm);
        double J[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) J[c] = (residual_of(rows[(size_t)(c + 1) * n + i], m, rm) - r0) * js.rdx[c];
        acc[1] += r0 * r0;
        int t = 7;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            acc[2 + a] += J[a] * r0;
#pragma unroll
            for (int c = a; c < 5; ++c) acc[t++] += J[a] * J[c];
        }
    }
    block_sum<NT, 23, FT>(acc, red, tid);
    if (tid == 0) {
        double* o = out + (size_t)p * 22;
        o[0] = (acc[22] > 0.0) ? 1e10 : acc[0];
        for (int i = 1; i < 22; ++i) o[i] = acc[i];
    }
}

// Outputs of a job whose base parameter set is invalid: NaN prices -> 1e10 loss.
template <int NT>
__device__ __forceinline__ void invalid_job(int what, const SurfaceDev& S, int p, double* out, double* out2,
                                            int tid) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int n = S.n_opt;
    if (what == W_PRICE) {
        for (int i = tid; i < n; i += NT) out[(size_t)p * n + i] = nan;
    } else if (what == W_LOSS) {
        if (tid == 0) out[p] = 1e10;
    } else if (what == W_NEQ) {
        if (tid < 22) out[(size_t)p * 22 + tid] = (tid == 0) ? 1e10 : nan;
    } else {
        for (int i = tid; i < n; i += NT) {
            out[(size_t)p * n + i] = nan;
            for (int c = 0; c < 5; ++c) out2[((size_t)p * n + i) * 5 + c] = nan;
        }
    }
}

// clamp + put-call parity of price_option_integration, heston.cpp:142-149.
// std::max(x, 0.0) keeps a NaN x (x < 0 is false), unlike fmax.
__device__ __forceinline__ double finish_price(double call, bool is_call, double fwd, double kdisc) {
    call = (call < 0.0) ? 0.0 : call;
    if (is_call) return call;
    const double put = call - fwd + kdisc;
    return (put < 0.0) ? 0.0 : put;
}

// The three final exponentials of a class-0 point (base, theta', v0': same stage B).  theta' and v0' move the exponent
// by dl = dkts B + dv0s Dq, ~1e-8 |B|: where |dl| <= 2^-17, phi' = phi e^{dl} = phi (1 + dl + dl^2/2 + dl^3/6) to 1e-22
// relative -- one cexp instead of three, and dl is formed from the exact parameter differences, so phi' - phi (what the
// finite difference reads) carries less rounding than two separate exponentials would.  Otherwise (finite-difference
// steps limited by a bound, NaN) the three full exponentials.  kts / v0s: of the three variants, base first.
__device__ __forceinline__ void class0_cexp(const double* er, const double* ei, const StageB& b, const double* kts,
                                            const double* v0s, double* pr, double* pi) {
    cplx dl[3];
    bool small = true;
#pragma unroll
    for (int g = 1; g < 3; ++g) {
        const double dk = kts[g] - kts[0], dv = v0s[g] - v0s[0];
        dl[g] = {fma(dk, b.B.re, dv * b.Dq.re), fma(dk, b.B.im, dv * b.Dq.im)};
        small = small && (fma(dl[g].re, dl[g].re, dl[g].im * dl[g].im) <= 5.8e-11);
    }
    if (small) {
        cexp_w<1>(er, ei, pr, pi);
        const cplx phi = {pr[0], pi[0]};
#pragma unroll
        for (int g = 1; g < 3; ++g) {
            cplx t = {fma(dl[g].re, 1.0 / 3.0, 1.0), dl[g].im * (1.0 / 3.0)};  // 1 + dl/3
            t = cmul(dl[g], t);
            t = {fma(t.re, 0.5, 1.0), t.im * 0.5};                             // 1 + dl/2 (1 + dl/3)
            t = cmul(dl[g], t);                                               // e^{dl} - 1
            const cplx u = cmul(phi, t);
            pr[g] = phi.re + u.re;
            pi[g] = phi.im + u.im;
        }
    } else {
        cexp_w<3>(er, ei, pr, pi);
    }
}

// ============================================================================================
// Fused Carr-Madan FFT job kernel
// ============================================================================================

// Threads per CTA of the fused kernel.  Measured on B200 (profiles/r01_shape_sweep.txt): 512 threads
// x 128 registers is the knee -- 640/768/1024 threads spill or idle (4096 points do not tile) and are
// 3-25 % slower; interleaving 2 or 4 grid points per thread does not fit the register file.
#ifndef HB_NT4096
#define HB_NT4096 512
#endif
constexpr int kNT4096 = HB_NT4096, kNT512 = 128;

// ONEVAR: the one-variant modes (W_PRICE, W_LOSS) and the six-variant modes (W_NEQ, W_JAC) are separate
// instantiations: each carries only its own tail logic (rigorous bound / dead masks + cexp_w<3>).
template <int N, int NT, bool DECIM, bool ONEVAR>
__global__ void __launch_bounds__(NT, 1)
fft_job_kernel(SurfaceDev S, GridConst gc, Bounds bd, const double* __restrict__ params, int ld, int P, int what,
               double* __restrict__ out, double* __restrict__ out2, double* __restrict__ scratch, int gmax,
               int split, unsigned long long* job_counter, const int* __restrict__ job_ids = nullptr,
               const int* __restrict__ p_count = nullptr, const int* __restrict__ jtab = nullptr) {
    // job_ids / p_count: the parameter sets routed to this kernel by prefix_scan_kernel (direct_kernel.cuh) and
    // their number, both on the device; null = all P sets in order.  jtab[p][maturity]: live-prefix bound of that
    // launch (prefix_bound.cuh): grid points at or beyond it are exact zeros and are not evaluated.
    if (p_count) P = *p_count;
    if (P <= 0) return;  // nothing was routed here: skip the table set-up
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* slices = reinterpret_cast<double2*>(smem_raw);
    double2* tw = slices + (size_t)gmax * N;
    double2* tw512 = tw + N / 8;
#ifndef HB_ZERO_AWARE
#define HB_ZERO_AWARE 1
#endif
    // zero-aware first two transform passes (fft_smem.cuh): only where butterfly `tid` owns thread `tid`'s points
    constexpr bool kZeroAware = HB_ZERO_AWARE && N == 4096 && N / 8 == NT;
    __shared__ int s_jlive[2][kMaxGroup];  // [group parity][slice]: 1 + highest non-zero point index
    __shared__ Group grps[2];  // descriptor of the running group and, prefetched during its K1, of the next one
    __shared__ JobState js;
    __shared__ double red[(NT / 32) * 23];
    const int tid = threadIdx.x;
    constexpr int V = ONEVAR ? 1 : 6;
    const int n = S.n_opt, M = S.n_mat;
    // groups of one parameter set: class 0 first, then kappa', sigma', rho' (6-variant mode)
    const int n_cls = (V == 1) ? 1 : 4;
    const int g0 = ((V > 1 ? 3 * M : M) + gmax - 1) / gmax;
    const int g1 = (M + gmax - 1) / gmax;
    const int groups_per_set = g0 + (n_cls - 1) * g1;
    // split = one job per (set, group) with rows in global memory and a separate finalize
    // kernel (small P: fills the SMs); otherwise one job per set, finalized in place.
    const long long n_jobs = split ? (long long)P * groups_per_set : (long long)P;

    fill_twiddles<N>(tw, tid, NT);
    fill_tw512(tw512, tid, NT);

    constexpr int PTS = DECIM ? 1 : (N + NT - 1) / NT;  // grid points owned by one thread
    static_assert(PTS <= 8, "one byte of dead-point flags per maturity");
    StageA ac[PTS];
    cplx tabc[PTS];
    double l1gc[PTS];  // 2 log|1 - g| (decayed-tail bound of the one-variant modes, heston_math.cuh)
    cplx l0c[PTS];     // L0 of the asymptotic stage B (heston_math.cuh), per (class, point) like stage A
    // deadm: one byte per maturity m, bit k: the base set's phi at this thread's k-th point and maturity m
    // has underflowed with a wide margin (see K1).  Only available when one CTA prices all classes of a set.
    // Indexed dynamically, once per group: lives in local memory, not in registers.
    constexpr int kDeadWords = 16;  // 128 maturities
    unsigned long long deadm[kDeadWords];
// Rigorous tail bound in the cached (N = 4096) one-variant path: worth +13 % before the asymptotic stage B,
// -4 % after it (a decayed point now costs two FMAs of stage B and the exponent): off.  The decimation path
// keeps it (+8 % at N = 16384, eta = 0.25, where 72 % of the grid has decayed and stage A is not cached).
#ifndef HB_BOUND
#define HB_BOUND 0
#endif
#ifndef HB_BOUND_DECIM
#define HB_BOUND_DECIM 1
#endif
#ifndef HB_F3
#define HB_F3 1
#endif
#ifndef HB_F3_PERT
#define HB_F3_PERT 1  // class-0 variants by expansion around the base exponential (class0_cexp)
#endif
#ifndef HB_TAIL
#define HB_TAIL 1
#endif
    const bool track_tail = HB_TAIL && gc.tail_ok && !DECIM && !split && V > 1 && M <= 8 * kDeadWords;

    // Jobs differ in cost (decayed tails take the short path of stage F), so after its first,
    // statically assigned job a CTA pulls the next one from a global counter.
    __shared__ long long s_job;
    auto next_job = [&](long long cur) -> long long {
        if (!job_counter) return cur + gridDim.x;
        __syncthreads();
        if (tid == 0) s_job = (long long)gridDim.x + (long long)atomicAdd(job_counter, 1ULL);
        __syncthreads();
        return s_job;
    };
    // group gi of a set -> (class, first sub-slice)
    auto group_at = [&](int gi, int& cls, int& n_total, int& t0) {
        const int ci = (gi < g0) ? 0 : 1 + (gi - g0) / g1;
        cls = (ci == 0) ? 0 : (ci == 1 ? 1 : (ci == 2 ? 3 : 4));  // base | kappa' | sigma' | rho'
        n_total = (V > 1 && cls == 0) ? 3 * M : M;
        t0 = ((gi < g0) ? gi : (gi - g0) % g1) * gmax;
    };

    // K3 works on all slices of a group at once: TPG threads per slice, two threads per option (one per
    // bracketing bin).  The option's constants are loaded one step ahead (before the last FFT pass /
    // before the previous batch is finished) so their latency is off the critical path.
#ifndef HB_BINSUM
#define HB_BINSUM 1
#endif
    struct K3Opt {
        int o, bin, orig, call;
        double s0, s1, frac, kdisc;
    };
#if HB_BINSUM
    // K3': one thread per option; `bin` = index of the lower bracketing bin in the maturity's distinct-bin list,
    // `call` = is_call | index of the upper one << 1
    constexpr int kMaxBins = 512;  // distinct bins per maturity (hb_surface_set splits longer maturities)
    __shared__ double s_xs[kMaxGroup][kMaxBins];
    constexpr int TPG = NT / kMaxGroup;
    const int k3_g = tid / TPG, k3_lt = tid - k3_g * TPG;
    // (o0, o1): the option range of this thread's slice, copied out of the group descriptor by the caller -- the
    // descriptor buffer may be refilled for group gi + 2 once the slices are released
    auto k3_load = [&](int o0, int o1, int base) -> K3Opt {
        K3Opt q;
        q.o = -1;
        q.bin = -1;
        q.orig = q.call = 0;
        q.s0 = q.s1 = q.frac = q.kdisc = 0.0;
        {
            const int o = o0 + base + k3_lt;
            if (o < o1) {
                q.o = o;
                q.bin = S.opt_b0[o];
                q.orig = S.opt_orig[o];
                q.call = (int)S.opt_call[o] | (S.opt_b1[o] << 1);
                q.s0 = S.opt_s0[o];
                q.s1 = S.opt_s1[o];
                q.frac = S.opt_frac[o];
                q.kdisc = S.opt_kdisc[o];
            }
        }
        return q;
    };
#else
    constexpr int TPG = (NT / kMaxGroup) & ~1;
    const int k3_g = tid / TPG, k3_lt = tid - k3_g * TPG, k3_half = k3_lt & 1;
    auto k3_load = [&](const Group& grp, int base) -> K3Opt {
        K3Opt q;
        q.o = -1;
        q.bin = -1;
        q.orig = q.call = 0;
        q.s0 = q.s1 = q.frac = q.kdisc = 0.0;
        if (k3_g < grp.count) {
            const int o = grp.s[k3_g].o0 + base + (k3_lt >> 1);
            if (o < grp.s[k3_g].o1) {
                q.o = o;
                q.bin = S.opt_bin[o];
                if (k3_half == 0) {
                    q.orig = S.opt_orig[o];
                    q.call = S.opt_call[o];
                    q.s0 = S.opt_s0[o];
                    q.s1 = S.opt_s1[o];
                    q.frac = S.opt_frac[o];
                    q.kdisc = S.opt_kdisc[o];
                }
            }
        }
        return q;
    };
#endif

    for (long long job = blockIdx.x; job < n_jobs; job = next_job(job)) {
        const int pj = split ? (int)(job / groups_per_set) : (int)job;
        const int p = job_ids ? job_ids[pj] : pj;
        const int gi_begin = split ? (int)(job % groups_per_set) : 0;
        const int gi_end = split ? gi_begin + 1 : groups_per_set;
        int cached_cls = -1;    // class whose stage A sits in ac[]
        int job_has_dead = -1;  // -1: not known yet (class 0 still running)
        if (track_tail) {
#pragma unroll
            for (int w = 0; w < kDeadWords; ++w) deadm[w] = 0ull;
        }
        __syncthreads();  // previous job's finalize has consumed rows/js
        if (tid == 0) {
#pragma unroll
            for (int g = 0; g < kMaxGroup; ++g) s_jlive[0][g] = s_jlive[1][g] = 0;
            job_setup(js, params, ld, p, bd, V);
            if (js.valid) {
                int cls, n_total, t0;
                group_at(gi_begin, cls, n_total, t0);
                fill_group(grps[gi_begin & 1], js, S, cls, V, t0, n_total, gmax);
            }
        }
        __syncthreads();
        if (!js.valid) {
            if (!split) invalid_job<NT>(what, S, p, out, out2, tid);
            else if (what == W_PRICE && gi_begin == 0) invalid_job<NT>(what, S, p, out, out2, tid);
            continue;
        }
        double* rows = (what == W_PRICE) ? out + (size_t)p * n
                                         : scratch + (size_t)(split ? p : (int)blockIdx.x) * 6 * n;
        if (gi_begin == 0) {
            for (int v = 0; v < V; ++v)
                for (int i = tid; i < S.n_intr; i += NT) rows[(size_t)v * n + S.intr_orig[i]] = S.intr_val[i];
        }
        for (int gi = gi_begin; gi < gi_end; ++gi) {
            int cls, n_total, t0;
            group_at(gi, cls, n_total, t0);
            const Group& grp = grps[gi & 1];  // filled before the previous barrier
            // The next group's descriptor is built by one thread while K1 runs (its global loads would
            // otherwise sit between two barriers with the whole CTA waiting); the other buffer was last
            // read before the barrier that ended group gi - 1.
            if (tid == NT - 1 && gi + 1 < gi_end) {
                int cls2, n2, t2;
                group_at(gi + 1, cls2, n2, t2);
                fill_group(grps[(gi + 1) & 1], js, S, cls2, V, t2, n2, gmax);
            }
            if (track_tail && cls != 0 && job_has_dead < 0) {
                // class 0 is done: does any thread of the CTA hold a decayed point?  (block-uniform)
                unsigned long long any = 0ull;
#pragma unroll
                for (int w = 0; w < kDeadWords; ++w) any |= deadm[w];
                job_has_dead = __syncthreads_or(any != 0ull);
            }
            const int count = grp.count;
            const int R = DECIM ? gc.R : 1;  // DECIM = false: the plain single-transform kernel (N == Nsub)
            // Stage A (and the Carr-Madan weight) depends on the class and the grid point only: compute it
            // once per (set, class) for this thread's points and keep it in a thread-private cache
            // (local memory, L2-resident: 80 B/point) instead of once per group of 3 slices.
            if (!DECIM && cls != cached_cls) {
                const ClassConst cc = grp.cc;
#pragma unroll 1
                for (int k = 0, j = tid; j < N; ++k, j += NT) {
                    const double wgt = gc.w0 * (j == 0 ? 1.0 : ((j & 1) ? -4.0 : 2.0));
                    ac[k] = stage_a_tab(cc, gc.eta * (double)j, gc.ui, gc.alpha, wgt, &tabc[k]);
                    l0c[k] = stage_b_l0(ac[k]);
                    // 2 log|1 - g| of the tail bound is -Re L0 (1 + g/(1-g) = 1/(1-g)); premise |g| <= 2 as in tail_l1g
                    if (ONEVAR && HB_BOUND) {
                        const double g2 = ac[k].g.re * ac[k].g.re + ac[k].g.im * ac[k].g.im;
                        l1gc[k] = (g2 <= 4.0) ? -l0c[k].re : HUGE_VAL;
                    }
                }
                cached_cls = cls;
            }
            for (int ph = 0; ph < R; ++ph) {
            HB_PROBE_T(pt0);
            unsigned live = 0u;  // bit 8 g + k: this thread's k-th point of slice g is non-zero (zero-aware K2)
            // ---- K1: characteristic function on the damped grid -> x_j in shared memory ----
            // One grid point per thread at a time: interleaving points explicitly was measured slower
            // (profiles/r01_shape_sweep.txt); the ILP comes from the interleaved chains inside stage B / F.
            {
                const ClassConst cc = grp.cc;
                // Decayed tail of a perturbed class: where the base set's log|phi| is below -800 the
                // kappa'/sigma'/rho' slice (parameters moved by 1.5e-8 relative) is exactly 0 as well --
                // the exponent would have to move by 54, i.e. a condition number above 4e6; measured
                // over the box it is below 200 (tests/test_host_math.py).  Points where every slice of
                // the group is dead skip stage B and F.  skip bit k <-> this thread's k-th point.
                unsigned skip = 0u;
                if (track_tail && cls != 0 && job_has_dead) {
                    skip = 0xffu;
                    for (int g = 0; g < count; ++g) {
                        const int mat = grp.s[g].mat;
                        skip &= (unsigned)(deadm[mat >> 3] >> ((mat & 7) * 8));
                    }
                    skip &= 0xffu;
                }
                if (jtab) {  // beyond the live prefix of every slice of the group
                    int jlim = 0;
                    for (int g = 0; g < count; ++g) jlim = max(jlim, jtab[(size_t)p * M + grp.s[g].mat]);
#pragma unroll
                    for (int k = 0; k < PTS; ++k)
                        if (ph + R * (tid + k * NT) >= jlim) skip |= 1u << k;
                }
                // class 0 records the base slice's decayed points (slice 0 of a class-0 group is the base set)
                const bool record = track_tail && cls == 0 && grp.s[0].variant == 0;
                // a class-0 group of the 6-variant mode is {base, theta', v0'} of ONE maturity: stage B is shared
                const bool share_b = V > 1 && cls == 0 && gmax == 3;
                unsigned dmask = 0u;
                // One-variant modes (prices, objective: population search): the rigorous decayed-tail bound of
                // heston_math.cuh decides from stage A alone where a slice is exactly 0 -- bit 3k + g of `gone`
                // -- so stage B is not evaluated there either.  Done up front for all points of the group: the
                // cache reads pipeline, and the main loop never touches the cache for a point it skips.
                // (For class 0 of the six-variant mode it was measured a loss: stage B is shared by three
                // slices there, the bound costs more than it saves -- 17.7 vs 18.5 M slices/s.)
                unsigned gone = 0u;
                if (!DECIM && ONEVAR && HB_BOUND) {
#pragma unroll 2
                    for (int k = 0; k < PTS; ++k) {
                        const TailPoint tp = tail_point(ac[k]);
                        const double l1g = l1gc[k];
                        unsigned bits = 0u;
                        for (int g = 0; g < count; ++g) {
                            const SubSlice& s = grp.s[g];
                            bits |= (tail_ub(tp, l1g, s.T, s.kts, s.v0s, s.lsm, gc.ui) < gc.cut - 4.0 ? 1u : 0u) << g;
                        }
                        gone |= bits << (3 * k);
                        if (bits == (1u << count) - 1u) skip |= 1u << k;
                    }
                }
                // Stage A of a point comes back from the thread-private cache (local memory, an L2 hit of
                // several hundred cycles).  It is fetched one point ahead, into the registers of the point
                // being finished, as soon as that point's last stage B has consumed them -- the latency then
                // hides behind the remaining stage F (ncu r01_g: 3.3 % of all samples sat on the first use).
                const int last_b = share_b ? 0 : count - 1;  // slice whose stage B is the last one of a point
                StageA a = {};
                cplx tab = {0.0, 0.0}, tab_n = {0.0, 0.0}, l0 = {0.0, 0.0};
                if (!DECIM) {
                    a = ac[0];
                    tab_n = tabc[0];
                    l0 = l0c[0];
                }
#pragma unroll 1
                for (int k = 0, j0 = tid; j0 < N; ++k, j0 += NT) {
                    HB_ASSERT(k < 8 && count >= 1 && count <= gmax && gmax <= kMaxGroup);
                    const int j = ph + R * j0;  // index on the full N-point grid
                    const double v = gc.eta * (double)j;
                    const int kn = (k + 1 < PTS) ? k + 1 : k;
                    if ((skip >> k) & 1u) {
                        for (int g = 0; g < count; ++g) sts_c(slices + (size_t)g * N, j0, {0.0, 0.0});
                        // fetch stage A only for a point that will be evaluated: back-to-back fetches into
                        // the same registers would serialise runs of skipped points on the L2 latency
                        if (!DECIM && !((skip >> kn) & 1u)) {
                            a = ac[kn];
                            tab_n = tabc[kn];
                            l0 = l0c[kn];
                        }
                        continue;
                    }
                    if (DECIM) {
                        // Simpson weight times e^{i b v_j} = (-1)^j   
                        const double wgt = gc.w0 * (j == 0 ? 1.0 : ((j & 1) ? -4.0 : 2.0));
                        a = stage_a_tab(cc, v, gc.ui, gc.alpha, wgt, &tab);
                        // Long grids reach far into the decayed tail (N = 16384, eta = 0.25: 72 % of all
                        // (point, maturity) pairs): the rigorous bound of heston_math.cuh, in its log-free
                        // form, says from stage A alone where a slice is exactly 0.
                        const TailPoint tp = tail_point(a);
                        const double l1g = tail_l1g_const(a);
                        unsigned bits = 0u;
                        for (int g = 0; g < (HB_BOUND_DECIM ? count : 0); ++g) {
                            const SubSlice& s = grp.s[g];
                            bits |= (tail_ub(tp, l1g, s.T, s.kts, s.v0s, s.lsm, gc.ui) < gc.cut - 4.0 ? 1u : 0u) << g;
                        }
                        gone = (gone & ~(7u << (3 * k))) | (bits << (3 * k));
                        if (bits == (1u << count) - 1u) {
                            for (int g = 0; g < count; ++g) sts_c(slices + (size_t)g * N, j0, {0.0, 0.0});
                            continue;
                        }
                        // L0 of the asymptotic / series stage B, only where some slice of the group can use it
                        // (the series form is not used here: without the cache its L0 would cost a clog per point and group)
                        if (a.d.re * grp.s[count - 1].T > kAsymDT) l0 = stage_b_l0(a);
                    } else {
                        tab = tab_n;
                    }
                    StageB b = {};
                    if (HB_F3 && share_b && count == 3) {
                        // {base, theta', v0'} of one maturity: one stage B, then the three final cexps in one
                        // interleaved evaluation (they underflow together: the sets differ by 1.5e-8 relative)
                        {
                            const double T0 = grp.s[0].T;
                            b = stage_b_auto(a, l0, T0, !DECIM);
                        }
                        if (!DECIM) {
                            a = ac[kn];
                            tab_n = tabc[kn];
                            l0 = l0c[kn];
                        }
                        double er[3], ei[3];
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            const SubSlice& s = grp.s[g];
                            er[g] = fma(s.kts, b.B.re, fma(s.v0s, b.Dq.re, -(gc.ui * s.lsm)));  // stage_f, heston.cpp:87-91
                            ei[g] = fma(s.kts, b.B.im, fma(s.v0s, b.Dq.im, v * s.lsm));
                        }
                        dmask |= (er[0] < gc.cut_dead ? 1u : 0u) << k;
                        double pr[3] = {0.0, 0.0, 0.0}, pi[3] = {0.0, 0.0, 0.0};
                        if (!(er[0] < gc.cut && er[1] < gc.cut && er[2] < gc.cut)) {
#if HB_F3_PERT
                            const double ks[3] = {grp.s[0].kts, grp.s[1].kts, grp.s[2].kts};
                            const double vs3[3] = {grp.s[0].v0s, grp.s[1].v0s, grp.s[2].v0s};
                            class0_cexp(er, ei, b, ks, vs3, pr, pi);
#else
                            cexp_w<3>(er, ei, pr, pi);
#endif
                        }
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            const bool zero = er[g] < gc.cut;  // exactly 0 as in stage_f
                            const cplx phi = {zero ? 0.0 : pr[g], zero ? 0.0 : pi[g]};
                            live |= (zero ? 0u : 1u) << (8 * g + k);
                            sts_c(slices + (size_t)g * N, j0, cmul(phi, tab));
                        }
                        continue;
                    }
#pragma unroll 1
                    for (int g = 0; g < count; ++g) {
                        const SubSlice& s = grp.s[g];
                        // this slice alone is exactly 0 here (never with a shared stage B: all or nothing there)
                        const bool g_gone = (ONEVAR || DECIM) && !share_b && ((gone >> (3 * k + g)) & 1u);
                        if ((g == 0 || !share_b) && !g_gone)
                            b = stage_b_auto(a, l0, s.T, !DECIM);
                        if (!DECIM && g == last_b && !(ONEVAR && ((skip >> kn) & 1u))) {
                            a = ac[kn];
                            tab_n = tabc[kn];
                            l0 = l0c[kn];
                        }
                        if (g_gone) {
                            sts_c(slices + (size_t)g * N, j0, {0.0, 0.0});
                            continue;
                        }
                        const SliceConst sc = {s.kts, s.v0s, s.lsm};
                        double er;
                        const cplx phi = stage_f(b, sc, v, gc.ui, &er, gc.cut);
                        live |= (er < gc.cut ? 0u : 1u) << (8 * g + k);
                        if (g == 0) dmask |= (er < gc.cut_dead ? 1u : 0u) << k;
                        sts_c(slices + (size_t)g * N, j0, cmul(phi, tab));
                    }
                }
                if (record) {
                    const int mat = grp.s[0].mat;
                    deadm[mat >> 3] |= (unsigned long long)dmask << ((mat & 7) * 8);
                }
                if (kZeroAware) {
                    // 1 + the highest non-zero point index of each slice, for the second pass (fft_smem.cuh)
                    if (tid == 0 && ph == 0) {
#pragma unroll
                        for (int g = 0; g < kMaxGroup; ++g) s_jlive[(gi + 1) & 1][g] = 0;  // next group's
                    }
#pragma unroll
                    for (int g = 0; g < kMaxGroup; ++g) {
                        const unsigned lm = (live >> (8 * g)) & 0xffu;
                        int jm = lm ? tid + NT * (31 - __clz(lm)) + 1 : 0;
                        jm = __reduce_max_sync(0xffffffffu, jm);
                        if ((tid & 31) == 0 && jm > 0) atomicMax(&s_jlive[gi & 1][g], jm);
                    }
                }
            }
            // ---- K2: in-place decimation-in-frequency passes in shared memory (fft_smem.cuh) ----
            // Pass 1: butterfly `tid` reads the points tid + r N/8 this thread has just written -> no barrier.
            HB_PROBE_T(pt1);
            if (N / 8 != NT) __syncthreads();
            if constexpr (kZeroAware) dif_pass_first<N, NT>(slices, count, tw, live, tid);
            else dif_pass<N, NT, N / 8, false>(slices, count, tw, tw512, grp.fmask, tid);
            HB_PROBE_T(pt2);
            __syncthreads();
            HB_PROBE_T(pt3);
            if (N >= 4096) {
                if constexpr (kZeroAware) dif_pass_second<N, NT>(slices, count, tw512, s_jlive[gi & 1], tid);
                else dif_pass<N, NT, (N >= 4096 ? N / 64 : 8), false>(slices, count, tw, tw512, grp.fmask, tid);
                // The S = 8 butterfly of thread 8 blk' + t' reads the slots 64 blk' + t' + 8 r', written in the
                // S = 64 pass by the threads 64 (blk' >> 3) + t' + 8 r': both sit in the same 64-thread group, so
                // this hand-over needs a barrier among two warps only (named barriers 1..8), not the whole CTA.
#if HB_BINSUM
            }
            // ---- K3': the quoted bins straight from the S = 64 pass, interpolation, clamp, parity -> price rows ----
            // After the passes down to stride 64 the 64 values that still have to be combined into X_m sit in ONE
            // contiguous block (its base spells the resolved low digits of m) and X_m = sum_t z[t] W_64^{t mu},
            // mu = m >> 3(L-2).  Carr-Madan quotes ~70 consecutive bins: eight lanes per DISTINCT bin each sum eight
            // terms (w8 powers from the constant bank, one W_64 twiddle from the table) and a shuffle tree adds
            // them -- no S = 8 pass, no masks, one CTA barrier less, a third of the FP64 work of that pass.
            const int my_o0 = (k3_g < count) ? grp.s[k3_g].o0 : 0, my_o1 = (k3_g < count) ? grp.s[k3_g].o1 : 0;
            K3Opt cur = k3_load(my_o0, my_o1, 0);
            // item -> (slice, index in the slice's distinct-bin list): the bins of a group's slices numbered consecutively
            int nb0 = grp.s[0].nb, nb1 = count > 1 ? grp.s[1].nb : 0, nb2 = count > 2 ? grp.s[2].nb : 0;
            const int total = nb0 + nb1 + nb2;
            auto item_at = [&](int it, int& g, int& i) -> bool {
                g = 0;
                i = it;
                if (i >= nb0) {
                    i -= nb0;
                    g = 1;
                    if (i >= nb1) {
                        i -= nb1;
                        g = 2;
                    }
                }
                return it < total;
            };
            // full-grid bins of this thread's two items of the first round, loaded ahead of the barrier
            int pre_m[2] = {0, 0};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                int g, i;
                if (item_at((tid >> 3) + u * (NT / 8), g, i)) pre_m[u] = S.bin_m[grp.s[g].b0 + i];
            }
            HB_PROBE_T(pt4);
            __syncthreads();  // every slice of the group has been through the S = 64 pass
            HB_PROBE_T(pt5);
            {
                constexpr int L = Log8<N>::value;
                const int part = tid & 7;
                // two items per thread and round: their load -> FMA -> shuffle chains overlap
#pragma unroll 1
                for (int it0 = 0; it0 < total; it0 += 2 * (NT / 8)) {  // block-uniform trip count (shuffles inside)
                    cplx y[2] = {{0.0, 0.0}, {0.0, 0.0}};
                    int gs[2], is[2], ms[2];
                    bool act[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        act[u] = item_at(it0 + (tid >> 3) + u * (NT / 8), gs[u], is[u]);
                        ms[u] = (it0 == 0) ? pre_m[u] : (act[u] ? S.bin_m[grp.s[gs[u]].b0 + is[u]] : 0);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (act[u]) {
                            const int m = ms[u] & (N - 1);
                            int base = 0;
#pragma unroll
                            for (int d = 0; d < L - 2; ++d) base |= ((m >> (3 * d)) & 7) << (3 * (L - 1 - d));
                            const int mu = m >> (3 * (L - 2));
                            const double2* sl = slices + (size_t)gs[u] * N;
                            HB_ASSERT(gs[u] >= 0 && gs[u] < count && is[u] >= 0 && is[u] < kMaxBins && base >= 0 &&
                                      base + 63 < N && mu >= 0 && mu < 64);
                            cplx acc = {0.0, 0.0};
#pragma unroll
                            for (int a = 0; a < 8; ++a) {
                                const cplx v = lds_c(sl, base + 8 * a + part);
                                const double wc = kW8c[(a * mu) & 7], ws = kW8s[(a * mu) & 7];
                                acc.re = fma(v.re, wc, fma(-v.im, ws, acc.re));
                                acc.im = fma(v.re, ws, fma(v.im, wc, acc.im));
                            }
                            const double2 w = tw512[512 + ((part * mu) & 63)];  // W_64^{part mu}
                            y[u].re = fma(acc.re, w.x, -(acc.im * w.y));
                            if (DECIM) y[u].im = fma(acc.re, w.y, acc.im * w.x);
                        }
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            y[u].re += __shfl_xor_sync(0xffffffffu, y[u].re, o);
                            if (DECIM) y[u].im += __shfl_xor_sync(0xffffffffu, y[u].im, o);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (act[u] && part == 0) {
                            double x = y[u].re;
                            if (DECIM) {  // Re(W_N^{ph m} Y[m mod Nsub]); exact angle reduction in integers
                                double sn, cs;
                                sincospi(-2.0 * (double)(((long long)ph * ms[u]) % gc.n_full) / (double)gc.n_full, &sn, &cs);
                                x = y[u].re * cs - y[u].im * sn;
                            }
                            s_xs[gs[u]][is[u]] = x;
                        }
                    }
                }
            }
            const int max_opt = grp.max_opt;
            const int gg = (k3_g < count) ? k3_g : 0;
            // copied out: after the barrier below the descriptor buffer may be refilled for group gi + 2
            const double s_disc = grp.s[gg].disc, s_fwd = grp.s[gg].fwd;
            const int s_variant = grp.s[gg].variant;
#ifdef HB_CHECK
            const int grp_nb = grp.s[gg].nb;
#endif
            // The slices have been read for the last time: release them (and the threads without an option to
            // finish) to the next group's K1 before the interpolation and the global stores.
            __syncthreads();
            HB_PROBE_T(pt6);
            {
                for (int base = 0; base < max_opt; base += TPG) {
                    if (base > 0) cur = k3_load(my_o0, my_o1, base);
                    if (cur.o >= 0) {
                        double* dst = rows + (size_t)s_variant * n + cur.orig;
                        double price = __longlong_as_double(0x7ff8000000000000LL);
                        HB_ASSERT(cur.orig >= 0 && cur.orig < n && s_variant >= 0 && s_variant < 6);
                        if (cur.bin >= 0) {
                            HB_ASSERT(cur.bin < kMaxBins && (cur.call >> 1) >= 0 && (cur.call >> 1) < kMaxBins &&
                                      cur.bin < grp_nb && (cur.call >> 1) < grp_nb);
                            const double c0 = cur.s0 * s_xs[gg][cur.bin];
                            const double c1 = cur.s1 * s_xs[gg][cur.call >> 1];
                            double call = s_disc * (c0 + (c1 - c0) * cur.frac);
                            if (ph > 0) call += *dst;
                            price = (ph == R - 1) ? finish_price(call, (cur.call & 1) != 0, s_fwd, cur.kdisc) : call;
                        }
                        *dst = price;
                    }
                }
            }
#else
                if (N == 4096 && NT == 512)
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + (tid >> 6)) : "memory");
                else
                    __syncthreads();
            }
            HB_PROBE_T(pt4);
            K3Opt cur = k3_load(grp, 0);
            dif_pass<N, NT, 8, true>(slices, count, tw, tw512, grp.fmask, tid);  // only the digits the strikes need
            HB_PROBE_T(pt5);
            __syncthreads();
            HB_PROBE_T(pt6);
            // ---- K3: last butterfly output at the two bracketing bins, log-strike interpolation
            // (accumulated over phases), clamp, parity -> price rows ----
            {
                const int max_opt = grp.max_opt;
                const int gg = (k3_g < count) ? k3_g : 0;
                // copied out: after the early barrier below the descriptor buffer may be refilled for group gi + 2
                const double s_disc = grp.s[gg].disc, s_fwd = grp.s[gg].fwd;
                const int s_variant = grp.s[gg].variant;
                const double2* sl = slices + (size_t)gg * N;
                if (max_opt <= 0) __syncthreads();
                for (int base = 0; base < max_opt; base += TPG / 2) {  // block-uniform trip count (shuffles inside)
                    const bool last = base + TPG / 2 >= max_opt;
                    K3Opt nxt = cur;
                    if (!last) nxt = k3_load(grp, base + TPG / 2);
                    double x = 0.0;
                    if (cur.bin >= 0) {
                        const int m = cur.bin + k3_half;
                        const cplx y = dif_bin<N, DECIM>(sl, tw512, m & (N - 1));
                        x = y.re;
                        if (DECIM) {  // Re(W_N^{ph m} Y[m mod Nsub]); exact angle reduction in integers
                            double sn, cs;
                            sincospi(-2.0 * (double)(((long long)ph * m) % gc.n_full) / (double)gc.n_full, &sn, &cs);
                            x = y.re * cs - y.im * sn;
                        }
                    }
                    // The slices have been read for the last time: release them (and the 40 % of the CTA that has
                    // no option to finish) to the next group's K1 before the interpolation and the global stores.
                    if (last) __syncthreads();
                    const double x1 = __shfl_xor_sync(0xffffffffu, x, 1);
                    if (cur.o >= 0 && k3_half == 0) {
                        double* dst = rows + (size_t)s_variant * n + cur.orig;
                        double price = __longlong_as_double(0x7ff8000000000000LL);
                        if (cur.bin >= 0) {
                            const double c0 = cur.s0 * x;
                            const double c1 = cur.s1 * x1;
                            double call = s_disc * (c0 + (c1 - c0) * cur.frac);
                            if (ph > 0) call += *dst;
                            price = (ph == R - 1) ? finish_price(call, cur.call != 0, s_fwd, cur.kdisc) : call;
                        }
                        *dst = price;
                    }
                    cur = nxt;
                }
            }
#endif
            HB_PROBE_T(pt7);
            HB_PROBE_ADD(0, pt0, pt1);  // K1 (incl. group set-up of the next group by one thread)
            HB_PROBE_ADD(1, pt1, pt2);  // pass 1
            HB_PROBE_ADD(2, pt2, pt3);  // barrier after pass 1 (waits for the slowest K1)
            HB_PROBE_ADD(3, pt3, pt4);  // pass 2 + pair barrier
            HB_PROBE_ADD(4, pt4, pt5);  // pass 3 (+ first K3 constants)
            HB_PROBE_ADD(5, pt5, pt6);  // barrier after pass 3
            HB_PROBE_ADD(6, pt6, pt7);  // K3 (incl. its release barrier)
            }
        }
        HB_PROBE_T(pf0);
        __syncthreads();  // the price rows of the last group are complete
        if (!split) finalize_job<NT, kFinalizeT<NT>()>(what, rows, S, js, p, out, out2, red, tid);
        HB_PROBE_T(pf1);
        HB_PROBE_ADD(7, pf0, pf1);  // finalize
    }
}

// Finalize for split launches: one CTA per parameter set, rows[p][6][n_opt] in global memory.
template <int NT, int FT = NT>
__global__ void __launch_bounds__(NT)
finalize_rows_kernel(SurfaceDev S, Bounds bd, const double* __restrict__ params, int ld, int P, int what,
                     const double* __restrict__ rows_buf, double* __restrict__ out, double* __restrict__ out2) {
    __shared__ JobState js;
    __shared__ double red[(NT / 32) * 23];
    const int tid = threadIdx.x;
    const int V = (what >= W_NEQ) ? 6 : 1;
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        __syncthreads();
        if (tid == 0) job_setup(js, params, ld, p, bd, V);
        __syncthreads();
        if (!js.valid) {
            invalid_job<NT>(what, S, p, out, out2, tid);
            continue;
        }
        finalize_job<NT, FT>(what, rows_buf + (size_t)p * 6 * S.n_opt, S, js, p, out, out2, red, tid);
    }
}

// ============================================================================================
// Reference-grid ("refgrid") job kernel: heston.cpp:94-151 restated as
//   psi_j = phi(v_j - 1.75 i)/(alpha^2+alpha-v_j^2 + i(2 alpha+1)v_j), v_j = 0.01 j, j=1..1023
//   call  = e^{-alpha ln K}/pi * e^{-rT} * 0.01 * sum_j Re(e^{-i v_j ln K} psi_j)
// The CF is evaluated once per slice instead of once per (strike, j).
// ============================================================================================

constexpr int kRefPoints = 1024;  // heston.cpp:126

// Explicit FMA forms: both refgrid kernels must round identically whatever the compiler would contract.
__device__ __forceinline__ cplx cmul_f(cplx a, cplx b) {
    return {fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re)};
}
__device__ __forceinline__ double re_mac(double acc, cplx t, double2 ps) {  // acc + Re(t * ps)
    return fma(t.re, ps.x, fma(-t.im, ps.y, acc));
}

template <int NT>
__global__ void __launch_bounds__(NT, 2)
refgrid_job_kernel(SurfaceDev S, GridConst gc, Bounds bd, const double* __restrict__ params, int ld, int P, int what,
                   double* __restrict__ out, double* __restrict__ out2, double* __restrict__ scratch, int split,
                   unsigned long long* job_counter) {
    __shared__ double2 psi[kRefPoints];
    __shared__ long long s_job;
    __shared__ JobState js;
    __shared__ double red[(NT / 32) * 23];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = 1;  // W_PRICE / W_LOSS only (launch_jobs sends W_NEQ / W_JAC to refgrid_job6_kernel)
    const int n = S.n_opt, M = S.n_mat;
    const double du = 0.01;  // heston.cpp:127
    const int slices_per_set = V * M;
    const long long n_jobs = split ? (long long)P * slices_per_set : (long long)P;

    auto next_job = [&](long long cur) -> long long {  // dynamic queue, as in fft_job_kernel
        if (!job_counter) return cur + gridDim.x;
        __syncthreads();
        if (tid == 0) s_job = (long long)gridDim.x + (long long)atomicAdd(job_counter, 1ULL);
        __syncthreads();
        return s_job;
    };
    for (long long job = blockIdx.x; job < n_jobs; job = next_job(job)) {
        const int p = split ? (int)(job / slices_per_set) : (int)job;
        const int si_begin = split ? (int)(job % slices_per_set) : 0;
        const int si_end = split ? si_begin + 1 : slices_per_set;
        __syncthreads();
        if (tid == 0) job_setup(js, params, ld, p, bd, V);
        __syncthreads();
        if (!js.valid) {
            if (!split) invalid_job<NT>(what, S, p, out, out2, tid);
            else if (what == W_PRICE && si_begin == 0) invalid_job<NT>(what, S, p, out, out2, tid);
            continue;
        }
        double* rows = (what == W_PRICE) ? out + (size_t)p * n
                                         : scratch + (size_t)(split ? p : (int)blockIdx.x) * 6 * n;
        if (si_begin == 0) {
            for (int v = 0; v < V; ++v)
                for (int i = tid; i < S.n_intr; i += NT) rows[(size_t)v * n + S.intr_orig[i]] = S.intr_val[i];
        }
        // Stage A and the Carr-Madan denominator depend on the grid point only (one variant per job here: the
        // six-variant modes run refgrid_job6_kernel): once per job instead of once per maturity.
        constexpr int PTS = kRefPoints / NT;
        StageA ac[PTS];
        cplx den[PTS];
        {
            const double* x = js.x[0];
            const ClassConst cc = {x[0], x[2] * x[2], x[3] * x[2]};
#pragma unroll 1
            for (int k = 0; k < PTS; ++k) {
                const double v = (double)(tid + k * NT) * du;
                ac[k] = stage_a(cc, v, gc.ui);
                den[k] = cm_inv_denominator(v, gc.alpha);
            }
        }
        for (int si = si_begin; si < si_end; ++si) {
            const int vnt = si / M, mat = si % M;
            const double* x = js.x[vnt];
            const double s2 = x[2] * x[2];
            const double T = S.mat_T[mat];
            const SliceConst sc = {x[0] * x[1] / s2, x[4] / s2, S.ln_spot + (S.rate - S.dividend) * T};
#pragma unroll 1
            for (int k = 0; k < PTS; ++k) {  // ac[] / den[] indexed dynamically: thread-private local memory (L1)
                const int j = tid + k * NT;
                cplx r = {0.0, 0.0};  // integrand(0) == 0, heston.cpp:110
                if (j > 0) {
                    const StageB b = stage_b(ac[k], T);
                    r = cmul(stage_f(b, sc, (double)j * du, gc.ui), den[k]);
                }
                psi[j] = make_double2(r.re, r.im);
            }
            __syncthreads();
            const double disc = S.mat_disc[mat], fwd = S.mat_fwd[mat];
            const int o1 = S.mat_off[mat + 1];
            for (int o = S.mat_off[mat] + warp; o < o1; o += NT / 32) {
                const double k = S.opt_lnk[o];
                // e^{-i v_j k}, j = lane + 32 a: base rotation times a 32-step recurrence
                double sb, cb, ss, cs;
                sincos_nb(-((double)lane * du) * k, &sb, &cb);
                sincos_nb(-(32.0 * du) * k, &ss, &cs);
                cplx twd = {cb, sb};
                const cplx st = {cs, ss};
                double sum = 0.0;
#pragma unroll 4
                for (int a = 0; a < kRefPoints / 32; ++a) {
                    const double2 ps = psi[lane + 32 * a];
                    sum = re_mac(sum, twd, ps);
                    twd = cmul_f(twd, st);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (lane == 0) {
                    const double call = S.opt_s0[o] * disc * (sum * du);  // heston.cpp:136-139
                    rows[(size_t)vnt * n + S.opt_orig[o]] = finish_price(call, S.opt_call[o] != 0, fwd, S.opt_kdisc[o]);
                }
            }
            __syncthreads();
        }
        if (!split) finalize_job<NT>(what, rows, S, js, p, out, out2, red, tid);
    }
}

// --------------------------------------------------------------------------------------------
// Six-variant refgrid kernel (normal equations / Jacobian, one CTA per parameter set).  All six
// variants of one maturity are resident (6 x 16 KiB of psi), so
//   * stage A is computed once per (set, class) and kept in thread-private local memory (4 classes x
//     2 points), stage B once per (class, maturity) -- {base, theta', v0'} share it -- instead of
//     A + B + F per slice: 4 B + 6 F per point and maturity instead of 6 (A + B + F);
//   * the per-strike twiddle recurrence e^{-i v_j ln K} is advanced once and applied to the six
//     psi vectors, and a warp sums two strikes per psi load (12 accumulators per lane).
// Per slice and strike the arithmetic (operations and order) is that of refgrid_job_kernel, so the two
// kernels agree to rounding (a few ulp: the compiler contracts the shared inline routines per call site).
// --------------------------------------------------------------------------------------------

constexpr int kRef6NT = 512;
constexpr size_t kRef6Smem = (size_t)6 * kRefPoints * sizeof(double2);

__global__ void __launch_bounds__(kRef6NT, 1)
refgrid_job6_kernel(SurfaceDev S, GridConst gc, Bounds bd, const double* __restrict__ params, int ld, int P, int what,
                    double* __restrict__ out, double* __restrict__ out2, double* __restrict__ scratch, int split,
                    unsigned long long* job_counter) {
    constexpr int NT = kRef6NT, PTS = kRefPoints / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* psi = reinterpret_cast<double2*>(smem_raw);  // [6][kRefPoints]
    __shared__ long long s_job;
    __shared__ JobState js;
    __shared__ double red[(NT / 32) * 23];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = S.n_opt, M = S.n_mat;
    const double du = 0.01;  // heston.cpp:127
    // class c -> variants sharing its (kappa, sigma, rho): 0 {base, theta', v0'}, kappa', sigma', rho'
    const int cls_variant[4] = {0, 1, 3, 4};

    auto next_job = [&](long long cur) -> long long {  // dynamic queue, as in fft_job_kernel
        if (!job_counter) return cur + gridDim.x;
        __syncthreads();
        if (tid == 0) s_job = (long long)gridDim.x + (long long)atomicAdd(job_counter, 1ULL);
        __syncthreads();
        return s_job;
    };
    // split = one job per (set, maturity) with rows in global memory and a separate finalize kernel
    // (small P: fills the SMs); otherwise one job per set, finalized in place.
    const long long n_jobs = split ? (long long)P * M : (long long)P;
    for (long long job = blockIdx.x; job < n_jobs; job = next_job(job)) {
        const int p = split ? (int)(job / M) : (int)job;
        const int mat_begin = split ? (int)(job % M) : 0, mat_end = split ? mat_begin + 1 : M;
        __syncthreads();
        if (tid == 0) job_setup(js, params, ld, p, bd, 6);
        __syncthreads();
        if (!js.valid) {
            if (!split) invalid_job<NT>(what, S, p, out, out2, tid);
            continue;
        }
        double* rows = scratch + (size_t)(split ? p : (int)blockIdx.x) * 6 * n;
        if (mat_begin == 0) {
            for (int v = 0; v < 6; ++v)
                for (int i = tid; i < S.n_intr; i += NT) rows[(size_t)v * n + S.intr_orig[i]] = S.intr_val[i];
        }
        // stage A per (class, point) and the Carr-Madan denominator per point
        StageA ac[4][PTS];
        cplx den[PTS];
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const int j = tid + k * NT;
            const double v = (double)j * du;
            den[k] = cm_inv_denominator(v, gc.alpha);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double* x = js.x[cls_variant[c]];
                const ClassConst cc = {x[0], x[2] * x[2], x[3] * x[2]};
                ac[c][k] = stage_a(cc, v, gc.ui);
            }
        }
        for (int mat = mat_begin; mat < mat_end; ++mat) {
            const double T = S.mat_T[mat];
            const double lsm = S.ln_spot + (S.rate - S.dividend) * T;
#pragma unroll 1
            for (int k = 0; k < PTS; ++k) {
                const int j = tid + k * NT;
                const double v = (double)j * du;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const StageB b = stage_b(ac[c][k], T);
                    const int nv = (c == 0) ? 3 : 1;
                    for (int i = 0; i < nv; ++i) {
                        const int vnt = (c == 0) ? (i == 0 ? 0 : (i == 1 ? 2 : 5)) : cls_variant[c];
                        const double* x = js.x[vnt];
                        const double s2 = x[2] * x[2];
                        const SliceConst sc = {x[0] * x[1] / s2, x[4] / s2, lsm};
                        cplx r = {0.0, 0.0};  // integrand(0) == 0, heston.cpp:110
                        if (j > 0) r = cmul(stage_f(b, sc, v, gc.ui), den[k]);
                        psi[vnt * kRefPoints + j] = make_double2(r.re, r.im);
                    }
                }
            }
            __syncthreads();
            const double disc = S.mat_disc[mat], fwd = S.mat_fwd[mat];
            const int o0 = S.mat_off[mat], o1 = S.mat_off[mat + 1];
            // a warp sums two strikes at a time over the six resident psi vectors
            for (int ob = o0 + 2 * warp; ob < o1; ob += 2 * (NT / 32)) {
                const bool two = ob + 1 < o1;
                const double k0 = S.opt_lnk[ob], k1 = S.opt_lnk[two ? ob + 1 : ob];
                // e^{-i v_j k}, j = lane + 32 a: base rotation times a 32-step recurrence
                double sb, cb, ss, cs;
                sincos_nb(-((double)lane * du) * k0, &sb, &cb);
                sincos_nb(-(32.0 * du) * k0, &ss, &cs);
                cplx tw0 = {cb, sb};
                const cplx st0 = {cs, ss};
                sincos_nb(-((double)lane * du) * k1, &sb, &cb);
                sincos_nb(-(32.0 * du) * k1, &ss, &cs);
                cplx tw1 = {cb, sb};
                const cplx st1 = {cs, ss};
                double sum0[6], sum1[6];
#pragma unroll
                for (int v = 0; v < 6; ++v) sum0[v] = sum1[v] = 0.0;
#pragma unroll 2
                for (int a = 0; a < kRefPoints / 32; ++a) {
#pragma unroll
                    for (int v = 0; v < 6; ++v) {
                        const double2 ps = psi[v * kRefPoints + lane + 32 * a];
                        sum0[v] = re_mac(sum0[v], tw0, ps);
                        sum1[v] = re_mac(sum1[v], tw1, ps);
                    }
                    tw0 = cmul_f(tw0, st0);
                    tw1 = cmul_f(tw1, st1);
                }
#pragma unroll
                for (int v = 0; v < 6; ++v) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        sum0[v] += __shfl_xor_sync(0xffffffffu, sum0[v], off);
                        sum1[v] += __shfl_xor_sync(0xffffffffu, sum1[v], off);
                    }
                }
                if (lane < 12 && (lane < 6 || two)) {
                    const int v = lane % 6, o = ob + lane / 6;
                    double sum = 0.0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) sum = (v == i) ? (lane < 6 ? sum0[i] : sum1[i]) : sum;
                    const double call = S.opt_s0[o] * disc * (sum * du);  // heston.cpp:136-139
                    rows[(size_t)v * n + S.opt_orig[o]] = finish_price(call, S.opt_call[o] != 0, fwd, S.opt_kdisc[o]);
                }
            }
            __syncthreads();
        }
        if (!split) finalize_job<NT>(what, rows, S, js, p, out, out2, red, tid);
    }
}

// ============================================================================================
// Black-Scholes implied volatility of model prices: HestonModel::implied_volatility,
// heston.cpp:311-349 (Newton from sqrt(v0), tol 1e-8, <= 100 steps, vol in [0.001, 5],
// vega < 1e-12 -> vol *= 1.5), with black_scholes_price / _vega of heston.cpp:275-309.
// One thread per (parameter set, option); prices come from hb_price.  Not a hot path: libdevice math.
// ============================================================================================

__device__ __forceinline__ double bs_norm_cdf(double x) { return 0.5 * (1.0 + erf(x / sqrt(2.0))); }  // :16-18

__global__ void implied_vol_kernel(const double* __restrict__ prices, const double* __restrict__ params, int ld, int P,
                                   int n, const double* __restrict__ strike, const double* __restrict__ maturity,
                                   const uint8_t* __restrict__ is_call, double spot, double rate, double dividend,
                                   double* __restrict__ out) {
    const size_t total = (size_t)P * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), p = (int)(idx / n);
        const double target = prices[idx], K = strike[i], T = maturity[i];
        const bool call = is_call[i] != 0;
        double vol = sqrt(params[(size_t)4 * ld + p]);  // sqrt(v0)
        if (target != target) {
            out[idx] = target;  // invalid option / parameter set
            continue;
        }
        if (T <= 0.0) {  // :316-318
            out[idx] = 0.0;
            continue;
        }
        const double fwd = spot * exp((rate - dividend) * T), disc = exp(-rate * T), sq = sqrt(T);
        const double sdiv = spot * exp(-dividend * T);
        for (int it = 0; it < 100; ++it) {
            const double vs = vol * sq;
            const double d1 = (log(fwd / K) + 0.5 * vol * vol * T) / vs, d2 = d1 - vs;
            const double bs = call ? sdiv * bs_norm_cdf(d1) - K * disc * bs_norm_cdf(d2)
                                   : K * disc * bs_norm_cdf(-d2) - sdiv * bs_norm_cdf(-d1);
            const double vega = (vol <= 0.0) ? 0.0 : sdiv * sq * 0.3989422804014327 * exp(-0.5 * d1 * d1);
            if (vega < 1e-12) {
                vol *= 1.5;
                continue;
            }
            const double diff = bs - target;
            if (fabs(diff) < 1e-8) break;
            vol = fmax(0.001, fmin(5.0, vol - diff / vega));
        }
        out[idx] = vol;
    }
}

// ============================================================================================
// Finite-difference Greeks: HestonModel::price_option_with_greeks, heston.cpp:168-217.  The nine price
// evaluations per option are nine launches of the pricing kernel on bumped surfaces / parameter sets;
// these two kernels build the v0 +/- eps parameter sets and apply the difference formulas.
// ============================================================================================

constexpr double kEpsSpotRel = 0.001, kEpsRate = 0.0001, kEpsTime = 1.0 / 365.0, kEpsVol = 0.001;  // :175-178

// out[0] = params with v0 + eps, out[1] = params with v0 - eps, both SoA [5][P]
__global__ void bump_v0_kernel(const double* __restrict__ params, int ld, int P, double* __restrict__ out) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const double x = params[(size_t)c * ld + p];
            out[(size_t)c * P + p] = (c == 4) ? x + kEpsVol : x;
            out[(size_t)(5 + c) * P + p] = (c == 4) ? x - kEpsVol : x;
        }
    }
}

// pr = 8 price planes [8][P*n]: base, S+, S-, r+, r-, T-, v0+, v0-
__global__ void greeks_combine_kernel(const double* __restrict__ pr, size_t plane, int n,
                                      const double* __restrict__ maturity, double spot, double* __restrict__ out) {
    const double eps_spot = spot * kEpsSpotRel;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (size_t)gridDim.x * blockDim.x) {
        const double price = pr[idx], up = pr[plane + idx], dn = pr[2 * plane + idx];
        const double rup = pr[3 * plane + idx], rdn = pr[4 * plane + idx], later = pr[5 * plane + idx];
        const double vup = pr[6 * plane + idx], vdn = pr[7 * plane + idx];
        double* g = out + idx * 5;
        g[0] = (up - dn) / (2.0 * eps_spot);                       // delta  :183
        g[1] = (up - 2.0 * price + dn) / (eps_spot * eps_spot);    // gamma  :186
        g[2] = (vup - vdn) / (2.0 * kEpsVol);                      // vega   :213
        g[3] = (maturity[idx % n] > kEpsTime) ? (later - price) / kEpsTime : 0.0;  // theta :195-200
        g[4] = (rup - rdn) / (2.0 * kEpsRate);                     // rho    :192
    }
}

// ============================================================================================
// K1 alone: characteristic function for arbitrary complex u
// ============================================================================================

__global__ void cf_kernel(const double* __restrict__ params, int ld, int P, const double* __restrict__ T, int n_T,
                          const double* __restrict__ ur, const double* __restrict__ ui, int n_u, double spot,
                          double rate, double dividend, double2* __restrict__ out) {
    const size_t total = (size_t)P * n_T * n_u;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % n_u);
        const size_t pm = idx / n_u;
        const int m = (int)(pm % n_T);
        const int p = (int)(pm / n_T);
        const cplx z = heston_cf(params[p], params[(size_t)ld + p], params[(size_t)2 * ld + p],
                                 params[(size_t)3 * ld + p], params[(size_t)4 * ld + p], ur[j], ui[j], T[m], spot,
                                 rate, dividend);
        out[idx] = make_double2(z.re, z.im);
    }
}

// ============================================================================================
// K2 alone: batched FFT, slices staged HBM -> shared memory with cp.async.bulk (TMA engine,
// SASS UBLKCP) behind an mbarrier, 3-deep ring so the copy of slice s+2 overlaps the
// transform of slice s.
// ============================================================================================

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <int N, int NT, int STAGES>
__global__ void __launch_bounds__(NT, 1) fft_batch_kernel(double2* __restrict__ data, int n_slices) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* tw = bufs + (size_t)STAGES * N;
    __shared__ __align__(8) uint64_t bar[STAGES];
    const int tid = threadIdx.x;
    constexpr uint32_t kBytes = (uint32_t)N * sizeof(double2);

    fill_twiddles<N>(tw, tid, NT);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_mine = (first < n_slices) ? (n_slices - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int it = 0; it < STAGES - 1 && it < n_mine; ++it) {
            mbar_expect_tx(&bar[it], kBytes);
            bulk_g2s(bufs + (size_t)it * N, data + (size_t)(first + it * stride) * N, kBytes, &bar[it]);
        }
    }
    for (int it = 0; it < n_mine; ++it) {
        const int st = it % STAGES;
        // refill the stage consumed in the previous iteration with slice it + STAGES - 1
        if (tid == 0 && it + STAGES - 1 < n_mine) {
            const int nx = (it + STAGES - 1) % STAGES;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar[nx], kBytes);
            bulk_g2s(bufs + (size_t)nx * N, data + (size_t)(first + (it + STAGES - 1) * stride) * N, kBytes,
                     &bar[nx]);
        }
        mbar_wait(&bar[st], (uint32_t)((it / STAGES) & 1));
        double2* buf = bufs + (size_t)st * N;
        fft_pass<N, NT, true>(buf, 1, 1, tw, 1, tid);
        for (int Ns = 8; Ns < N; Ns *= 8) fft_pass<N, NT, false>(buf, 1, 1, tw, Ns, tid);
        double2* dst = data + (size_t)(first + it * stride) * N;
        for (int i = tid; i < N; i += NT) dst[i] = buf[swz(i)];
        __syncthreads();  // stage st fully read before it is refilled next iteration
    }
}

// ============================================================================================
// FP64 pipe probe: 8 independent DFMA chains per thread
// ============================================================================================

__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace hb
