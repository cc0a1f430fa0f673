/*
 * heston_b200.h -- C ABI of libheston_b200.so: batched Heston Carr-Madan pricing and
 * calibration objective on NVIDIA B200 (sm_100a), FP64.
 *
 * This is the drop-in boundary for ONE path of dharvpat/PDE: what the pybind11 module
 * `quant_cpp.heston` (src/cpp/bindings/heston_bindings.cpp:15-254) and the calibrator's
 * pricing loop (src/python/quant_trading/calibration/heston_calibrator.py:486-586) do on
 * the CPU.  Plain pointers and sizes only; no torch / pybind types.  All citations are
 * file:line in the reference repository.
 *
 * Conventions
 *   - Every function returns an hb_status; hb_last_error() gives a thread-local message.
 *   - "d_" pointers are DEVICE pointers owned by the caller (e.g. torch.Tensor.data_ptr());
 *     "h_" / unprefixed const double* in *_host and hb_model_* functions are HOST pointers.
 *   - Parameter sets on the device are SoA: d_params[c * ld + p], c in {kappa, theta, sigma,
 *     rho, v0}, p < P, ld >= P (reference: AoS struct HestonParameters, heston.hpp:42-109).
 *     Host entry points take AoS double[P][5] (what numpy / std::vector<HestonParameters> hold).
 *   - Launches are asynchronous on `stream` (a cudaStream_t, NULL = legacy default stream);
 *     *_host functions synchronise before returning.
 *   - Batched calls never fail on an invalid parameter set or option: that set/option yields
 *     NaN prices and the objective's 1e10 sentinel (heston_calibrator.py:507-508, :583-584).
 *     The scalar hb_model_* functions return HB_ERR_INVALID_* with the reference's message
 *     (std::invalid_argument -> ValueError, heston.cpp:28-35, :156-164).
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     HB_ERR_CUDA.
 *   - Threading: a plan owns its workspace (price rows, job counter, staging buffers), so the calls on
 *     ONE plan must be stream-ordered -- one stream at a time, or externally serialised; different plans
 *     are independent and may be driven from different host threads and streams concurrently (the
 *     reference's model objects are const and share nothing either, heston.hpp:117-258).
 */
#ifndef HESTON_B200_H
#define HESTON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_plan hb_plan;

typedef enum {
    HB_OK = 0,
    HB_ERR_INVALID_ARGUMENT = 1,  /* bad strike/spot/maturity/size: heston.cpp:156-164, :229-231 */
    HB_ERR_INVALID_PARAMETER = 2, /* HestonParameters::validate failed: heston.hpp:81-100 */
    HB_ERR_CUDA = 3,              /* CUDA runtime error or no device */
    HB_ERR_STATE = 4              /* call order (e.g. no surface set) or unsupported configuration */
} hb_status;

typedef enum {
    /* The reference's actual arithmetic: 1023-point direct quadrature of the damped
     * integrand, dv = 0.01, alpha = 0.75 (price_option_integration, heston.cpp:94-151). */
    HB_MODE_REFGRID = 0,
    /* Carr-Madan N-point FFT + linear log-strike interpolation
     * (docs/models/heston-model.md:89-106; no reference code -- SURVEY.md F1). */
    HB_MODE_FFT = 1
} hb_mode;

/* Width of one hb_normal_eq row: loss, ||r||^2, J^T r (5), upper triangle of J^T J (15). */
#define HB_NEQ_WIDTH 22

int hb_version(void);               /* 100 * major + minor */
const char* hb_last_error(void);    /* thread-local, never NULL */
int hb_device_count(void);          /* 0 when no CUDA device is usable */

/* ---- plan: grid constants, tables, workspace ------------------------------------------ */

/* mode FFT: n_grid = 512 or a power of two in [4096, 65536] (grids above 4096 are transformed as
 * R = n_grid/4096 decimated on-chip sub-transforms), eta > 0, alpha > 0 (reference docs: 4096, 0.25; alpha
 * 0.75 from heston.hpp:261).  mode REFGRID: n_grid/eta are ignored (1024 / 0.01 fixed by
 * heston.cpp:126-127), alpha must be 0.75.  `device` is a CUDA ordinal. */
int hb_plan_create(int mode, int n_grid, double eta, double alpha, int device, hb_plan** out);
int hb_plan_destroy(hb_plan* plan);

/* Flat option list exactly as HestonCalibrator passes it (heston_calibrator.py:293-303):
 * strike[i], maturity[i], is_call[i] (0/1), market[i] (may be NULL when only prices are
 * wanted).  HOST pointers; the plan groups options by distinct maturity, builds interpolation
 * tables and uploads them.  Options with strike<=0 or maturity<0 (or spot<=0) price to NaN;
 * maturity==0 prices to intrinsic value (heston.cpp:97-100). */
int hb_surface_set(hb_plan* plan, int n_opt, const double* strike, const double* maturity, const uint8_t* is_call,
                   const double* market, double spot, double rate, double dividend);

/* Box used by the finite-difference step rule (scipy least_squares 2-point with bounds,
 * heston_calibrator.py:465-477).  Default: the calibrator's DEFAULT_BOUNDS (:201-207). */
int hb_set_bounds(hb_plan* plan, const double* lb5, const double* ub5);

/* Significance cut of the Carr-Madan (FFT) modes.  The damped integrand decays exponentially in v; a grid
 * point whose |phi| lies below e^cut is treated as exactly 0 (as the reference's own double arithmetic does
 * once exp underflows, cut = -746).  cut is chosen per surface so that ALL dropped points together cannot move
 * any price by more than `abs_price_error`:  e^cut * sum_j |w_j / denominator_j| * max_options(e^{-rT}
 * e^{-alpha k}/pi) = abs_price_error.  Default 2^-80 (8.3e-25, twelve orders below the 1e-12 absolute parity
 * tolerance); 0 = exact mode (only true underflow is skipped); values above 1e-12 are rejected.  REFGRID mode
 * restates the reference's arithmetic and ignores it.  hb_plan_log_cut returns the cut in force (after
 * hb_surface_set). */
int hb_plan_set_truncation(hb_plan* plan, double abs_price_error);
double hb_plan_log_cut(const hb_plan* plan);

int hb_plan_n_options(const hb_plan* plan);
int hb_plan_n_maturities(const hb_plan* plan); /* distinct maturities > 0 = slices per parameter set */

/* ---- batched device entry points -------------------------------------------------------- */

/* d_prices[p * n_opt + i]: replaces HestonCalibrator._price_options (heston_calibrator.py:538-586)
 * for P parameter sets at once. */
int hb_price(hb_plan* plan, const double* d_params, int ld, int P, double* d_prices, void* stream);

/* d_loss[p] = sum(((model - market)/market)^2), or 1e10 if any price is NaN or <= 0:
 * replaces _compute_objective (heston_calibrator.py:486-513). */
int hb_objective(hb_plan* plan, const double* d_params, int ld, int P, double* d_loss, void* stream);

/* d_out[p * 22 + ...] = { loss, ||r||^2, J^T r [5], triu(J^T J) [15] } with r from
 * _compute_residuals (heston_calibrator.py:515-536) and J the forward-difference Jacobian
 * scipy.optimize.least_squares(jac='2-point', bounds) builds from it. */
int hb_normal_eq(hb_plan* plan, const double* d_params, int ld, int P, double* d_out, void* stream);

/* Full residual vectors d_res[p * n_opt + i] and Jacobians d_jac[(p * n_opt + i) * 5 + j]
 * (for small P: one Levenberg-Marquardt iterate). */
int hb_jacobian(hb_plan* plan, const double* d_params, int ld, int P, double* d_res, double* d_jac, void* stream);

/* Black-Scholes implied volatilities of the model prices, d_iv[p * n_opt + i]: replaces a loop of
 * HestonModel::implied_volatility (heston.cpp:311-349; models/heston.py:313-343 builds surfaces from it).
 * Same Newton iteration, start and clamps as the reference. */
int hb_implied_vol(hb_plan* plan, const double* d_params, int ld, int P, double* d_iv, void* stream);

/* Finite-difference Greeks of every (parameter set, option), d_greeks[(p * n_opt + i) * 5 + {delta, gamma,
 * vega, theta, rho}]: replaces a loop of HestonModel::price_option_with_greeks (heston.cpp:168-217) -- the
 * same nine price evaluations per option (spot +/- 0.1 %, rate +/- 1e-4, maturity - 1/365, v0 +/- 1e-3)
 * and the same difference formulas; theta is 0 for maturity <= 1/365.  A set whose v0 - 1e-3 is not
 * positive gets NaN vega (the reference's constructor throws there). */
int hb_greeks(hb_plan* plan, const double* d_params, int ld, int P, double* d_greeks, void* stream);

/* Characteristic function phi(u_j; T_m) for every parameter set:
 * d_out[((p * n_T + m) * n_u + j) * 2 + {0,1}]; replaces HestonModel::characteristic_function
 * (heston.cpp:74-92).  d_T, d_ur, d_ui are device arrays. */
int hb_cf(const double* d_params, int ld, int P, const double* d_T, int n_T, const double* d_ur, const double* d_ui,
          int n_u, double spot, double rate, double dividend, double* d_out, void* stream);

/* Batched in-shared-memory FFT of n_slices complex128 slices of length n (512 or 4096),
 * forward sign, in place on the device (bulk-async staged).  Exposed for parity tests of
 * the transform stage and for the unfused pipeline. */
int hb_fft_batch(double* d_data, int n, int n_slices, void* stream);

int hb_sync(void* stream);

/* ---- SABR (Hagan 2002) smile path: the sibling hot loop ------------------------
 * The reference holds two formulas that differ in their guards; `flavour` selects which is restated:
 *   HB_SABR_CPP  SABRModel::implied_volatility (src/cpp/models/sabr.cpp:130-192; bound as
 *                quant_cpp.sabr.SABRModel.implied_volatility / implied_volatilities, bindings/sabr_bindings.cpp);
 *                NaN where the reference throws std::invalid_argument
 *   HB_SABR_PY   SABRCalibrator.sabr_implied_vol (calibration/sabr_calibrator.py:159-258), what the
 *                calibration objective evaluates
 * Parameter sets are SoA: d_params[c * ld + p], c = alpha, rho, nu; beta is fixed per call (sabr.cpp:19-32). */
#define HB_SABR_CPP 0
#define HB_SABR_PY 1

/* d_vols[p * n + i] for P parameter sets on one smile (forward, maturity, n strikes): replaces a loop of
 * implied_volatility / SABRModel::implied_volatilities (sabr.cpp:194-207). */
int hb_sabr_vols(int flavour, double beta, double forward, double maturity, int n, const double* d_strikes,
                 const double* d_params, int ld, int P, double* d_vols, void* stream);

/* Calibration objective of calibrate_single_maturity (sabr_calibrator.py:316-324) for every (smile, candidate):
 * d_loss[m * P + p] = sum_i w_i (sigma_i - market_i)^2, smiles in CSR form (d_off[n_smiles + 1] into d_strikes /
 * d_market / d_weights, weights already normalised as in :291-293; at most max_strikes <= 512 per smile),
 * candidates d_params[(m * 3 + c) * ld + p].  One launch for all maturities x candidates instead of one
 * Python formula evaluation per (candidate, strike). */
int hb_sabr_objective(double beta, int n_smiles, const double* d_forward, const double* d_maturity, const int* d_off,
                      int max_strikes, const double* d_strikes, const double* d_market, const double* d_weights,
                      const double* d_params, int ld, int P, double* d_loss, void* stream);

/* Host-pointer variant of hb_sabr_vols: h_params AoS double[P][3]; copies included. */
int hb_sabr_vols_host(int flavour, double beta, double forward, double maturity, int n, const double* h_strikes,
                      const double* h_params, int P, double* h_vols);

/* ---- host-pointer entry points (what a pybind11 / ctypes binding calls) --------------------
 * h_params is AoS double[P][5].  Each call copies inputs to the device, runs the same kernels
 * as above and copies the result back before returning. */
int hb_price_host(hb_plan* plan, const double* h_params, int P, double* h_prices);
int hb_objective_host(hb_plan* plan, const double* h_params, int P, double* h_loss);
int hb_implied_vol_host(hb_plan* plan, const double* h_params, int P, double* h_iv);
int hb_greeks_host(hb_plan* plan, const double* h_params, int P, double* h_greeks);
int hb_normal_eq_host(hb_plan* plan, const double* h_params, int P, double* h_out);
int hb_jacobian_host(hb_plan* plan, const double* h_params, int P, double* h_res, double* h_jac);

/* ---- scalar drop-ins for quant_cpp.heston.HestonModel -----------------------------------------
 * params5 = {kappa, theta, sigma, rho, v0}.  Reference-faithful ("refgrid") arithmetic. */

/* HestonParameters::validate (heston.hpp:81-100). */
int hb_model_validate(const double* params5);
/* HestonModel::characteristic_function (heston.cpp:74-92), out2 = {re, im}. */
int hb_model_cf(const double* params5, double u_re, double u_im, double T, double spot, double rate, double dividend,
                double* out2, int device);
/* HestonModel::price_options (heston.cpp:220-245): n_maturity must be 1 or n. */
int hb_model_price_options(const double* params5, int n, const double* strikes, int n_maturity,
                           const double* maturities, double spot, double rate, double dividend, int is_call,
                           double* out, int device);

/* ---- measurement helpers ------------------------------------------------------------------------ */

/* Sustained DFMA throughput of the device in TFLOP/s (2 flops per DFMA): the roofline
 * denominator of this FP64-pipe-bound path (MEASURED_PEAKS.json has no FP64 entry). */
int hb_measure_fp64_peak(int device, double seconds, double* tflops);
/* Kernel launches issued by this library in this process so far. */
uint64_t hb_launch_count(void);
/* Kernel durations of a plan's pricing launches, measured with CUDA events on the launch stream (bench.py's roofline:
 * the duration of the dominant kernel inside the timed region).  enable != 0: every hb_price / hb_objective /
 * hb_normal_eq / hb_jacobian launch records an event pair around each of its kernels (at most 1024 pairs are kept). */
int hb_plan_profile(hb_plan* plan, int enable);
/* Waits for the recorded launches, then ms3 = summed durations of { prefix scan, direct-sum job kernel, transform /
 * refgrid job kernel (with its finalize) } and n3 = { kernel invocations recorded, parameter sets the LAST launch
 * routed to the direct-sum kernel, ... to the transform kernel } (-1: that launch was not routed); clears the record. */
int hb_plan_profile_read(hb_plan* plan, double* ms3, long long* n3);

#ifdef __cplusplus
}
#endif
#endif /* HESTON_B200_H */
