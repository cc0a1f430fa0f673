// sabr_b200.cu -- batched SABR (Hagan 2002) implied volatilities and smile-calibration objective for sm_100a,
// behind the C ABI of include/heston_b200.h (section "SABR").  SURVEY.md 8f rank 4: the sibling hot loop of
// the Heston objective -- closed form, one thread per (smile, candidate).
//
// The reference holds two formulas that differ in their guards; both are restated with the reference's
// operation order:
//   HB_SABR_CPP  SABRModel::implied_volatility          src/cpp/models/sabr.cpp:34-192
//   HB_SABR_PY   SABRCalibrator.sabr_implied_vol        src/python/quant_trading/calibration/sabr_calibrator.py:159-258
//                and the objective of calibrate_single_maturity, :316-324
// This translation unit is compiled with -fmad=false: near the money the formula takes log(1 + eps) of an
// argument formed by a few additions, so a contracted FMA in that argument would move the volatility by
// up to 1e-6 relative against the reference (its own conditioning); with IEEE mul/add/div/sqrt the only
// differences left are libdevice's log and pow (<= 2 ulp).
// No CPU code path: every entry point fails without a CUDA device.
#include "../../include/heston_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

extern "C" int hb_internal_fail(int code, const char* msg);  // heston_b200.cu: sets hb_last_error()
extern "C" void hb_internal_count_launch(void);

namespace {

constexpr double kEps = 1e-10;           // sabr.cpp:14
constexpr double kAtmThreshold = 1e-6;   // sabr.cpp:17

#define SB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return hb_internal_fail(HB_ERR_CUDA, (std::string(#expr) + ": " + cudaGetErrorString(_e)).c_str()); \
    } while (0)

// ---- SABRModel::implied_volatility, sabr.cpp:130-192 (NaN where the reference throws) -----------------

__device__ double chi_function(double z, double rho) {  // sabr.cpp:34-60
    if (fabs(z) < kEps) return z * (1.0 + 0.5 * rho * z + (2.0 * rho * rho - 1.0) / 6.0 * z * z);
    const double sqrt_term = sqrt(1.0 - 2.0 * rho * z + z * z);
    double numerator = sqrt_term + z - rho;
    const double denominator = 1.0 - rho;
    if (fabs(denominator) < kEps) return (z < 1.0) ? z / (1.0 - z) : z / (z - 1.0);
    if (numerator <= 0.0) numerator = kEps;
    return log(numerator / denominator);
}

__device__ double sabr_vol_cpp(double beta, double K, double F, double T, double alpha, double rho, double nu) {
    if (!(K > 0.0) || !(F > 0.0) || !(alpha > 0.0) || !(fabs(rho) < 1.0) || !(nu >= 0.0) || !(T >= 0.0))
        return __longlong_as_double(0x7ff8000000000000LL);
    const double omb = 1.0 - beta;
    if (T < kEps) return alpha / pow(sqrt(F * K), 1.0 - beta);  // :152-155
    const double log_fk = log(F / K);
    if (fabs(log_fk) < kAtmThreshold) {  // atm_volatility, sabr.cpp:95-128
        const double f_power = pow(F, omb);
        const double sigma_atm = alpha / f_power;
        const double term1 = (omb * omb / 24.0) * alpha * alpha / (f_power * f_power);
        const double term2 = (rho * beta * nu * alpha) / (4.0 * f_power);
        const double term3 = ((2.0 - 3.0 * rho * rho) / 24.0) * nu * nu;
        return sigma_atm * (1.0 + (term1 + term2 + term3) * T);
    }
    const double fk_mid = sqrt(F * K);
    const double fk_power = pow(fk_mid, omb);
    const double z = (nu < kEps || alpha < kEps) ? 0.0 : (nu / alpha) * pow(fk_mid, 1.0 - beta) * log_fk;  // :62-74
    const double chi_z = chi_function(z, rho);
    const double z_over_chi = (fabs(z) < kEps) ? 1.0 : z / chi_z;
    const double l2 = log_fk * log_fk;
    const double numerator_correction = 1.0 + (omb * omb / 24.0) * l2 + (pow(omb, 4.0) / 1920.0) * l2 * l2;
    const double denominator = fk_power * numerator_correction;
    const double sigma_base = (alpha / denominator) * z_over_chi;
    // compute_correction_factor, sabr.cpp:76-93
    const double term1 = (omb * omb / 24.0) * (alpha * alpha) / (fk_power * fk_power);
    const double term2 = (rho * beta * nu * alpha) / (4.0 * fk_power);
    const double term3 = ((2.0 - 3.0 * rho * rho) / 24.0) * nu * nu;
    return sigma_base * (1.0 + (term1 + term2 + term3) * T);
}

// ---- SABRCalibrator.sabr_implied_vol, sabr_calibrator.py:159-258 -------------------------------------
// The candidate-independent pieces of a strike (log(F/K), (FK)^((1-beta)/2), the denominator correction,
// F^(1-beta) for the ATM branch) are formed once per (smile, strike) with the reference's operations.
struct StrikeConst {
    double log_FK, FK_beta, denom_term, F_beta;
    int atm;
};

__device__ StrikeConst strike_const(double beta, double K, double F) {
    StrikeConst c;
    c.atm = fabs(F - K) < 1e-10;  // :184
    c.F_beta = pow(F, 1.0 - beta);  // :246
    const double FK = F * K;
    c.log_FK = log(F / K);
    c.FK_beta = pow(FK, (1.0 - beta) / 2.0);
    const double omb = 1.0 - beta, omb2 = omb * omb;
    double denom_term = 1.0 + omb2 / 24.0 * c.log_FK * c.log_FK;     // :216
    denom_term += omb2 * omb2 / 1920.0 * pow(c.log_FK, 4.0);          // :217
    c.denom_term = denom_term;
    return c;
}

__device__ double sabr_vol_py(const StrikeConst& c, double beta, double T, double alpha, double rho, double nu) {
    const double omb = 1.0 - beta, omb2 = omb * omb;
    if (c.atm) {  // _sabr_atm_vol, :226-258
        const double term1 = omb * omb / 24.0 * alpha * alpha / (c.F_beta * c.F_beta);
        const double term2 = rho * beta * nu * alpha / (4.0 * c.F_beta);
        const double term3 = (2.0 - 3.0 * rho * rho) * nu * nu / 24.0;
        return alpha / c.F_beta * (1.0 + (term1 + term2 + term3) * T);
    }
    const double z = (nu / alpha) * c.FK_beta * c.log_FK;                 // :193
    const double sqrt_term = sqrt(1.0 - 2.0 * rho * z + z * z);           // :196
    const double x_z = log((sqrt_term + z - rho) / (1.0 - rho));          // :197
    const double zeta = (fabs(x_z) < 1e-10) ? 1.0 : z / x_z;              // :200-203
    const double term1 = omb2 / 24.0 * alpha * alpha / (c.FK_beta * c.FK_beta);
    const double term2 = rho * beta * nu * alpha / (4.0 * c.FK_beta);
    const double term3 = (2.0 - 3.0 * rho * rho) * nu * nu / 24.0;
    const double bracket = 1.0 + (term1 + term2 + term3) * T;
    const double sigma = (alpha / (c.FK_beta * c.denom_term)) * zeta * bracket;
    return (sigma > 1e-6) ? sigma : ((sigma != sigma) ? sigma : 1e-6);  // max(sigma, 1e-6), NaN kept
}

__global__ void sabr_vols_kernel(int flavour, double beta, double F, double T, int n, const double* __restrict__ K,
                                 const double* __restrict__ params, int ld, int P, double* __restrict__ vols) {
    const size_t total = (size_t)P * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), p = (int)(idx / n);
        const double alpha = params[p], rho = params[(size_t)ld + p], nu = params[(size_t)2 * ld + p];
        vols[idx] = (flavour == HB_SABR_CPP) ? sabr_vol_cpp(beta, K[i], F, T, alpha, rho, nu)
                                             : sabr_vol_py(strike_const(beta, K[i], F), beta, T, alpha, rho, nu);
    }
}

// One CTA row per smile (blockIdx.y), candidates along x; the smile's strike constants, market vols and
// weights are staged in shared memory once per CTA.
constexpr int kObjThreads = 128;
constexpr int kMaxSmileStrikes = 512;

__global__ void __launch_bounds__(kObjThreads)
sabr_objective_kernel(double beta, const double* __restrict__ forward, const double* __restrict__ maturity,
                      const int* __restrict__ off, const double* __restrict__ strikes,
                      const double* __restrict__ market, const double* __restrict__ weights,
                      const double* __restrict__ params, int ld, int P, double* __restrict__ loss) {
    __shared__ StrikeConst sc[kMaxSmileStrikes];
    __shared__ double smk[kMaxSmileStrikes], sw[kMaxSmileStrikes];
    const int m = blockIdx.y;
    const int o0 = off[m], n = off[m + 1] - o0;
    const double F = forward[m], T = maturity[m];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sc[i] = strike_const(beta, strikes[o0 + i], F);
        smk[i] = market[o0 + i];
        sw[i] = weights[o0 + i];
    }
    __syncthreads();
    const double* pm = params + (size_t)m * 3 * ld;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const double alpha = pm[p], rho = pm[(size_t)ld + p], nu = pm[(size_t)2 * ld + p];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {  // np.sum(weights * errors), :322-324 (sequential order)
            const double e = sabr_vol_py(sc[i], beta, T, alpha, rho, nu) - smk[i];
            s += sw[i] * (e * e);
        }
        loss[(size_t)m * P + p] = s;
    }
}

int check_beta(double beta) {
    if (!(beta >= 0.0 && beta <= 1.0)) {  // sabr.cpp:21-25
        char buf[96];
        std::snprintf(buf, sizeof buf, "SABR: beta must be in [0, 1], got %f", beta);
        return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, buf);
    }
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_sabr_vols(int flavour, double beta, double forward, double maturity, int n, const double* d_strikes,
                 const double* d_params, int ld, int P, double* d_vols, void* stream) {
    if (flavour != HB_SABR_CPP && flavour != HB_SABR_PY) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "unknown SABR flavour");
    int rc = check_beta(beta);
    if (rc) return rc;
    if (n < 0 || P < 0 || ld < P) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "need n >= 0 and 0 <= P <= ld");
    if (n == 0 || P == 0) return HB_OK;
    if (!d_strikes || !d_params || !d_vols) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "NULL device pointer");
    int dev = 0, sms = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t total = (size_t)P * n;
    const int grid = (int)std::min<size_t>((total + 127) / 128, (size_t)sms * 16);
    sabr_vols_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(flavour, beta, forward, maturity, n, d_strikes, d_params, ld,
                                                            P, d_vols);
    hb_internal_count_launch();
    SB_CUDA(cudaGetLastError());
    return HB_OK;
}

int hb_sabr_objective(double beta, int n_smiles, const double* d_forward, const double* d_maturity, const int* d_off,
                      int max_strikes, const double* d_strikes, const double* d_market, const double* d_weights,
                      const double* d_params, int ld, int P, double* d_loss, void* stream) {
    int rc = check_beta(beta);
    if (rc) return rc;
    if (n_smiles < 0 || P < 0 || ld < P) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "need n_smiles >= 0 and 0 <= P <= ld");
    if (max_strikes < 0 || max_strikes > kMaxSmileStrikes)
        return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "a smile holds at most 512 strikes");
    if (n_smiles == 0 || P == 0) return HB_OK;
    if (!d_forward || !d_maturity || !d_off || !d_strikes || !d_market || !d_weights || !d_params || !d_loss)
        return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "NULL device pointer");
    int dev = 0, sms = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int gx = std::max(1, std::min((P + kObjThreads - 1) / kObjThreads, std::max(1, sms * 8 / n_smiles)));
    dim3 grid(gx, n_smiles);
    sabr_objective_kernel<<<grid, kObjThreads, 0, (cudaStream_t)stream>>>(beta, d_forward, d_maturity, d_off, d_strikes,
                                                                         d_market, d_weights, d_params, ld, P, d_loss);
    hb_internal_count_launch();
    SB_CUDA(cudaGetLastError());
    return HB_OK;
}

// Host-pointer variants: h_params AoS [P][3] (alpha, rho, nu); copies included.
int hb_sabr_vols_host(int flavour, double beta, double forward, double maturity, int n, const double* h_strikes,
                      const double* h_params, int P, double* h_vols) {
    if (n < 0 || P < 0) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "negative size");
    if (n == 0 || P == 0) return check_beta(beta);
    if (!h_strikes || !h_params || !h_vols) return hb_internal_fail(HB_ERR_INVALID_ARGUMENT, "NULL host pointer");
    std::vector<double> soa((size_t)3 * P);
    for (int p = 0; p < P; ++p)
        for (int c = 0; c < 3; ++c) soa[(size_t)c * P + p] = h_params[(size_t)p * 3 + c];
    double *dK = nullptr, *dX = nullptr, *dV = nullptr;
    SB_CUDA(cudaMalloc(&dK, (size_t)n * 8));
    SB_CUDA(cudaMalloc(&dX, soa.size() * 8));
    SB_CUDA(cudaMalloc(&dV, (size_t)P * n * 8));
    cudaMemcpy(dK, h_strikes, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, soa.data(), soa.size() * 8, cudaMemcpyHostToDevice);
    int rc = hb_sabr_vols(flavour, beta, forward, maturity, n, dK, dX, P, P, dV, nullptr);
    cudaError_t e = cudaMemcpy(h_vols, dV, (size_t)P * n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dK);
    cudaFree(dX);
    cudaFree(dV);
    if (rc) return rc;
    if (e != cudaSuccess) return hb_internal_fail(HB_ERR_CUDA, cudaGetErrorString(e));
    return HB_OK;
}

}  // extern "C"
