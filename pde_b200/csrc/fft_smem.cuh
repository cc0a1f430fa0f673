// fft_smem.cuh -- in-shared-memory Stockham autosort FFT (complex128, forward sign
// e^{-2 pi i jm/N}) for one or several length-N slices that stay on chip.
//
// Layout.  A slice is N complex128 values (16 B each) in shared memory; element i lives
// at 16-byte slot swz(i) = i ^ ((i >> 3) & 7).  With LDS.128/STS.128 a warp access is
// served in four 8-lane phases of 128 B; the XOR swizzle makes every phase of every
// pass (stride-N/8 reads, stride-Ns writes, including the stride-8 writes of the first
// pass) hit eight distinct 16-byte bank groups, i.e. conflict-free, at the price of two
// integer ops per access instead of 12.5 % padding (3 x 64 KiB slices + tables must fit
// the 227 KiB of one sm_100a CTA).
//
// Algorithm.  Radix-8 decimation-in-time passes, Ns = 1, 8, 64, ...: butterfly b reads
// x[b + r N/8], multiplies by W_N^{r k}, k = (b mod Ns) N/(8 Ns), does an in-register
// 8-point DFT and writes y[(b/Ns) 8 Ns + (b mod Ns) + r Ns].  The pass is in place
// (load -> __syncthreads -> store), so one thread owns exactly one butterfly per slice
// (N/8 <= block size).  When several slices are transformed together the barrier of
// slice s doubles as the store/load fence of slices s-1/s+1, so a pass costs one barrier
// per slice instead of two.  W_N^k for k < N/8 comes from a shared table (N/8 entries,
// 8 KiB at N = 4096); the powers W^2..W^7 are formed by multiplication (<= 3 products
// deep, ~1e-16 relative).
//
// The last pass can be PRUNED: Carr-Madan needs Re X_m only at the bins bracketing the
// quoted strikes, so only the listed butterflies (need_q) are evaluated there -- and
// because the last pass of an autosort FFT reads and writes the same slots
// (Ns = N/8 => out index = in index) it needs no barrier between load and store.
//
// Reference: none -- /root/reference has no FFT.  Spec: SURVEY.md
// Appendix B step 5; checked against numpy.fft in tests/test_gpu_parity.py.
#pragma once
#include "heston_math.cuh"

namespace hb {

__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }

__device__ __forceinline__ cplx lds_c(const double2* s, int i) {
    double2 v = s[swz(i)];
    return {v.x, v.y};
}
__device__ __forceinline__ void sts_c(double2* s, int i, cplx v) { s[swz(i)] = make_double2(v.re, v.im); }

#define HB_BF(a, b)                           \
    {                                         \
        cplx _t = {a.re - b.re, a.im - b.im}; \
        a.re += b.re;                         \
        a.im += b.im;                         \
        b = _t;                               \
    }

// 8-point forward DFT in registers.  On return v[] holds X in bit-reversed order:
// X0=v0 X1=v4 X2=v2 X3=v6 X4=v1 X5=v5 X6=v3 X7=v7.
__device__ __forceinline__ void dft8(cplx (&v)[8]) {
    const double s = 0.70710678118654752440;
    HB_BF(v[0], v[4]);
    HB_BF(v[1], v[5]);
    HB_BF(v[2], v[6]);
    HB_BF(v[3], v[7]);
    {  // v5 *= (1-i)/sqrt2 ; v6 *= -i ; v7 *= (-1-i)/sqrt2
        cplx t = v[5];
        v[5] = {(t.re + t.im) * s, (t.im - t.re) * s};
        t = v[6];
        v[6] = {t.im, -t.re};
        t = v[7];
        v[7] = {(t.im - t.re) * s, -(t.re + t.im) * s};
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        HB_BF(v[h + 0], v[h + 2]);
        HB_BF(v[h + 1], v[h + 3]);
        cplx t = v[h + 3];
        v[h + 3] = {t.im, -t.re};
        HB_BF(v[h + 0], v[h + 1]);
        HB_BF(v[h + 2], v[h + 3]);
    }
}
__host__ __device__ constexpr int bitrev3(int r) { return ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1); }

// W^1..W^7 from W (the 7 twiddles of one butterfly).
__device__ __forceinline__ void twiddle_powers(cplx w, cplx (&p)[8]) {
    p[1] = w;
    p[2] = cmul(w, w);
    p[3] = cmul(p[2], w);
    p[4] = cmul(p[2], p[2]);
    p[5] = cmul(p[4], w);
    p[6] = cmul(p[3], p[3]);
    p[7] = cmul(p[4], p[3]);
}

// Fill the W_N^k table, k < N/8 (sincospi keeps the argument reduction exact).
template <int N>
__device__ __forceinline__ void fill_twiddles(double2* tw, int tid, int nthreads) {
    for (int k = tid; k < N / 8; k += nthreads) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)N, &sn, &cs);
        tw[k] = make_double2(cs, sn);
    }
}

// One full radix-8 pass, in place, over slices base + g*N for g < gmax; only slices
// g < count hold data (count is block-uniform), but every thread executes all gmax
// barriers.  A thread owns PER = ceil((N/8)/NT) butterflies and keeps all of them in
// registers across the barrier (the pass is in place).  LINEAR_IN: the pass reads an
// unswizzled slice (as landed by a bulk copy) and writes the swizzled layout.
#ifndef HB_FFT_LEAN_TWIDDLES
#define HB_FFT_LEAN_TWIDDLES 1
#endif
template <int N, int NT, bool LINEAR_IN = false>
__device__ __forceinline__ void fft_pass(double2* base, int count, int gmax, const double2* tw, int Ns, int tid) {
    constexpr int NB = N / 8;
    constexpr int PER = (NB + NT - 1) / NT;
    static_assert(NB % NT == 0 || NB < NT, "butterflies must tile over the block");
#pragma unroll 1
    for (int g = 0; g < gmax; ++g) {
        double2* s = base + (size_t)g * N;
        cplx v[PER][8];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int b = tid + u * NT;
            if (b < NB && g < count) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (LINEAR_IN) {
                        double2 t = s[b + r * NB];
                        v[u][r] = {t.x, t.y};
                    } else {
                        v[u][r] = lds_c(s, b + r * NB);
                    }
                }
            }
        }
        __syncthreads();  // loads of slice g done; stores of slice g-1 visible
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int b = tid + u * NT;
            if (b < NB && g < count) {
                const int bm = b & (Ns - 1);
                if (Ns > 1) {
#if HB_FFT_LEAN_TWIDDLES
                    // running power W^r: 4 live doubles instead of 14 (register budget of the fused kernel)
                    const double2 w = tw[bm * (NB / Ns)];
                    const cplx w1 = {w.x, w.y};
                    cplx wp = w1;
                    v[u][1] = cmul(v[u][1], wp);
#pragma unroll
                    for (int r = 2; r < 8; ++r) {
                        wp = cmul(wp, w1);
                        v[u][r] = cmul(v[u][r], wp);
                    }
#else
                    double2 w = tw[bm * (NB / Ns)];
                    cplx p[8];
                    twiddle_powers({w.x, w.y}, p);
#pragma unroll
                    for (int r = 1; r < 8; ++r) v[u][r] = cmul(v[u][r], p[r]);
#endif
                }
                dft8(v[u]);
                const int j0 = (b - bm) * 8 + bm;
#pragma unroll
                for (int r = 0; r < 8; ++r) sts_c(s, j0 + r * Ns, v[u][bitrev3(r)]);
            }
        }
    }
    if (gmax == 1) __syncthreads();  // a lone slice has no neighbour barrier to fence its stores
}

// Last pass (Ns = N/8) only for the butterflies q in need_q[0..n_need): outputs land in
// place at slots q + r N/8.  No barrier inside; the caller syncs before and after.
template <int N, int NT>
__device__ __forceinline__ void fft_last_pass_pruned(double2* s, const double2* tw, const int* need_q, int n_need,
                                                     int tid) {
    constexpr int NB = N / 8;
    for (int t = tid; t < n_need; t += NT) {
        const int q = need_q[t];
        double2 w = tw[q];
        cplx p[8], v[8];
        twiddle_powers({w.x, w.y}, p);
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = lds_c(s, q + r * NB);
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], p[r]);
        dft8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts_c(s, q + r * NB, v[bitrev3(r)]);
    }
}

}  // namespace hb
