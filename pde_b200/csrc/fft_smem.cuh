// fft_smem.cuh -- in-shared-memory radix-8 FFTs (complex128, forward sign e^{-2 pi i jm/N}) for
// one or several length-N slices that stay on chip.  Two factorisations live here:
//   * fft_pass:  Stockham autosort decimation-in-time passes producing the whole spectrum in natural
//                order -- used by the unfused fft_batch_kernel (HBM -> smem -> HBM);
//   * dif_pass / dif_bin:  in-place decimation-in-frequency passes pruned to the bins Carr-Madan
//                quotes -- used by the fused fft_job_kernel (second half of this file).
//
// Layout.  A slice is N complex128 values (16 B each) in shared memory; element i lives
// at 16-byte slot swz(i) = i ^ ((i >> 3) & 7).  With LDS.128/STS.128 a warp access is
// served in four 8-lane phases of 128 B; the XOR swizzle makes every phase of every
// pass of both factorisations (strides N/8, N/64, 8, 1) hit eight distinct 16-byte bank
// groups, i.e. conflict-free, at the price of two integer ops per access instead of
// 12.5 % padding (3 x 64 KiB slices + tables must fit the 227 KiB of one sm_100a CTA).
//
// Autosort DIT (fft_pass).  Passes Ns = 1, 8, 64, ...: butterfly b reads x[b + r N/8],
// multiplies by W_N^{r k}, k = (b mod Ns) N/(8 Ns), does an in-register 8-point DFT and
// writes y[(b/Ns) 8 Ns + (b mod Ns) + r Ns].  The pass is in place (load -> __syncthreads
// -> store), so one thread owns exactly one butterfly per slice (N/8 <= block size).
// When several slices are transformed together the barrier of slice s doubles as the
// store/load fence of slices s-1/s+1.  W_N^k for k < N/8 comes from a shared table; the
// powers W^2..W^7 are formed by multiplication (~1e-16 relative).
//
// Reference: none -- /root/reference has no FFT.  Spec: SURVEY.md
// Appendix B step 5; checked against numpy.fft in tests/test_gpu_parity.py (fft_pass) and
// through price parity with the oracle's plain radix-2 FFT (dif_pass).
#pragma once
#include "heston_math.cuh"

namespace hb {

// Index checks of the diagnostic build (-DHB_CHECK; compute-sanitizer is closed on the GPU pool): every slice access
// of the fused kernel's transform / epilogue and every table index is asserted; the whole GPU test-suite runs over
// that build once per round (PDE_B200_LIB=...libheston_b200_check.so).  Compiled out of the product library.
#ifdef HB_CHECK
#include <assert.h>
#define HB_ASSERT(cond) assert(cond)
#else
#define HB_ASSERT(cond)
#endif

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
__device__ __forceinline__ void dft8_tail(cplx (&v)[8]);
__device__ __forceinline__ void dft8(cplx (&v)[8]) {
    HB_BF(v[0], v[4]);
    HB_BF(v[1], v[5]);
    HB_BF(v[2], v[6]);
    HB_BF(v[3], v[7]);
    dft8_tail(v);
}
// The same transform when v[4..7] are known zeros (decayed tail of the integrand): the first butterfly
// stage degenerates to copies.  Same operations on the same values as dft8 -> identical bits.
__device__ __forceinline__ void dft8_lo4(cplx (&v)[8]) {
    v[4] = v[0];
    v[5] = v[1];
    v[6] = v[2];
    v[7] = v[3];
    dft8_tail(v);
}
__device__ __forceinline__ void dft8_tail(cplx (&v)[8]) {
    const double s = 0.70710678118654752440;
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

// ================================================================================================
// In-place decimation-in-frequency transform of the fused kernel (N = 8^L).
//
// Carr-Madan reads Re X_m only at the few bins bracketing the quoted strikes (a window of ~70
// consecutive m for strikes 80..120 at N = 4096).  A decimation-in-FREQUENCY factorisation prunes
// that much better than the autosort DIT passes above: pass i (stride S = N/8^i) resolves digit
// m_{i-1} of the output index m = m_0 + 8 m_1 + 64 m_2 + ..., so
//   * the passes are in place on the thread's own 8 slots -- no load/barrier/store split, one
//     barrier per pass for all slices of a group, and the FIRST pass needs no barrier at all
//     because butterfly `tid` reads exactly the points tid + r N/8 that thread `tid` wrote in K1;
//   * the second-to-last pass (S = 8) only forms the outputs whose digit m_{L-2} occurs in the
//     window (fmask, typically 2 of 8) -- one twiddle product per kept output instead of 7;
//   * the last pass (S = 1) is not a pass: the interpolation epilogue evaluates the single output
//     X_m = sum_r y[pos(m) + r] w8^{r m_{L-1}} it needs (dif_bin).
// X_m ends up at the digit-reversed position pos(m) = m_0 8^{L-1} + m_1 8^{L-2} + ... + m_{L-1}.
// Twiddles: the first pass of a grid above 512 points multiplies output f by W_N^{t f}, formed as
// running powers of W_N^t (table tw, N/8 entries); every later pass has 8 S <= 512 and reads
// W_{8S}^{t f} = W_512^{t f 512/(8S)} directly from a 512-entry table (tw512), which moves work
// from the FP64 pipe (the binding one) to the idle LSU.
// ================================================================================================

constexpr int kMaxGroupFft = 3;  // slices resident at once (kernels.cuh: kMaxGroup)

template <int N>
struct Log8 {
    static constexpr int value = 1 + Log8<N / 8>::value;
};
template <>
struct Log8<1> {
    static constexpr int value = 0;
};

// Twiddle tables behind the W_N one: W_512^k (512 entries), W_64^k (64), w8^k (8), each compact so that
// consecutive lanes read consecutive 16-byte slots (a strided walk through the 512-entry table would put
// all lanes of the S = 8 pass on one bank group).
constexpr int kTwSmall = 512 + 64 + 8;
__device__ __forceinline__ void fill_tw512(double2* t512, int tid, int nthreads) {
    for (int k = tid; k < kTwSmall; k += nthreads) {
        const int n = (k < 512) ? 512 : (k < 576 ? 64 : 8);
        const int kk = (k < 512) ? k : (k < 576 ? k - 512 : k - 576);
        double sn, cs;
        sincospi(-2.0 * (double)kk / (double)n, &sn, &cs);
        t512[k] = make_double2(cs, sn);
    }
}

// One DIF pass with stride S over slices g < count.  MASKED: only outputs f with bit f of fmask[g] set
// are formed (the others are never read again).
template <int N, int NT, int S, bool MASKED>
__device__ __forceinline__ void dif_pass(double2* base, int count, const double2* tw, const double2* tw512,
                                         const unsigned* fmask, int tid) {
    constexpr int NB = N / 8;
    constexpr int PER = (NB + NT - 1) / NT;
    static_assert(NB % NT == 0 || NB < NT, "butterflies must tile over the block");
    static_assert(8 * S == N || 8 * S == 512 || 8 * S == 64, "twiddle table for this stride");
    constexpr bool RUNNING = (8 * S == N) && (N > 512);
    const double2* tws = (8 * S == 64) ? tw512 + 512 : tw512;  // W_{8S}^k
    constexpr int TMASK = (8 * S == 64) ? 63 : 511;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int b = tid + u * NT;
        if (b < NB) {
            const int t = b & (S - 1);
            const int p0 = ((b - t) << 3) + t;
            // twiddles of this butterfly, shared by all slices of the group (a full pass needs all seven;
            // the masked pass looks its one or two up per slice)
            cplx wf[8];
            if (!MASKED) {
                if (RUNNING) {
                    const double2 w = tw[t];
                    wf[1] = {w.x, w.y};
#pragma unroll
                    for (int f = 2; f < 8; ++f) wf[f] = cmul(wf[f - 1], wf[1]);
                } else {
#pragma unroll
                    for (int f = 1; f < 8; ++f) {
                        const double2 w = tws[(t * f) & TMASK];
                        wf[f] = {w.x, w.y};
                    }
                }
            }
            // (Issuing the loads of slice g + 1 before slice g is transformed was tried: unrolled or with
            // a register rotation it costs 0.9-2 KB of spills at 128 registers -- profiles/r01_shape_sweep.txt.)
#pragma unroll 1
            for (int g = 0; g < count; ++g) {
                double2* sl = base + (size_t)g * N;
                const unsigned keep = MASKED ? fmask[g] : 0xffu;
                cplx v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) v[r] = lds_c(sl, p0 + r * S);
                dft8(v);  // X_f in v[bitrev3(f)]
#pragma unroll
                for (int f = 1; f < 8; ++f) {
                    if (!MASKED) {
                        v[bitrev3(f)] = cmul(v[bitrev3(f)], wf[f]);
                    } else if ((keep >> f) & 1u) {
                        const double2 w = tws[(t * f) & TMASK];
                        v[bitrev3(f)] = cmul(v[bitrev3(f)], {w.x, w.y});
                    }
                }
#pragma unroll
                for (int f = 0; f < 8; ++f)
                    if (!MASKED || ((keep >> f) & 1u)) sts_c(sl, p0 + f * S, v[bitrev3(f)]);
            }
        }
    }
}

// ---- zero-aware first two passes (N / 8 == NT: butterfly `tid` of the first pass owns exactly the points
// tid + r N/8 that thread `tid` produced in K1) -------------------------------------------------------------
// The integrand decays: beyond some index every point of a slice is an exact zero (exp underflow, or the plan's
// significance cut).  K1 hands over what it knows -- bit 8 g + k of `live`: this thread's k-th point of slice g is
// non-zero -- and the first pass skips what is trivially zero:
//   no live point      nothing to do (the slots already hold K1's zeros)
//   only row 0 live    X_f = v0 for every f: seven twiddle products, no butterfly
//   rows 0..3 live     dft8_lo4
// Results are bit-identical to transforming the stored zeros.
template <int N, int NT>
__device__ __forceinline__ void dif_pass_first(double2* base, int count, const double2* tw, unsigned live, int tid) {
    static_assert(N / 8 == NT && N > 512, "first pass on the thread's own points, running twiddle powers");
    constexpr int S = N / 8;
    if (live == 0u) return;
    cplx wf[8];
    {
        const double2 w = tw[tid];
        wf[1] = {w.x, w.y};
#pragma unroll
        for (int f = 2; f < 8; ++f) wf[f] = cmul(wf[f - 1], wf[1]);
    }
#pragma unroll 1
    for (int g = 0; g < count; ++g) {
        const unsigned lm = (live >> (8 * g)) & 0xffu;
        if (lm == 0u) continue;
        double2* sl = base + (size_t)g * N;
        HB_ASSERT(tid >= 0 && tid + 7 * S < N);
        if (lm == 1u) {
            const cplx v0 = lds_c(sl, tid);
#pragma unroll
            for (int f = 1; f < 8; ++f) sts_c(sl, tid + f * S, cmul(v0, wf[f]));
            continue;
        }
        cplx v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = ((lm >> r) != 0u) ? lds_c(sl, tid + r * S) : cplx{0.0, 0.0};
        if (lm < 16u) dft8_lo4(v);
        else dft8(v);
#pragma unroll
        for (int f = 1; f < 8; ++f) v[bitrev3(f)] = cmul(v[bitrev3(f)], wf[f]);
#pragma unroll
        for (int f = 0; f < 8; ++f) sts_c(sl, tid + f * S, v[bitrev3(f)]);
    }
}

// Second pass (S = N/64).  jlive[g] = 1 + the highest non-zero point index of slice g (0: none).  If it is at
// most N/8 only row 0 of the first pass was live, every sub-slice f holds y_f[t] = x_t W^{t f} with support
// t < jlive, and the butterfly of thread (c, t) reads the elements t + r S of sub-slice c: rows with
// t + r S >= jlive are zeros and are neither loaded nor added.
template <int N, int NT>
__device__ __forceinline__ void dif_pass_second(double2* base, int count, const double2* tw512, const int* jlive,
                                                int tid) {
    static_assert(N / 8 == NT && N == 4096, "stride N/64 = 64, W_512 twiddles");
    constexpr int S = N / 64;
    const int t = tid & (S - 1);
    const int p0 = ((tid - t) << 3) + t;
    int nr[kMaxGroupFft];
    int any = 0;
#pragma unroll
    for (int g = 0; g < kMaxGroupFft; ++g) {
        const int J = (g < count) ? jlive[g] : 0;
        HB_ASSERT(J >= 0 && J <= N && p0 >= 0 && p0 + 7 * S < N);
        nr[g] = (J > N / 8) ? 8 : max(0, min(8, (J - t + S - 1) / S));
        any |= nr[g];
    }
    if (any == 0) return;
    cplx wf[8];
#pragma unroll
    for (int f = 1; f < 8; ++f) {
        const double2 w = tw512[(t * f) & 511];
        wf[f] = {w.x, w.y};
    }
#pragma unroll
    for (int g = 0; g < kMaxGroupFft; ++g) {
        if (g >= count || nr[g] == 0) continue;
        double2* sl = base + (size_t)g * N;
        cplx v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = (r < nr[g]) ? lds_c(sl, p0 + r * S) : cplx{0.0, 0.0};
        if (nr[g] <= 4) dft8_lo4(v);
        else dft8(v);
#pragma unroll
        for (int f = 1; f < 8; ++f) v[bitrev3(f)] = cmul(v[bitrev3(f)], wf[f]);
#pragma unroll
        for (int f = 0; f < 8; ++f) sts_c(sl, p0 + f * S, v[bitrev3(f)]);
    }
}

// Digit of m resolved by the masked (S = 8) pass, and the position of X_m's last-pass inputs.
template <int N>
__host__ __device__ constexpr int dif_mask_digit(int m) {
    return (m >> (3 * (Log8<N>::value - 2))) & 7;
}
template <int N>
__device__ __forceinline__ int dif_pos(int m) {
    constexpr int L = Log8<N>::value;
    int p = 0;
#pragma unroll
    for (int i = 0; i < L - 1; ++i) p |= ((m >> (3 * i)) & 7) << (3 * (L - 1 - i));
    return p;
}

// X_m of a slice that went through every pass down to S = 8: the one needed output of its last
// radix-8 butterfly (w8^k from the compact table behind W_512 and W_64).  WANT_IM = false returns Re X_m only (im = 0).
// w8^k = e^{-2 pi i k/8} from the constant bank (f is the same for every lane of a warp in practice: the quoted
// bins share their top digit, so the dynamic index is uniform)
__constant__ double kW8c[8] = {1.0, 0.70710678118654752440, 0.0, -0.70710678118654752440,
                               -1.0, -0.70710678118654752440, 0.0, 0.70710678118654752440};
__constant__ double kW8s[8] = {0.0, -0.70710678118654752440, -1.0, -0.70710678118654752440,
                               0.0, 0.70710678118654752440, 1.0, 0.70710678118654752440};

template <int N, bool WANT_IM>
__device__ __forceinline__ cplx dif_bin(const double2* sl, const double2* tw512, int m) {
    const int p = dif_pos<N>(m);
    const int f = (m >> (3 * (Log8<N>::value - 1))) & 7;
    cplx acc = {0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const cplx v = lds_c(sl, p + r);
        const double2 w = make_double2(kW8c[(r * f) & 7], kW8s[(r * f) & 7]);
        acc.re = fma(v.re, w.x, fma(-v.im, w.y, acc.re));
        if (WANT_IM) acc.im = fma(v.re, w.y, fma(v.im, w.x, acc.im));
    }
    return acc;
}

}  // namespace hb
