// direct_kernel.cuh -- the Carr-Madan job kernel for slices whose integrand has decayed: live prefix + direct sums.
//
// With the plan's significance cut (kernels.cuh, GridConst::cut) a slice of the calibrator's box keeps a short
// PREFIX of the damped grid (median 216 of 4096 points, mean 378) and Carr-Madan reads ~70 consecutive bins.  At
// that size the transform is not an FFT problem: inputs x outputs / N ~ 3.5, so a pruned radix-8 network does as
// many operations as the plain sums X_m = sum_{j<J} x_j W^{jm}, and spends its time on barriers and on 4096-slot
// passes over zeros.  This kernel therefore
//   * bounds the live prefix J of every (class, maturity) rigorously from the class constants (prefix_bound.cuh) and
//     touches nothing beyond it: stage A, stage B, stage F and the sums all run over j < J only (points inside the
//     prefix are still tested one by one against the cut, exactly as the transform kernel tests all N);
//   * flattens the (maturity, point) pairs of a parameter set into WAVES of <= CAP points, so that every thread of the
//     CTA has characteristic-function work whatever the prefix length (one task = one class at one point of one
//     maturity: stage B once, then the final cexp of each variant of the class), and all six variants of a maturity
//     are resident together;
//   * forms the bins by direct summation with the conjugate-pair identity: around the centre m_c of a maturity's
//     bins, with z_j = x_j W^{j m_c} (the rotation is folded into the Carr-Madan weight),
//         X_{m_c +- mu} = P +- Q,   P = sum_j Re z_j cos(2 pi j mu/N),   Q = sum_j Im z_j sin(2 pi j mu/N),
//     i.e. ONE fused multiply-add per (point, bin, slice); one thread owns a (chunk of <= 64 points, pair mu) and
//     all six slices, so its twiddle (an exact table value advanced by a two-way interleaved recurrence over at most
//     32 steps) is shared by twelve accumulators;
//   * adds the chunk partials of a maturity in chunk order (fixed: the bits do not depend on how the waves were
//     packed, on the launch path, the CTA shape or the rank), then interpolates / clamps / applies parity as K3 does.
// Two CTA barriers per wave (~400 points x 6 slices; the finish step of a wave overlaps the next wave's K1) instead of
// three per group of three slices, and two or more CTAs per SM so that one CTA's barrier is another's compute.
// Everything after the prices (residuals, SciPy forward differences, normal equations) is kernels.cuh's finalize.
// Host dispatch (heston_b200.cu): FFT-mode plans with a significance cut, <= kDMaxPairs (six variants) / kDMaxPairs1
// (one variant) pairs per maturity; prefix_scan_kernel below bounds the live prefix of every (set, maturity) first and
// sends the sets whose prefix is long (slow decay) to fft_job_kernel, which costs O(N log N) whatever the prefix.
#pragma once
#include "kernels.cuh"
#include "prefix_bound.cuh"

namespace hb {

// CTA shape, measured on B200 (profiles/r02_direct_shapes.txt): two CTAs per SM beat one of 512 / 640 / 768 threads --
// the phases of a wave (stage B/F tasks, direct sums, finish) are separated by CTA barriers, and a second CTA fills the
// FP64 pipe while the first one waits (256 x 2 against 512 x 1: +4.5 % normal equations, +32 % objective); 384 threads
// per CTA (80 registers, ~300 B of spills, 24 warps per SM) add another 3 % over 256 (126 registers, 16 warps).
#ifndef HB_DNT
#define HB_DNT 384
#endif
#ifndef HB_DCAP
#define HB_DCAP 448
#endif
#ifndef HB_DITEMS
#define HB_DITEMS 384
#endif
#ifndef HB_DCTAS
#define HB_DCTAS 2
#endif
#ifndef HB_DNT1
#define HB_DNT1 256
#endif
#ifndef HB_DCAP1
#define HB_DCAP1 192
#endif
#ifndef HB_DITEMS1
#define HB_DITEMS1 128
#endif
#ifndef HB_DCTAS1
#define HB_DCTAS1 5
#endif
#ifndef HB_DPULL
#define HB_DPULL 64  // K1 tasks per trip to the queue (32 or 64)
#endif
#ifndef HB_DPREFETCH
#define HB_DPREFETCH 0
#endif
#ifndef HB_DCHUNK
#define HB_DCHUNK 64
#endif
constexpr int kDChunk = HB_DCHUNK;  // points per chunk (one DFT item = chunk x pair)
constexpr int kDMaxPairs = 64;    // conjugate pairs per maturity the six-variant kernel takes
constexpr int kDMaxPairs1 = 256;  // ... the one-variant kernel (1/6 of the shared memory per pair: config 5's 200 strikes fit)
constexpr int kDMaxMat = 512;     // maturities per surface (prefix table in shared memory)
constexpr int kDMaxCh = 64;       // chunks per wave
constexpr int kDMaxSeg = 48;      // maturity pieces per wave
constexpr int kDAFields = 5;      // stage-A cache: num, L0, d, g, q1
constexpr int kDMapMax = 4096 / 32 + 1;  // coarse slot / item maps of a wave (CAP, ITEMS <= 4096)

struct DirectDev {
    const double2* tw;    // [2 n_full]  (cos, sin)(pi k / n_full)
    const double2* tab;   // [n_full]    Simpson weight x e^{i b v_j} / (alpha^2 + alpha - v^2 + i (2 alpha + 1) v)
    const int* blk;       // [nblk + 1]  block boundaries of the prefix bound
    int nblk;
    int n_full;           // grid length N
    const int* mat_c2;    // [n_mat]     doubled centre of the maturity's bins (m_lo + m_hi)
    const int* mat_rot;   // [n_mat]     row of `tabrot` holding tab_j W^{j m_c} for this maturity's centre, -1 = none
    const double2* tabrot;  // [rows][n_full]
    const int* pair_off;  // [n_mat + 1]
    const int* pair_d2;   // doubled offsets |2m - c2| of the maturity's bins, distinct, ascending
    const int* opt_pq0;   // [n_sorted]  lower bracketing bin: (pair index << 1) | (bin above the centre), -1 = off grid
    const int* opt_pq1;   // upper bracketing bin
    double2* acache;      // [grid][classes][kDAFields][n_full]  stage-A cache of the CTA's current job
    const int* job_ids;   // optional: the parameter sets this launch prices (null = 0..P-1)
    const int* jtab;      // optional [P][n_mat]: prefix lengths precomputed by prefix_scan_kernel (null = compute here)
    const int* p_count;   // optional: number of entries of job_ids, written on the device by prefix_scan_kernel
    int max_pairs;        // most conjugate pairs any maturity of the surface has
};

template <bool ONEVAR>
struct DirectCfg {
    static constexpr int V = ONEVAR ? 1 : 6;
    static constexpr int NCLS = ONEVAR ? 1 : 4;
    // the one-variant kernel (prices, objective) holds 1/6 of the shared memory per point: five CTAs of 256 threads per SM
    // with waves of 768 points measured +30 % over the six-variant kernel's shape (objective 49.5 -> 64.5 M slices/s; 128
    // threads x 6 CTAs is as fast but its finalize would sum over 128 threads: the loss column of the normal equations
    // must equal the objective's bit for bit)
    static constexpr int NT = ONEVAR ? HB_DNT1 : HB_DNT;       // threads per CTA
    static constexpr int CTAS = ONEVAR ? HB_DCTAS1 : HB_DCTAS;  // CTAs per SM
    static constexpr int MAXP = ONEVAR ? kDMaxPairs1 : kDMaxPairs;      // conjugate pairs per maturity
    static constexpr int CAP = ONEVAR ? 4 * HB_DCAP1 : HB_DCAP;        // wave capacity in points
    static constexpr int ITEMS = ONEVAR ? 4 * HB_DITEMS1 : HB_DITEMS;  // (chunk, pair) items per wave
    // nblk: blocks of the prefix bound (the block table of an unrouted launch aliases the wave buffers)
    static constexpr size_t smem_bytes(int nblk = 0) {
        const size_t buffers = (size_t)V * CAP * 16 + (size_t)V * ITEMS * 16 + (size_t)2 * V * MAXP * 16;
        const size_t blocks = (size_t)NCLS * (nblk + (nblk + 7) / 8) * sizeof(PrefixBlock);
        return buffers > blocks ? buffers : blocks;
    }
};

struct DirectWave {
    int nch, nslots, nitems, nseg, more;
    int next_m, next_c;  // cursor behind this wave: maturity and chunk within it
    int c_mat[kDMaxCh], c_j0[kDMaxCh], c_len[kDMaxCh], c_seg[kDMaxCh];
    int c_slot0[kDMaxCh + 1], c_item0[kDMaxCh + 1];
    int c_rot[kDMaxCh];                     // tabrot row offset (row * n_full) or -1
    double c_T[kDMaxCh], c_lsm[kDMaxCh];    // maturity, ln S0 + (r - q) T
    int s_mat[kDMaxSeg], s_c0[kDMaxSeg], s_c1[kDMaxSeg], s_flags[kDMaxSeg];  // flags: 1 = from j = 0, 2 = ends the maturity
    int s_t0[kDMaxSeg + 1];  // finish tasks per slice before piece sg: options of an ending piece, pairs of a cut one
    // coarse maps: the chunk holding slot 32 g / item 32 g (the exact chunk is at most a few steps further on)
    unsigned char smap[kDMapMax], imap[kDMapMax];
};

// chunking of a prefix of J points: nc chunks of len points (the last one shorter); depends on J alone
__device__ __forceinline__ void direct_chunking(int J, int& nc, int& len) {
    nc = (J + kDChunk - 1) / kDChunk;
    len = (nc > 0) ? (J + nc - 1) / nc : 0;
}

// variants of class index ci (0: {base, theta', v0'}, 1: kappa', 2: sigma', 3: rho') and its parameter-class id
__device__ __forceinline__ int direct_cls_variant(int ci) { return ci == 0 ? 0 : (ci == 1 ? 1 : (ci == 2 ? 3 : 4)); }

template <int CAP, int ITEMS>
__device__ void direct_build_wave(DirectWave& w, const int* s_J, const SurfaceDev& S, const DirectDev& D, int m, int c,
                                  int m_end) {
    int nch = 0, nslots = 0, nitems = 0, nseg = 0;
    while (m < m_end) {
        const int np = D.pair_off[m + 1] - D.pair_off[m];
        if (np == 0) {  // nothing on the grid: a piece without chunks, so that its options still get their NaN
            if (nseg >= kDMaxSeg) break;
            w.s_mat[nseg] = m;
            w.s_c0[nseg] = w.s_c1[nseg] = nch;
            w.s_flags[nseg] = 3;
            ++nseg;
            ++m;
            c = 0;
            continue;
        }
        int nc, len;
        const int J = s_J[m];
        direct_chunking(J, nc, len);
        int added = 0;
        while (c < nc) {
            const int l = min(len, J - c * len);
            if (nch >= kDMaxCh || nslots + l > CAP || nitems + np > ITEMS) break;
            if (added == 0) {
                if (nseg >= kDMaxSeg) break;
                w.s_mat[nseg] = m;
                w.s_c0[nseg] = nch;
                w.s_flags[nseg] = (c == 0) ? 1 : 0;
                ++nseg;
            }
            w.c_mat[nch] = m;
            if (added == 0) {
                const double T = S.mat_T[m];
                w.c_T[nch] = T;
                w.c_lsm[nch] = S.ln_spot + (S.rate - S.dividend) * T;
                w.c_rot[nch] = D.mat_rot[m] >= 0 ? D.mat_rot[m] * D.n_full : -1;
            } else {
                w.c_T[nch] = w.c_T[nch - 1];
                w.c_lsm[nch] = w.c_lsm[nch - 1];
                w.c_rot[nch] = w.c_rot[nch - 1];
            }
            w.c_j0[nch] = c * len;
            w.c_len[nch] = l;
            w.c_seg[nch] = nseg - 1;
            w.c_slot0[nch] = nslots;
            w.c_item0[nch] = nitems;
            nslots += l;
            nitems += np;
            ++nch;
            ++c;
            ++added;
        }
        if (added) w.s_c1[nseg - 1] = nch;
        if (c == nc) {
            if (added) w.s_flags[nseg - 1] |= 2;
            ++m;
            c = 0;
        } else {
            break;  // wave full
        }
    }
    HB_ASSERT(nch <= kDMaxCh && nseg <= kDMaxSeg && nslots <= CAP && nitems <= ITEMS);
    w.c_slot0[nch] = nslots;
    w.c_item0[nch] = nitems;
    w.s_t0[0] = 0;
    for (int sg = 0; sg < nseg; ++sg) {
        const int mm = w.s_mat[sg];
        w.s_t0[sg + 1] = w.s_t0[sg] + ((w.s_flags[sg] & 2) ? S.mat_off[mm + 1] - S.mat_off[mm]
                                                            : D.pair_off[mm + 1] - D.pair_off[mm]);
    }
    for (int g = 0, c2 = 0; 32 * g < nslots; ++g) {
        while (w.c_slot0[c2 + 1] <= 32 * g) ++c2;
        w.smap[g] = (unsigned char)c2;
    }
    for (int g = 0, c2 = 0; 32 * g < nitems; ++g) {
        while (w.c_item0[c2 + 1] <= 32 * g) ++c2;
        w.imap[g] = (unsigned char)c2;
    }
    w.nch = nch;
    w.nslots = nslots;
    w.nitems = nitems;
    w.nseg = nseg;
    w.next_m = m;
    w.next_c = c;
    w.more = (m < m_end) ? 1 : 0;
}

// index of the chunk whose [start[c], start[c+1]) holds x (start ascending, start[n] = total > x), from the coarse map
__device__ __forceinline__ int direct_find(const int* start, const unsigned char* coarse, int x) {
    int c = coarse[x >> 5];
    while (start[c + 1] <= x) ++c;
    return c;
}

// Prefix length of one (class, maturity): the start of the first block such that it and every later block is dead for
// all nv variants of the class.  Two levels: super-blocks of kDSuper fine blocks (one bound over the merged range,
// looser but valid) are scanned from the end of the grid first, then the fine blocks below the first live one.
constexpr int kDSuper = 8;
__device__ __forceinline__ int direct_nsuper(int nblk) { return (nblk + kDSuper - 1) / kDSuper; }
__device__ __forceinline__ int direct_prefix(const PrefixBlock* fine, const PrefixBlock* super, const int* blk, int nblk,
                                             double T, double cst, const double* kts, const double* v0s, int nv,
                                             double cut) {
    // ONE loop over super-blocks (from the end, while dead) and then fine blocks (downwards, while dead), with one call
    // site of the bound: the bound's code is several hundred instructions, and a second inlined copy made the scan
    // kernel miss its instruction cache 40 % of the time (profiles/r02_f_prefix_scan_kernel_bench_launch.txt)
    int sb = direct_nsuper(nblk) - 1, k = nblk - 1;
    bool in_super = true;
    while (k >= 0) {
        const PrefixTerms t = prefix_terms(in_super ? super[sb] : fine[k], T);  // shared by the variants of the class
        bool dead = true;
        for (int i = 0; i < nv; ++i) dead = dead && (prefix_ub_of(t, kts[i], v0s[i], cst) < cut - kPrefixMargin);
        if (in_super) {
            if (dead) {
                k = kDSuper * sb - 1;  // every fine block of this super-block is dead
                --sb;
            } else {
                in_super = false;  // its fine blocks, from the last one downwards
            }
        } else {
            if (!dead) break;
            --k;
        }
    }
    return blk[k + 1];  // all blocks dead: blk[0] = 1 (point 0 is always evaluated)
}
// fine and super blocks of a class: entry i < nblk is fine block i, entry nblk + s super-block s
__device__ __forceinline__ PrefixBlock direct_block_at(const PrefixClass& pc, const int* blk, int nblk, int i, double eta) {
    const int k0 = (i < nblk) ? i : kDSuper * (i - nblk);
    const int k1 = (i < nblk) ? i + 1 : min(nblk, k0 + kDSuper);
    return prefix_block(pc, eta * (double)blk[k0], eta * (double)(blk[k1] - 1));
}

template <bool ONEVAR>
__global__ void __launch_bounds__(DirectCfg<ONEVAR>::NT, DirectCfg<ONEVAR>::CTAS)
direct_job_kernel(SurfaceDev S, DirectDev D, GridConst gc, Bounds bd, const double* __restrict__ params, int ld, int P,
                  int what, double* __restrict__ out, double* __restrict__ out2, double* __restrict__ scratch,
                  int pieces, unsigned long long* job_counter) {
    using Cfg = DirectCfg<ONEVAR>;
    constexpr int V = Cfg::V, NCLS = Cfg::NCLS, CAP = Cfg::CAP, ITEMS = Cfg::ITEMS, NT = Cfg::NT, MAXP = Cfg::MAXP;
    static_assert(ITEMS >= MAXP, "a wave holds at least one chunk of the widest maturity");
    static_assert(kFinalizeT<NT>() == 256, "every Carr-Madan job kernel sums its finalize over 256 threads (same bits)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* ybuf = reinterpret_cast<double2*>(smem_raw);  // [V][CAP]   z_j of the wave's points
    double2* part = ybuf + (size_t)V * CAP;                // [V][ITEMS] (P, Q) of every (chunk, pair)
    // [2][V][MAXP] running sums of a maturity cut by a wave; by wave parity: the first piece of a wave may read
    // the sums its predecessor left while the last piece of the same wave leaves its own
    double2* carry = part + (size_t)V * ITEMS;
    PrefixBlock* s_blocks = reinterpret_cast<PrefixBlock*>(smem_raw);  // [NCLS][nblk], job set-up only (aliases ybuf)
    __shared__ DirectWave waves[3];  // running wave, the one being built, and the previous one (its finish step may still run)
    __shared__ JobState js;
    __shared__ double red[(NT / 32) * 23];
    __shared__ int s_J[kDMaxMat];
    __shared__ ClassConst s_cc[4];
    __shared__ double s_kts[6], s_v0s[6];
    __shared__ int s_jall;
    __shared__ int s_task;  // K1 task queue of the running wave (blocks of 32 tasks, pulled per warp)
    __shared__ long long s_job;
    const int tid = threadIdx.x;
    const int n = S.n_opt, M = S.n_mat, NF = D.n_full;
    const unsigned tw_mask = 2u * (unsigned)NF - 1u;  // n_full is a power of two
    const long long n_jobs = (long long)(D.p_count ? *D.p_count : P) * pieces;
    double2* acache = D.acache + (size_t)blockIdx.x * NCLS * kDAFields * NF;
    auto afield = [&](int ci, int f) -> double2* { return acache + ((size_t)ci * kDAFields + f) * NF; };

    auto next_job = [&](long long cur) -> long long {
        if (!job_counter) return cur + gridDim.x;
        __syncthreads();
        if (tid == 0) s_job = (long long)gridDim.x + (long long)atomicAdd(job_counter, 1ULL);
        __syncthreads();
        return s_job;
    };

    for (long long job = blockIdx.x; job < n_jobs; job = next_job(job)) {
        const int pj = (int)(job / pieces), piece = (int)(job % pieces);
        const int p = D.job_ids ? D.job_ids[pj] : pj;
        const int m_begin = (int)((long long)M * piece / pieces), m_end = (int)((long long)M * (piece + 1) / pieces);
        __syncthreads();  // the previous job's finalize has consumed rows / js / the descriptors
#ifdef HB_DBG_POISON
        {  // diagnostic build: anything read before it is written shows up as NaN
            const double2 nan2 = make_double2(__longlong_as_double(0x7ff8000000000000LL), __longlong_as_double(0x7ff8000000000000LL));
            for (int i = tid; i < V * CAP + V * ITEMS + 2 * V * MAXP; i += NT) ybuf[i] = nan2;
            for (int i = tid; i < NCLS * kDAFields * NF; i += NT) acache[i] = nan2;
        }
        __syncthreads();
#endif
        if (tid == 0) {
            job_setup(js, params, ld, p, bd, V);
            s_jall = 0;
            s_task = 0;
        }
        for (int m = tid; m < M; m += NT) s_J[m] = 0;
        __syncthreads();
        if (!js.valid) {
            if (pieces == 1 || (what == W_PRICE && piece == 0)) invalid_job<NT>(what, S, p, out, out2, tid);
            continue;
        }
        double* rows = (what == W_PRICE) ? out + (size_t)p * n
                                         : scratch + (size_t)(pieces > 1 ? p : (int)blockIdx.x) * 6 * n;
        if (piece == 0) {
            for (int v = 0; v < V; ++v)
                for (int i = tid; i < S.n_intr; i += NT) rows[(size_t)v * n + S.intr_orig[i]] = S.intr_val[i];
        }
        // class constants and per-variant slice constants, as fill_group forms them
        if (tid < NCLS) {
            const double* xc = js.x[direct_cls_variant(tid)];
            s_cc[tid] = {xc[0], xc[2] * xc[2], xc[3] * xc[2]};
        }
        if (tid >= 32 && tid < 32 + V) {
            const int v = tid - 32;
            const int cls = (v == 1 || v == 3 || v == 4) ? v : 0;
            const double* xc = js.x[cls];
            const double* xv = js.x[v];
            const double s2 = xc[2] * xc[2];
            s_kts[v] = xv[0] * xv[1] / s2;  // kappa*theta/sigma^2, heston.cpp:65
            s_v0s[v] = xv[4] / s2;
        }
        __syncthreads();
        // ---- live prefix of every maturity (prefix_bound.cuh) ----
        if (D.jtab) {
            for (int m = m_begin + tid; m < m_end; m += NT) s_J[m] = D.jtab[(size_t)p * M + m];
        } else {
            const int nb2 = D.nblk + direct_nsuper(D.nblk);  // fine + super blocks per class
            for (int i = tid; i < NCLS * nb2; i += NT) {
                const int ci = i / nb2;
                const PrefixClass pc = prefix_class(s_cc[ci], gc.alpha);
                s_blocks[i] = direct_block_at(pc, D.blk, D.nblk, i - ci * nb2, gc.eta);
            }
            __syncthreads();
            const int nm = m_end - m_begin;
            for (int i = tid; i < NCLS * nm; i += NT) {  // one (class, maturity) per thread
                const int ci = i / nm, m = m_begin + (i - ci * nm);
                const double T = S.mat_T[m];
                const double cst = -gc.ui * (S.ln_spot + (S.rate - S.dividend) * T);
                double kts[3], v0s[3];
                int nv = 1;
                if (ci == 0) {
                    kts[0] = s_kts[0];
                    v0s[0] = s_v0s[0];
                    if (V > 1) {
                        kts[1] = s_kts[2];
                        v0s[1] = s_v0s[2];
                        kts[2] = s_kts[5];
                        v0s[2] = s_v0s[5];
                        nv = 3;
                    }
                } else {
                    const int v = direct_cls_variant(ci);
                    kts[0] = s_kts[v];
                    v0s[0] = s_v0s[v];
                }
                const PrefixBlock* fb = s_blocks + (size_t)ci * nb2;
                const int J = direct_prefix(fb, fb + D.nblk, D.blk, D.nblk, T, cst, kts, v0s, nv, gc.cut);
                atomicMax(&s_J[m], min(J, NF));
            }
            __syncthreads();
        }
        {
            int jm = 0;
            for (int m = m_begin + tid; m < m_end; m += NT) jm = max(jm, s_J[m]);
            jm = __reduce_max_sync(0xffffffffu, jm);
            if ((tid & 31) == 0 && jm > 0) atomicMax(&s_jall, jm);
        }
        __syncthreads();  // also: the prefix blocks (aliasing ybuf) have been read for the last time
        // ---- stage A (+ L0 of the asymptotic stage B) once per (class, point of the longest prefix) ----
        {
#ifndef HB_DBG_NOFILL
#define HB_DBG_NOFILL 0
#endif
            const int jall = HB_DBG_NOFILL ? 0 : s_jall;
            for (int i = tid; i < NCLS * jall; i += NT) {
                const int ci = i / jall, j = i - ci * jall;
                const StageA a = stage_a(s_cc[ci], gc.eta * (double)j, gc.ui);
                const cplx l0 = stage_b_l0(a);
                afield(ci, 0)[j] = make_double2(a.num.re, a.num.im);
                afield(ci, 1)[j] = make_double2(l0.re, l0.im);
                afield(ci, 2)[j] = make_double2(a.d.re, a.d.im);
                afield(ci, 3)[j] = make_double2(a.g.re, a.g.im);
                afield(ci, 4)[j] = make_double2(a.q1.re, a.q1.im);
            }
            if (tid == 0) direct_build_wave<CAP, ITEMS>(waves[0], s_J, S, D, m_begin, 0, m_end);
        }
        __syncthreads();
        for (int wi = 0;; ++wi) {
            const DirectWave& w = waves[wi % 3];
            // the next wave's descriptor is built by one thread while K1 runs: no barrier separates the finish step of
            // wave wi - 1 from this point, hence three buffers (buffer (wi + 1) % 3 was last read in wave wi - 2)
            if (tid == NT - 1 && w.more)
                direct_build_wave<CAP, ITEMS>(waves[(wi + 1) % 3], s_J, S, D, w.next_m, w.next_c, m_end);
            const int nslots = w.nslots, nch = w.nch, nitems = w.nitems;
            // ---- K1: one task = (class, point of a maturity): stage B, then stage F of the class's variants ----
            // Tasks differ in cost (class 0 carries three final cexps, low grid points the full stage B): warps pull
            // blocks of 32 tasks from a queue, class 0 first.
#ifndef HB_DBG_NOK1
#define HB_DBG_NOK1 0
#endif
            const int ntasks = HB_DBG_NOK1 ? 0 : NCLS * nslots;
            for (int i = ntasks, first = 1, half = 0, blk0 = 0;; first = 0) {
                if (ONEVAR) {  // uniform tasks: static assignment
                    i = first ? tid : i + NT;
                    if (i >= ntasks) break;
                } else {  // blocks of 64 tasks: lane l takes tasks l and l + 32 of the block
                    if (first || half || HB_DPULL == 32) {
                        if ((tid & 31) == 0) blk0 = atomicAdd(&s_task, HB_DPULL);
                        blk0 = __shfl_sync(0xffffffffu, blk0, 0);
                        if (blk0 >= ntasks) break;
                        half = 0;
                        i = blk0 + (tid & 31);
#if HB_DPREFETCH
                        // the lane's second task of the block: pull its stage-A cache lines (L2 hits of several hundred
                        // cycles) towards L1 while the first task is evaluated
                        if (i + 32 < ntasks) {
                            const int i2 = i + 32;
                            const int ci2 = (i2 >= nslots) + (i2 >= 2 * nslots) + (i2 >= 3 * nslots), slot2 = i2 - ci2 * nslots;
                            const int c2 = direct_find(w.c_slot0, w.smap, slot2);
                            const int j2 = w.c_j0[c2] + (slot2 - w.c_slot0[c2]);
#pragma unroll
                            for (int f = 0; f < kDAFields; ++f)
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(afield(ci2, f) + j2));
                            if (w.c_rot[c2] >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(D.tabrot + w.c_rot[c2] + j2));
                        }
#endif
                    } else {
                        half = 1;
                        i = blk0 + 32 + (tid & 31);
                    }
                    if (i >= ntasks) continue;
                }
                const int ci = (i >= nslots) + (i >= 2 * nslots) + (i >= 3 * nslots), slot = i - ci * nslots;
                const int c = direct_find(w.c_slot0, w.smap, slot);
                const int m = w.c_mat[c], j = w.c_j0[c] + (slot - w.c_slot0[c]);
                HB_ASSERT(c >= 0 && c < nch && j >= 0 && j < NF && slot >= 0 && slot < CAP && j < s_jall && j < s_J[m] &&
                          ci >= 0 && ci < NCLS && slot - w.c_slot0[c] < w.c_len[c]);
                const double T = w.c_T[c], lsm = w.c_lsm[c];
                const double v = gc.eta * (double)j;
                StageA a;
                cplx l0;
                {
                    const double2 f0 = afield(ci, 0)[j], f1 = afield(ci, 1)[j], f2 = afield(ci, 2)[j],
                                  f3 = afield(ci, 3)[j], f4 = afield(ci, 4)[j];
                    a.num = {f0.x, f0.y};
                    l0 = {f1.x, f1.y};
                    a.d = {f2.x, f2.y};
                    a.g = {f3.x, f3.y};
                    a.q1 = {f4.x, f4.y};
                }
                // z_j = phi_j tab_j W^{j m_c}: weight and rotation to the centre of the maturity's bins
                cplx tabrot;
                if (w.c_rot[c] >= 0) {  // precomputed per distinct centre (hb_surface_set)
                    const double2 tr = D.tabrot[w.c_rot[c] + j];
                    tabrot = {tr.x, tr.y};
                } else {
                    const double2 tb = D.tab[j];
                    const double2 rt = D.tw[((unsigned)j * (unsigned)D.mat_c2[m]) & tw_mask];
                    tabrot = cmul({tb.x, tb.y}, {rt.x, -rt.y});
                }
                const StageB b = stage_b_auto(a, l0, T, true);
                if (!ONEVAR && ci == 0) {
                    const int vs[3] = {0, 2, 5};
                    double er[3], ei[3];
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const double kts = s_kts[vs[g]], v0s = s_v0s[vs[g]];
                        er[g] = fma(kts, b.B.re, fma(v0s, b.Dq.re, -(gc.ui * lsm)));  // stage_f, heston.cpp:87-91
                        ei[g] = fma(kts, b.B.im, fma(v0s, b.Dq.im, v * lsm));
                    }
                    double pr[3] = {0.0, 0.0, 0.0}, pi[3] = {0.0, 0.0, 0.0};
                    if (!(er[0] < gc.cut && er[1] < gc.cut && er[2] < gc.cut)) {
                        const double ks[3] = {s_kts[0], s_kts[2], s_kts[5]}, vs3[3] = {s_v0s[0], s_v0s[2], s_v0s[5]};
                        class0_cexp(er, ei, b, ks, vs3, pr, pi);  // one cexp + two expansions where the steps are small
                    }
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const bool zero = er[g] < gc.cut;  // exactly 0 as in stage_f
                        const cplx z = cmul({zero ? 0.0 : pr[g], zero ? 0.0 : pi[g]}, tabrot);
                        ybuf[(size_t)vs[g] * CAP + slot] = make_double2(z.re, z.im);
                    }
                } else {
                    const int vv = direct_cls_variant(ci);
                    const SliceConst sc = {s_kts[vv], s_v0s[vv], lsm};
                    const cplx z = cmul(stage_f(b, sc, v, gc.ui, nullptr, gc.cut), tabrot);
                    ybuf[(size_t)vv * CAP + slot] = make_double2(z.re, z.im);
                }
            }
            __syncthreads();
            // every warp has left K1 and none re-enters it before the next barrier: reset the task queue here
            if (tid == 0) s_task = 0;
            // ---- direct sums: item = (chunk, conjugate pair), all V slices ----
#ifndef HB_DBG_NODFT
#define HB_DBG_NODFT 0
#endif
            for (int it = tid; it < (HB_DBG_NODFT ? 0 : nitems); it += NT) {
                const int c = direct_find(w.c_item0, w.imap, it);
                const int m = w.c_mat[c], pi_ = it - w.c_item0[c];
                const unsigned d2 = (unsigned)D.pair_d2[D.pair_off[m] + pi_];
                const int len = w.c_len[c];
                const unsigned k0 = ((unsigned)w.c_j0[c] * d2) & tw_mask;
                // two interleaved twiddle recurrences (even / odd points), each advanced by W^{2 mu}
                const double2 w2 = D.tw[(2u * d2) & tw_mask];
                double2 t0 = D.tw[k0], t1 = D.tw[(k0 + d2) & tw_mask];
                double aP[V], aQ[V];
#pragma unroll
                for (int v = 0; v < V; ++v) aP[v] = aQ[v] = 0.0;
                HB_ASSERT(c >= 0 && c < nch && it < ITEMS && pi_ >= 0 && pi_ < MAXP && len >= 1 && len <= kDChunk &&
                          w.c_slot0[c] + len <= CAP && pi_ < D.pair_off[m + 1] - D.pair_off[m]);
                const double2* yb = ybuf + w.c_slot0[c];
                int k = 0;
                for (; k + 1 < len; k += 2) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const double2 y0 = yb[(size_t)v * CAP + k], y1 = yb[(size_t)v * CAP + k + 1];
                        aP[v] = fma(y0.x, t0.x, aP[v]);
                        aQ[v] = fma(y0.y, t0.y, aQ[v]);
                        aP[v] = fma(y1.x, t1.x, aP[v]);
                        aQ[v] = fma(y1.y, t1.y, aQ[v]);
                    }
                    const double2 n0 = {fma(t0.x, w2.x, -(t0.y * w2.y)), fma(t0.x, w2.y, t0.y * w2.x)};
                    const double2 n1 = {fma(t1.x, w2.x, -(t1.y * w2.y)), fma(t1.x, w2.y, t1.y * w2.x)};
                    t0 = n0;
                    t1 = n1;
                }
                if (k < len) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const double2 y0 = yb[(size_t)v * CAP + k];
                        aP[v] = fma(y0.x, t0.x, aP[v]);
                        aQ[v] = fma(y0.y, t0.y, aQ[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < V; ++v) part[(size_t)v * ITEMS + it] = make_double2(aP[v], aQ[v]);
            }
            __syncthreads();
            // ---- finish: chunk partials of a maturity added in chunk order (P and Q separately), X = P +- Q at the two
            // bracketing bins, interpolation, clamp, parity.  A maturity cut by the wave leaves its running sums in
            // `carry` (by wave parity: the first piece of a wave may read what its predecessor left while the last
            // piece leaves its own).  No barrier up to the next K1: that one only writes ybuf. ----
            {
                const double2* carry_in = carry + (size_t)(wi & 1) * V * MAXP;
                double2* carry_out = carry + (size_t)((wi + 1) & 1) * V * MAXP;
                auto pair_sum = [&](int sg, int v, int pi_) -> double2 {
                    double2 acc = (w.s_flags[sg] & 1) ? make_double2(0.0, 0.0) : carry_in[v * MAXP + pi_];
                    HB_ASSERT(sg >= 0 && sg < w.nseg && v >= 0 && v < V && pi_ >= 0 && pi_ < MAXP && w.s_c1[sg] <= w.nch);
                    for (int cc = w.s_c0[sg]; cc < w.s_c1[sg]; ++cc) {
                        HB_ASSERT(w.c_item0[cc] + pi_ < w.c_item0[cc + 1] && w.c_item0[cc] + pi_ < ITEMS);
                        const double2 q = part[(size_t)v * ITEMS + w.c_item0[cc] + pi_];
                        acc.x += q.x;
                        acc.y += q.y;
                    }
                    return acc;
                };
                // tasks of all pieces flattened: (slice v, option) of an ending piece, (slice v, pair) of a cut one
#ifndef HB_DBG_NOFIN
#define HB_DBG_NOFIN 0
#endif
                const int ntask = HB_DBG_NOFIN ? 0 : w.s_t0[w.nseg];
                for (int i = tid; i < V * ntask; i += NT) {
                    int v = 0, t = i;
                    if (V > 1) {
                        v = (i >= ntask) + (i >= 2 * ntask) + (i >= 3 * ntask) + (i >= 4 * ntask) + (i >= 5 * ntask);
                        t = i - v * ntask;
                    }
                    int sg = 0;
                    while (w.s_t0[sg + 1] <= t) ++sg;
                    t -= w.s_t0[sg];
                    const int m = w.s_mat[sg];
                    if (!(w.s_flags[sg] & 2)) {  // only the last piece of a wave: keep the running sums
                        carry_out[v * MAXP + t] = pair_sum(sg, v, t);
                        continue;
                    }
                    const int o = S.mat_off[m] + t;
                    const int q0 = D.opt_pq0[o], q1 = D.opt_pq1[o];
                    double price = __longlong_as_double(0x7ff8000000000000LL);
                    if (q0 >= 0) {
                        HB_ASSERT((q0 >> 1) < MAXP && (q1 >> 1) < MAXP);
                        const double2 a0 = pair_sum(sg, v, q0 >> 1), a1 = pair_sum(sg, v, q1 >> 1);
                        const double x0 = (q0 & 1) ? a0.x + a0.y : a0.x - a0.y;
                        const double x1 = (q1 & 1) ? a1.x + a1.y : a1.x - a1.y;
                        const double c0 = S.opt_s0[o] * x0, c1 = S.opt_s1[o] * x1;
                        const double call = S.mat_disc[m] * (c0 + (c1 - c0) * S.opt_frac[o]);
                        price = finish_price(call, S.opt_call[o] != 0, S.mat_fwd[m], S.opt_kdisc[o]);
                    }
                    HB_ASSERT(o >= S.mat_off[m] && o < S.mat_off[m + 1] && S.opt_orig[o] >= 0 && S.opt_orig[o] < n);
                    rows[(size_t)v * n + S.opt_orig[o]] = price;
                }
            }
#ifdef HB_DBG_SYNC
            __syncthreads();
#endif
            if (!w.more) break;
        }
        __syncthreads();  // the price rows of the job are complete
#ifndef HB_DBG_NOFINAL
#define HB_DBG_NOFINAL 0
#endif
        if (pieces == 1 && !HB_DBG_NOFINAL) finalize_job<NT, kFinalizeT<NT>()>(what, rows, S, js, p, out, out2, red, tid);
    }
}

// Prefix lengths of every (parameter set, maturity), one warp per set: the table the direct kernel reads, and the
// routing of the sets by mean prefix length -- short ones to the direct kernel, long ones (slow decay: the direct sums
// cost O(prefix x bins), the transform O(N log N) whatever the prefix) to fft_job_kernel.  The order of the two lists
// is not deterministic (atomics); the result of a set does not depend on where it is priced in a list.
constexpr int kScanWarps = 4;
template <bool ONEVAR>
__global__ void __launch_bounds__(32 * kScanWarps)
prefix_scan_kernel(SurfaceDev S, DirectDev D, GridConst gc, Bounds bd, const double* __restrict__ params, int ld, int P,
                   int* __restrict__ jtab, int threshold, int* __restrict__ short_ids, int* __restrict__ long_ids,
                   int* __restrict__ counts) {
    constexpr int V = ONEVAR ? 1 : 6, NCLS = ONEVAR ? 1 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb2 = D.nblk + direct_nsuper(D.nblk);  // fine + super blocks of one class at a time
    PrefixBlock* blocks = reinterpret_cast<PrefixBlock*>(smem_raw) + (size_t)warp * nb2;
    __shared__ JobState jss[kScanWarps];
    __shared__ double s_kts[kScanWarps][6], s_v0s[kScanWarps][6];
    __shared__ int s_jm[kScanWarps][kDMaxMat];
    JobState& js = jss[warp];
    const int M = S.n_mat;
    for (int p = blockIdx.x * kScanWarps + warp; p < P; p += gridDim.x * kScanWarps) {
        __syncwarp();
        if (lane == 0) job_setup(js, params, ld, p, bd, V);
        __syncwarp();
        if (!js.valid) {
            for (int m = lane; m < M; m += 32) jtab[(size_t)p * M + m] = 1;
            if (lane == 0) short_ids[atomicAdd(&counts[0], 1)] = p;
            continue;
        }
        if (lane < V) {
            const int v = lane;
            const int cls = (v == 1 || v == 3 || v == 4) ? v : 0;
            const double s2 = js.x[cls][2] * js.x[cls][2];
            s_kts[warp][v] = js.x[v][0] * js.x[v][1] / s2;
            s_v0s[warp][v] = js.x[v][4] / s2;
        }
        for (int m = lane; m < M; m += 32) s_jm[warp][m] = 0;
        for (int ci = 0; ci < NCLS; ++ci) {
            __syncwarp();
            {
                const double* xc = js.x[direct_cls_variant(ci)];
                const ClassConst cc = {xc[0], xc[2] * xc[2], xc[3] * xc[2]};
                const PrefixClass pc = prefix_class(cc, gc.alpha);
                for (int k = lane; k < nb2; k += 32) blocks[k] = direct_block_at(pc, D.blk, D.nblk, k, gc.eta);
            }
            __syncwarp();
            double kts[3], v0s[3];
            int nv = 1;
            if (ci == 0) {
                kts[0] = s_kts[warp][0];
                v0s[0] = s_v0s[warp][0];
                if (V > 1) {
                    kts[1] = s_kts[warp][2];
                    v0s[1] = s_v0s[warp][2];
                    kts[2] = s_kts[warp][5];
                    v0s[2] = s_v0s[warp][5];
                    nv = 3;
                }
            } else {
                const int v = direct_cls_variant(ci);
                kts[0] = s_kts[warp][v];
                v0s[0] = s_v0s[warp][v];
            }
            for (int m = lane; m < M; m += 32) {  // one maturity per lane
                const double T = S.mat_T[m];
                const double cst = -gc.ui * (S.ln_spot + (S.rate - S.dividend) * T);
                const int J = direct_prefix(blocks, blocks + D.nblk, D.blk, D.nblk, T, cst, kts, v0s, nv, gc.cut);
                s_jm[warp][m] = max(s_jm[warp][m], min(J, D.n_full));
            }
        }
        __syncwarp();
        long long total = 0;
        for (int m = lane; m < M; m += 32) {
            const int J = s_jm[warp][m];
            jtab[(size_t)p * M + m] = J;
            total += J;
        }
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) {
            if (total <= (long long)threshold * M) short_ids[atomicAdd(&counts[0], 1)] = p;
            else long_ids[atomicAdd(&counts[1], 1)] = p;
        }
    }
}

}  // namespace hb
