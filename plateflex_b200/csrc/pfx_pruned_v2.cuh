// pfx_pruned_v2.cuh -- second-generation kernels of the band-pruned engine for padded axes of
// N >= 2048 samples (grids from 512 cells per side up).  Same mathematics as pfx_pruned_kernels.cuh
// (see pfx_cwt_pruned.cu), different mapping:
//
//   * a line of N = L*r output samples is produced in parts of CH = max(2048, L) elements, each part the
//     L-point FFTs of RC = CH/L consecutive residues.  Parts are independent given the K support samples,
//     so chunking costs no arithmetic and keeps a line part at 32 KB of shared memory whatever N is
//     (3 CTAs of 256 threads per SM instead of 1 at N = 8192);
//   * with CH fixed every thread owns the same radix-8 butterfly (u, rho) in every part of every line:
//     all shared-memory indices, strides and digit reversals are compile-time constants, and the K
//     support samples a thread needs are the same for every part, so they are fetched from global
//     memory once per line straight into registers -- optionally one line group ahead (PREFETCH) so the
//     load latency hides behind the remaining stages of the current group;
//   * the first-stage twiddles w_N^(a (rho+1)) advance from part to part by one complex multiply
//     (rho -> rho + RC) instead of table reads.
#pragma once
#include "pfx_pruned.hpp"
#include "pfx_pruned_kernels.cuh"

namespace pfx {
namespace v2 {

using prn::cadd;
using prn::cmul;
using prn::csub;
using prn::dft;

constexpr int NT = 256;

template <int LOGL>
struct Geo {
    static_assert(LOGL >= 5 && LOGL <= 11, "v2 kernels need 32 <= L <= 2048");
    static constexpr int LOGCH = LOGL > 11 ? LOGL : 11;
    static constexpr int LOGRC = LOGCH - LOGL;
    static constexpr int RC = 1 << LOGRC;
    static constexpr int L = 1 << LOGL;
    static constexpr bool PAD = LOGRC < 3;
    static constexpr int NST = prn_nstages(LOGL);
    static constexpr int PLANE = (PAD ? (L + (L >> 3)) : L) << LOGRC;  // double2 elements per line part
    static constexpr int LR0 = prn_lrad(LOGL, 0);
    static_assert(LR0 == 3, "first stage is radix 8");
    static constexpr int LOGBF = LOGL - 3;          // log2 of first-stage butterflies per FFT
    static constexpr int STEP = 1 << LOGBF;         // distance between the inputs of a first-stage butterfly
    static constexpr int IPT0 = (1 << LOGCH) / 8 / NT;  // first-stage butterflies per thread and part
    static constexpr int STW = prn_stw_total(LOGL);
};

template <bool PAD>
__device__ __forceinline__ constexpr int padded(int e)
{
    return PAD ? e + (e >> 3) : e;
}

struct V2Args {
    PrunedArgs p;
    int logr;     // log2(N / L) of the transformed axis
    int nparts;   // N / CH
};

// twiddle of stage S >= 1 for sub-FFT position q0: w_L'^{q0}, L' = M*R  -> root index q0 << (logN - logm - lr)
template <int LOGL>
__device__ __forceinline__ void load_stage_twiddles(double2 *stw, int logN, const double2 *__restrict__ root)
{
#pragma unroll
    for (int s = 1; s < prn_nstages(LOGL); s++) {
        const int logm = prn_logm(LOGL, s);
        const int rs = logN - (logm + prn_lrad(LOGL, s));
        for (int q0 = threadIdx.x; q0 < (1 << logm); q0 += NT) stw[prn_stw_off(LOGL, s) + q0] = __ldg(root + (q0 << rs));
    }
}

// a <- a + w b, b <- a - w b in six fused multiply-adds (the difference as 2a - (a + w b)).
__device__ __forceinline__ void bfly(double2 &a, double2 &b, const double2 w)
{
    const double pr = fma(w.x, b.x, fma(-w.y, b.y, a.x));
    const double pi = fma(w.x, b.y, fma(w.y, b.x, a.y));
    b = make_double2(fma(2.0, a.x, -pr), fma(2.0, a.y, -pi));
    a = make_double2(pr, pi);
}
// a <- a + w b only (four fused multiply-adds): the last stage of a line whose upper half is never stored
__device__ __forceinline__ void bfly_plus(double2 &a, const double2 b, const double2 w)
{
    a = make_double2(fma(w.x, b.x, fma(-w.y, b.y, a.x)), fma(w.x, b.y, fma(w.y, b.x, a.y)));
}
__device__ __forceinline__ double2 csqr(const double2 w) { return make_double2(fma(w.x, w.x, -(w.y * w.y)), 2.0 * (w.x * w.y)); }
__device__ __forceinline__ double2 times_i(const double2 w) { return make_double2(-w.y, w.x); }

// Twiddled radix-R butterfly, X[k] = sum_j W_R^{jk} w^j v[j] (W_R = exp(+2 pi i / R)), as log2(R) levels of
// radix-2 butterflies whose twiddles are w^(R/2), ..., w^2, w times the fixed eighth roots: 12 x 6 + 10
// instructions for R = 8 where "multiply the inputs by w^j, then a plain DFT" takes 108.  Output k is left in
// v[bitrev(k)]; with HALF only X[0 .. R/2) are computed (v[bitrev(k)] for k >= R/2 is garbage).
template <int R, bool HALF>
__device__ __forceinline__ void tw_dft(double2 (&v)[R], const double2 w)
{
    if constexpr (R == 8) {
        constexpr double H = 0.70710678118654752440;
        const double2 w2 = csqr(w), w4 = csqr(w2);
#pragma unroll
        for (int j = 0; j < 4; j++) bfly(v[j], v[j + 4], w4);
        const double2 iw2 = times_i(w2);
        bfly(v[0], v[2], w2);
        bfly(v[1], v[3], w2);
        bfly(v[4], v[6], iw2);
        bfly(v[5], v[7], iw2);
        const double2 w8 = make_double2(H * (w.x - w.y), H * (w.x + w.y));  // w * exp(i pi / 4)
        if constexpr (HALF) {
            bfly_plus(v[0], v[1], w);
            bfly_plus(v[4], v[5], w8);
            bfly_plus(v[2], v[3], times_i(w));
            bfly_plus(v[6], v[7], times_i(w8));
        } else {
            bfly(v[0], v[1], w);
            bfly(v[4], v[5], w8);
            bfly(v[2], v[3], times_i(w));
            bfly(v[6], v[7], times_i(w8));
        }
    } else {
        static_assert(R == 4, "radix 4 or 8");
        const double2 w2 = csqr(w);
        bfly(v[0], v[2], w2);
        bfly(v[1], v[3], w2);
        if constexpr (HALF) {
            bfly_plus(v[0], v[1], w);
            bfly_plus(v[2], v[3], times_i(w));
        } else {
            bfly(v[0], v[1], w);
            bfly(v[2], v[3], times_i(w));
        }
    }
}
template <int R>
__device__ __forceinline__ constexpr int bitrev(int k)
{
    return R == 8 ? (((k & 1) << 2) | (k & 2) | ((k >> 2) & 1)) : (((k & 1) << 1) | ((k >> 1) & 1));
}

// Stage S >= 1 of all CPC lines of a part, in place; the last stage stores through store(c, q, rho, v).
// HALF (last stage only): the upper half of the line is never stored (N = 2 nout), so it is not computed.
template <int LOGL, int S, int CPC, bool HALF, class Store>
__device__ __forceinline__ void stage(double2 *sm, const double2 *stw_all, const Store &store)
{
    using G = Geo<LOGL>;
    constexpr int LR = prn_lrad(LOGL, S), R = 1 << LR, LOGM = prn_logm(LOGL, S), M = 1 << LOGM;
    constexpr bool LAST = (S == G::NST - 1);
    constexpr int NBF = (1 << G::LOGCH) >> LR;  // butterflies per line part
    constexpr int IPT = NBF / NT;
    static_assert(IPT >= 1, "a stage has at least one butterfly per thread");
    const double2 *stw = stw_all + prn_stw_off(LOGL, S);
#pragma unroll
    for (int it = 0; it < IPT; it++) {
        const int b = it * NT + threadIdx.x;
        const int rho = b & (G::RC - 1);
        const int u = b >> G::LOGRC;
        const int q0 = u & (M - 1);
        const int base = ((u >> LOGM) << (LOGM + LR)) + q0;
        const double2 w = stw[q0];
#pragma unroll
        for (int c = 0; c < CPC; c++) {
            double2 *p = sm + c * G::PLANE + rho;
            double2 v[R];
#pragma unroll
            for (int k = 0; k < R; k++) v[k] = p[padded<G::PAD>(base + k * M) << G::LOGRC];
            if constexpr (LAST) {
                if (HALF) {
                    tw_dft<R, true>(v, w);
#pragma unroll
                    for (int k = 0; k < R / 2; k++) store(c, base + k * M, rho, v[bitrev<R>(k)]);
                } else {
                    tw_dft<R, false>(v, w);
#pragma unroll
                    for (int k = 0; k < R; k++) store(c, base + k * M, rho, v[bitrev<R>(k)]);
                }
            } else {
                tw_dft<R, false>(v, w);
#pragma unroll
                for (int k = 0; k < R; k++) p[padded<G::PAD>(base + k * M) << G::LOGRC] = v[bitrev<R>(k)];
            }
        }
    }
}

template <int LOGL, int S, int CPC, bool HALF, class Store>
__device__ __forceinline__ void run_stages(double2 *sm, const double2 *stw, const Store &store)
{
    if constexpr (S < Geo<LOGL>::NST) {
        __syncthreads();
        stage<LOGL, S, CPC, HALF, Store>(sm, stw, store);
        run_stages<LOGL, S + 1, CPC, HALF, Store>(sm, stw, store);
    }
}

// Shared driver of both passes.  grp.load(g, c, m) returns the weighted support sample m of line c of group g;
// grp.store(g, c, q, rho_global, v) consumes output element q of residue rho_global of that line.
// MULTI: the line has more than one part (N > CH); then the support samples stay in registers across the
// parts and the first-stage twiddles advance by one complex multiply per part.  PREFETCH: the samples of the
// next group are fetched while the later stages of the current one run.
template <int LOGL, int CPC, bool MULTI, bool PREFETCH, class Group>
__device__ __forceinline__ void run_lines(double2 *sm, const double2 *stw, const AxisDesc &A, int nparts, int lo, int K,
                                          const double2 *__restrict__ root, Group &grp)
{
    using G = Geo<LOGL>;
    static_assert(G::IPT0 == 1, "v2 kernels: one first-stage butterfly per thread (L <= 2048)");
    constexpr int STEP = G::STEP, L = G::L;
    constexpr bool KEEP = MULTI || PREFETCH;  // support samples persist in registers
    const int Nm = A.N - 1;
    const bool half = 2 * A.nout == A.N;  // power-of-two grid: outputs q >= L/2 are never stored
    const int l0 = lo & (L - 1), lh = l0 >> G::LOGBF, ll = l0 & (STEP - 1);

    // this thread's first-stage butterfly (u, rho): fixed for the whole kernel
    const int rho = threadIdx.x & (G::RC - 1), u = threadIdx.x >> G::LOGRC;
    const int nlow = prn::digit_reverse<LOGL>(u);
    const int m0 = (nlow - lo) & (L - 1);  // support index of input 0
    const bool flag = nlow >= ll;          // which of the two candidate positions holds the wrap of the support index
    const int a0 = lo + m0;
    double2 *const out0 = sm + (padded<G::PAD>(u << 3) << G::LOGRC) + rho;
    // t0 = w_N^(a0 (rho+1)); g = w_N^(STEP (rho+1)): ratio between consecutive inputs; h = w_N^(-L (rho+1)): wrap
    const double2 t0i = __ldg(root + (int)(((long long)a0 * (rho + 1)) & Nm));
    const double2 g0i = __ldg(root + (((rho + 1) << G::LOGBF) & Nm));
    const double2 h0i = __ldg(root + ((A.N - (((rho + 1) << LOGL) & Nm)) & Nm));

    double2 x[KEEP ? CPC : 1][8];
    auto fetch = [&](int gi) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int m = (m0 + k * STEP) & (L - 1);
#pragma unroll
            for (int c = 0; c < CPC; c++) x[KEEP ? c : 0][k] = (m < K) ? grp.load(gi, c, m) : make_double2(0., 0.);
        }
    };

    int gi = grp.first();
    if (PREFETCH && grp.valid(gi)) fetch(gi);
    for (; grp.valid(gi); gi = grp.next(gi)) {
        if (KEEP && !PREFETCH) fetch(gi);
        double2 t0 = t0i, g = g0i, h = h0i;
        for (int part = 0; part < (MULTI ? nparts : 1); part++) {
            // ---- first stage: twiddle the support samples, radix-8 butterfly into shared memory
            const double2 one = make_double2(1., 0.);
            const double2 ha = flag ? h : one, hb = flag ? one : h;
#pragma unroll
            for (int c = 0; c < CPC; c++) {
                double2 v[8];
                double2 t = t0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k > 0) {
                        t = cmul(t, g);
                        if (k == lh) t = cmul(t, ha);
                        if (k == lh + 1) t = cmul(t, hb);
                    }
                    double2 xin;
                    if (KEEP) {
                        xin = x[c][k];
                    } else {
                        const int m = (m0 + k * STEP) & (L - 1);
                        xin = (m < K) ? grp.load(gi, c, m) : make_double2(0., 0.);
                    }
                    v[k] = cmul(xin, t);
                }
                dft<8>(v);
#pragma unroll
                for (int k = 0; k < 8; k++) out0[c * G::PLANE + (k << G::LOGRC)] = v[k];
            }
            if (MULTI && part + 1 < nparts) {
                // rho -> rho + RC
                t0 = cmul(t0, __ldg(root + (int)(((long long)a0 * G::RC) & Nm)));
                g = cmul(g, __ldg(root + ((G::RC << G::LOGBF) & Nm)));
                h = cmul(h, __ldg(root + ((A.N - ((G::RC << LOGL) & Nm)) & Nm)));
            }
            if (PREFETCH && part == (MULTI ? nparts : 1) - 1) {
                const int gn = grp.next(gi);
                if (grp.valid(gn)) fetch(gn);
            }
            const int rho0 = part << G::LOGRC;
            auto store = [&](int c, int q, int rr, double2 v) { grp.store(gi, c, q, rho0 + rr, v); };
            if (half)
                run_stages<LOGL, 1, CPC, true>(sm, stw, store);
            else
                run_stages<LOGL, 1, CPC, false>(sm, stw, store);
            __syncthreads();  // the next first stage overwrites the planes
        }
    }
}

// ---- pass 2: x-direction transforms, CPC consecutive output rows per group ---------------------------------
template <int CPC>
struct Pass2Group {
    const double2 *Tt;  // T[t][m][j]
    double2 *Wt;        // W[t][j][i]
    int ny, nx, ngroups, logr;
    __device__ __forceinline__ int first() const { return blockIdx.x; }
    __device__ __forceinline__ bool valid(int g) const { return g < ngroups; }
    __device__ __forceinline__ int next(int g) const { return g + gridDim.x; }
    __device__ __forceinline__ double2 load(int g, int c, int m) const
    {
        const int j = g * CPC + c;
        return (CPC == 1 || j < ny) ? __ldg(Tt + (size_t)m * ny + j) : make_double2(0., 0.);
    }
    __device__ __forceinline__ void store(int g, int c, int q, int rho, double2 v) const
    {
        const int i = (q << logr) + rho, j = g * CPC + c;
        if (i < nx && (CPC == 1 || j < ny)) Wt[(size_t)j * nx + i] = v;
    }
};

template <int LOGL, int CPC, bool MULTI, bool PREFETCH>
__global__ void __launch_bounds__(NT, (CPC == 1 && !PREFETCH && !MULTI) ? 3 : 2) pass2_kernel(V2Args a)
{
    using G = Geo<LOGL>;
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ax;
    const int t = blockIdx.y, ia = t % PFX_NA;
    double2 *stw = smem + CPC * G::PLANE;
    load_stage_twiddles<LOGL>(stw, A.logN, a.p.rootx);
    Pass2Group<CPC> grp;
    grp.Tt = a.p.T + (size_t)t * sd.kamax * a.p.ny;
    grp.Wt = a.p.W + (size_t)t * a.p.ny * a.p.nx;
    grp.ny = a.p.ny;
    grp.nx = a.p.nx;
    grp.ngroups = (a.p.ny + CPC - 1) / CPC;
    grp.logr = a.logr;
    __syncthreads();
    run_lines<LOGL, CPC, MULTI, PREFETCH>(smem, stw, A, a.nparts, sd.alo[ia], sd.ka[ia], a.p.rootx, grp);
}

// ---- pass 1: y-direction transforms, CPC consecutive spectrum columns per group ---------------------------
template <int CPC>
struct Pass1Group {
    const double2 *spec;  // padded spectrum [a][b] (y-wavenumber fastest)
    const double *tv;     // v1 weights of this azimuth [L]
    const double *tu;     // (c/P) u1 weights of this azimuth [Lx]
    double2 *T;           // intermediate (layout: prn_t_column)
    const PrunedArgs *pa;
    int t;
    int NX, NYm, ny, alo, ka, blo, ngroups, logr;
    __device__ __forceinline__ int first() const { return blockIdx.x; }
    __device__ __forceinline__ bool valid(int g) const { return g < ngroups; }
    __device__ __forceinline__ int next(int g) const { return g + gridDim.x; }
    __device__ __forceinline__ double2 load(int g, int c, int mb) const
    {
        const int m = g * CPC + c;
        if (CPC > 1 && m >= ka) return make_double2(0., 0.);
        const double2 x = __ldg(spec + (size_t)((alo + m) & (NX - 1)) * (NYm + 1) + ((blo + mb) & NYm));
        const double wv = tv[mb];
        return make_double2(x.x * wv, x.y * wv);
    }
    __device__ __forceinline__ void store(int g, int c, int q, int rho, double2 v) const
    {
        const int j = (q << logr) + rho, m = g * CPC + c;
        if (j < ny && (CPC == 1 || m < ka)) {
            const double wu = tu[m];
            size_t js;
            const size_t o = prn_t_column(*pa, t, m, alo, js);
            T[o + (size_t)j * js] = make_double2(v.x * wu, v.y * wu);
        }
    }
};

template <int LOGL, int CPC, bool MULTI, bool PREFETCH>
__global__ void __launch_bounds__(NT, (CPC == 1 && !PREFETCH && !MULTI) ? 3 : 2) pass1_kernel(V2Args a)
{
    using G = Geo<LOGL>;
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ay;
    const int t = blockIdx.y, ia = t % PFX_NA, g = t / PFX_NA;
    const int ngroups = (sd.ka[ia] + CPC - 1) / CPC;
    if ((int)blockIdx.x >= ngroups) return;
    double2 *stw = smem + CPC * G::PLANE;
    load_stage_twiddles<LOGL>(stw, A.logN, a.p.rooty);
    Pass1Group<CPC> grp;
    grp.spec = g ? a.p.spec1 : a.p.spec0;
    grp.tv = a.p.tabv + sd.tabv + ((size_t)ia << LOGL);
    grp.tu = a.p.tabu + sd.tabu + (size_t)ia * sd.ax.L;
    grp.T = a.p.T;
    grp.pa = &a.p;
    grp.t = t;
    grp.NX = a.p.NX;
    grp.NYm = a.p.NY - 1;
    grp.ny = a.p.ny;
    grp.alo = sd.alo[ia];
    grp.ka = sd.ka[ia];
    grp.blo = sd.blo[ia];
    grp.ngroups = ngroups;
    grp.logr = a.logr;
    __syncthreads();
    run_lines<LOGL, CPC, MULTI, PREFETCH>(smem, stw, A, a.nparts, sd.blo[ia], sd.kb[ia], a.p.rooty, grp);
}

template <int LOGL>
inline size_t smem_bytes(int cpc)
{
    return ((size_t)cpc * Geo<LOGL>::PLANE + Geo<LOGL>::STW) * sizeof(double2);
}

}  // namespace v2

// launchers (pfx_pruned_v2_p1.cu, pfx_pruned_v2_p2.cu).  variant = cpc | (prefetch << 4) | (multi << 5)
// resident: receives the number of CTAs of this kernel that fit on one SM
int v2_configure_p1(int logL, int variant, int *resident);
int v2_configure_p2(int logL, int variant, int *resident);
int v2_launch_p1(const v2::V2Args &a, int logL, int variant, dim3 grid, cudaStream_t st);
int v2_launch_p2(const v2::V2Args &a, int logL, int variant, dim3 grid, cudaStream_t st);
size_t v2_smem_bytes(int logL, int cpc);

}  // namespace pfx
