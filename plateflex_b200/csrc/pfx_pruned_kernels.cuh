// pfx_pruned_kernels.cuh -- the shared-memory FFT kernels of the band-pruned engine,
// specialised at compile time on the FFT length (LOGL), the lines per CTA (CPC) and the
// padding mode so that every stage's radix, stride and digit reversal is a constant.
// See pfx_cwt_pruned.cu for the algorithm.
#pragma once
#include "pfx_pruned.hpp"

namespace pfx {
namespace prn {

constexpr int NT = PRN_NT;

__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-register inverse DFT (exp(+i...)) of R = 2, 4, 8 points, natural order in and out.
template <int R>
__host__ __device__ __forceinline__ void dft(double2 (&v)[R])
{
    if constexpr (R == 2) {
        const double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else {
        double2 e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; k++) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        dft<R / 2>(e);
        dft<R / 2>(o);
#pragma unroll
        for (int k = 0; k < R / 2; k++) {
            constexpr double H = 0.70710678118654752440;
            double2 t;
            const int k8 = k * (8 / R);  // twiddle exp(2 pi i k8 / 8)
            if (k8 == 0) t = o[k];
            else if (k8 == 2) t = make_double2(-o[k].y, o[k].x);
            else if (k8 == 1) t = make_double2(H * (o[k].x - o[k].y), H * (o[k].x + o[k].y));
            else t = make_double2(-H * (o[k].x + o[k].y), H * (o[k].x - o[k].y));
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
}

template <bool PAD>
__device__ __forceinline__ int padded(int e)
{
    return PAD ? e + (e >> 3) : e;
}

struct SmemTables {
    const double2 *stw;  // stage s (>= 1) at stw + prn_stw_off(LOGL, s), 2^logm(s) entries: w_{M*R}^{q0}
    const double2 *g;    // [rc]  w_N^{(L/R1)(rho+1)}: step between the R1 inputs of a first-stage butterfly
    const double2 *h;    // [rc]  w_N^{-L(rho+1)}: correction when the support index wraps modulo L
};

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

template <int LOGL>
__device__ __forceinline__ void load_residue_factors(double2 *gt, double2 *ht, const AxisDesc &A, int rho0,
                                                     const double2 *__restrict__ root)
{
    for (int rho = threadIdx.x; rho < A.rc; rho += NT) {
        const int e = rho0 + rho + 1;
        gt[rho] = __ldg(root + ((e << (LOGL - prn_lrad(LOGL, 0))) & (A.N - 1)));
        ht[rho] = __ldg(root + ((A.N - ((e << LOGL) & (A.N - 1))) & (A.N - 1)));
    }
}

// Stage S >= 1: radix R butterflies combining sub-FFTs of size M = 2^LOGM.  Element e of the FFT
// of residue rho of line c sits at sm[c*plane + (padded(e) << logrc) + rho].  The last stage hands its
// results straight to store(c, i, value) with i = q*r + rho0 + rho the output sample (consecutive lanes
// hold consecutive i), skipping the shared-memory round trip; outputs beyond nout are dropped.
template <int LOGL, int S, bool PAD, bool LAST, class Store>
__device__ __forceinline__ void fft_stage(double2 *sm, int cpc, int plane, const AxisDesc &A, int rho0,
                                          const double2 *stw_all, const Store &store)
{
    constexpr int LR = prn_lrad(LOGL, S), R = 1 << LR, LOGM = prn_logm(LOGL, S), M = 1 << LOGM, LOGBF = LOGL - LR;
    const double2 *stw = stw_all + prn_stw_off(LOGL, S);
    const int logrc = A.logrc;
    const int total = cpc << (LOGBF + logrc);
    const int rcm = (1 << logrc) - 1;
    const int es = (PAD ? (M + (M >> 3)) : M) << logrc;  // element stride between butterfly inputs
    for (int w = threadIdx.x; w < total; w += NT) {
        const int rho = w & rcm;
        const int uc = w >> logrc;
        const int u = uc & ((1 << LOGBF) - 1);
        const int c = uc >> LOGBF;
        const int q0 = u & (M - 1);
        const int base = ((u >> LOGM) << (LOGM + LR)) + q0;
        double2 *p = sm + c * plane + (padded<PAD>(base) << logrc) + rho;
        double2 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = p[k * es];
        if (q0 != 0) {
            const double2 w1 = stw[q0];
            double2 t = w1;
            v[1] = cmul(v[1], t);
#pragma unroll
            for (int k = 2; k < R; k++) {
                t = cmul(t, w1);
                v[k] = cmul(v[k], t);
            }
        }
        dft<R>(v);
        if constexpr (LAST) {
            const int i0 = (base << A.logr) + rho0 + rho;  // base == q0 == u in the last stage
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int i = i0 + ((k * M) << A.logr);
                if (i < A.nout) store(c, i, v[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) p[k * es] = v[k];
        }
    }
}

template <int LOGL, int S, bool PAD, class Store>
__device__ __forceinline__ void run_stages(double2 *sm, int cpc, int plane, const AxisDesc &A, int rho0,
                                           const double2 *stw, const Store &store)
{
    if constexpr (S < prn_nstages(LOGL)) {
        __syncthreads();
        fft_stage<LOGL, S, PAD, (S == prn_nstages(LOGL) - 1), Store>(sm, cpc, plane, A, rho0, stw, store);
        run_stages<LOGL, S + 1, PAD, Store>(sm, cpc, plane, A, rho0, stw, store);
    }
}

// natural index of the first input of first-stage butterfly u: the digits of u (radices of stages
// 1.. in order, least significant first) re-assembled most significant first.
template <int LOGL>
__device__ __forceinline__ int digit_reverse(int u)
{
    int nlow = 0;
#pragma unroll
    for (int s = 1; s < prn_nstages(LOGL); s++) {
        const int lr = prn_lrad(LOGL, s);
        nlow = (nlow << lr) | (u & ((1 << lr) - 1));
        u >>= lr;
    }
    return nlow;
}

// First stage fused with the input twiddles.  load(c, m) returns the weighted input m = a - lo of
// line c.  The twiddle of input k of a butterfly is exp(2 pi i a_k (rho+1)/N) with a_k = lo +
// ((n_k - lo) mod L), n_k = nlow + k*L/R: one table read tw[nlow*r + rho] for k = 0, then
// t_k = t_{k-1} * g[rho], times h[rho] once when the support index wraps.
template <int LOGL, int CPC, bool PAD, class Load>
__device__ __forceinline__ void fft_first_stage(double2 *sm, const Load &load, const AxisDesc &A, int plane, int lo, int K,
                                                int rho0, const double2 *__restrict__ tw, const SmemTables &tb)
{
    constexpr int LR = prn_lrad(LOGL, 0), R = 1 << LR, LOGBF = LOGL - LR, STEP = 1 << LOGBF, L = 1 << LOGL;
    const int logrc = A.logrc;
    const int total = 1 << (LOGBF + logrc);
    const int rcm = (1 << logrc) - 1;
    for (int w = threadIdx.x; w < total; w += NT) {
        const int rho = w & rcm;
        const int u = w >> logrc;
        const int nlow = digit_reverse<LOGL>(u);
        const int m0 = (nlow - lo) & (L - 1);
        const double2 t0 = __ldg(tw + ((size_t)nlow << A.logr) + rho0 + rho);
        const double2 g = tb.g[rho], h = tb.h[rho];
        double2 *p = sm + (padded<PAD>(u << LR) << logrc) + rho;
#pragma unroll
        for (int c = 0; c < CPC; c++) {
            double2 v[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int m = (m0 + k * STEP) & (L - 1);
                v[k] = (m < K) ? load(c, m) : make_double2(0., 0.);
            }
            double2 t = t0;
            v[0] = cmul(v[0], t);
#pragma unroll
            for (int k = 1; k < R; k++) {
                t = cmul(t, g);
                if (((m0 + k * STEP) & (L - 1)) < STEP) t = cmul(t, h);  // wrapped between k-1 and k
                v[k] = cmul(v[k], t);
            }
            dft<R>(v);
#pragma unroll
            for (int k = 0; k < R; k++) p[c * plane + (k << logrc)] = v[k];
        }
    }
}

// All lines of the CTA: first stage into shared memory, remaining stages in place, the last one
// storing through store(c, i, value).  A single-stage plan (L <= 8) copies out of shared memory instead.
template <int LOGL, int CPC, bool PAD, class Load, class Store>
__device__ __forceinline__ void fft_lines(double2 *sm, const Load &load, const Store &store, const AxisDesc &A, int plane,
                                          int lo, int K, int rho0, const double2 *__restrict__ tw, const SmemTables &tb)
{
    fft_first_stage<LOGL, CPC, PAD>(sm, load, A, plane, lo, K, rho0, tw, tb);
    if constexpr (prn_nstages(LOGL) > 1) {
        run_stages<LOGL, 1, PAD, Store>(sm, CPC, plane, A, rho0, tb.stw, store);
    } else {
        __syncthreads();
        const int nq = (A.nout + A.r - 1) >> A.logr;
        const int rcm = A.rc - 1;
#pragma unroll
        for (int c = 0; c < CPC; c++)
            for (int o = threadIdx.x; o < (nq << A.logrc); o += NT) {
                const int rho = o & rcm, q = o >> A.logrc;
                const int i = (q << A.logr) + rho0 + rho;
                if (i < A.nout) store(c, i, sm[c * plane + (padded<PAD>(q) << A.logrc) + rho]);
            }
    }
    __syncthreads();
}

// shared memory carve-up, in double2 elements: [data][staged inputs][stage twiddles][g][h]
template <int LOGL>
__device__ __forceinline__ void carve(double2 *smem, const AxisDesc &A, int cpc, int plane, double2 *&y, double2 *&stw,
                                      double2 *&gt, double2 *&ht)
{
    y = smem + cpc * plane;
    stw = y + ((A.r > 1) ? (cpc << LOGL) : 0);
    gt = stw + prn_stw_total(LOGL);
    ht = gt + A.rc;
}

// Pass 1: for transform t = blockIdx.y and spectrum columns a = alo + m (m = blockIdx.x, + gridDim.x, ..):
// the y-direction transform of v1(b)*X[b][a] over the support rows, outputs j < ny, times (c/P)*u1(a).
template <int LOGL, bool PAD>
__global__ void __launch_bounds__(NT, 3) pass1_kernel(PrunedArgs a)
{
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.sd;
    const AxisDesc &A = sd.ay;
    const int t = blockIdx.y, ia = t % PFX_NA, g = t / PFX_NA;
    if ((int)blockIdx.x >= sd.ka[ia]) return;
    const int K = sd.kb[ia], lo = sd.blo[ia];
    const int plane = (padded<PAD>((1 << LOGL) - 1) + 1) << A.logrc;
    double2 *sm = smem, *y, *stw, *gt, *ht;
    carve<LOGL>(smem, A, 1, plane, y, stw, gt, ht);
    load_stage_twiddles<LOGL>(stw, A.logN, a.rooty);
    const int nsplit = A.r >> A.logrc;
    if (nsplit == 1) load_residue_factors<LOGL>(gt, ht, A, 0, a.rooty);
    const SmemTables tb{stw, gt, ht};
    const double2 *spec = g ? a.spec1 : a.spec0;
    const double *tv = a.tabv + sd.tabv + ((size_t)ia << LOGL);
    const double2 *tw = a.twb + sd.twb + ((size_t)ia << A.logN);
    const int NYl = a.NY, NYm = a.NY - 1;
    const bool staged = A.r > 1;
    for (int m = blockIdx.x; m < sd.ka[ia]; m += gridDim.x) {
        const double2 *scol = spec + (size_t)((sd.alo[ia] + m) & (a.NX - 1)) * NYl;  // spectrum is [a][b]
        auto from_global = [&](int, int mb) {
            const double2 x = __ldg(scol + ((lo + mb) & NYm));
            const double wv = tv[mb];
            return make_double2(x.x * wv, x.y * wv);
        };
        auto from_smem = [&](int, int mb) { return y[mb]; };
        if (staged)
            for (int mb = threadIdx.x; mb < K; mb += NT) y[mb] = from_global(0, mb);
        __syncthreads();
        const double wu = a.tabu[sd.tabu + (size_t)ia * sd.ax.L + m];
        size_t tjs;
        double2 *Trow = a.T + prn_t_column(a, t, m, sd.alo[ia], tjs);
        for (int part = 0; part < nsplit; part++) {
            const int rho0 = part << A.logrc;
            if (nsplit > 1) {
                load_residue_factors<LOGL>(gt, ht, A, rho0, a.rooty);
                __syncthreads();
            }
            auto store = [&](int, int j, double2 v) { Trow[(size_t)j * tjs] = make_double2(v.x * wu, v.y * wu); };
            if (staged)
                fft_lines<LOGL, 1, PAD>(sm, from_smem, store, A, plane, lo, K, rho0, tw, tb);
            else
                fft_lines<LOGL, 1, PAD>(sm, from_global, store, A, plane, lo, K, rho0, tw, tb);
        }
    }
}

// Pass 2: for transform t = blockIdx.y and output rows j0 .. j0+CPC-1 (row groups blockIdx.x, + gridDim.x,
// ..): the x-direction transform of T[t][:, j], outputs i < nx, into W[t][j][i].  The admissibility
// correction -c*E*Z is applied by the consumers of W (PlaneView::zc).
template <int LOGL, int CPC, bool PAD>
__global__ void __launch_bounds__(NT, 3) pass2_kernel(PrunedArgs a)
{
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.sd;
    const AxisDesc &A = sd.ax;
    const int t = blockIdx.y, ia = t % PFX_NA;
    const int K = sd.ka[ia], lo = sd.alo[ia];
    const int plane = (padded<PAD>((1 << LOGL) - 1) + 1) << A.logrc;
    double2 *sm = smem, *y, *stw, *gt, *ht;
    carve<LOGL>(smem, A, CPC, plane, y, stw, gt, ht);
    load_stage_twiddles<LOGL>(stw, A.logN, a.rootx);
    const int nsplit = A.r >> A.logrc;
    if (nsplit == 1) load_residue_factors<LOGL>(gt, ht, A, 0, a.rootx);
    const SmemTables tb{stw, gt, ht};
    const double2 *Tt = a.T + (size_t)t * sd.kamax * a.ny;
    const double2 *tw = a.twa + sd.twa + ((size_t)ia << A.logN);
    const int nyl = a.ny;
    const bool staged = A.r > 1;
    const int ngroups = (a.ny + CPC - 1) / CPC;
    for (int cg = blockIdx.x; cg < ngroups; cg += gridDim.x) {
        const int j0 = cg * CPC;
        auto from_global = [&](int c, int m) {
            return (j0 + c < nyl) ? __ldg(Tt + (size_t)m * nyl + j0 + c) : make_double2(0., 0.);
        };
        auto from_smem = [&](int c, int m) { return y[(c << LOGL) + m]; };
        if (staged) {
#pragma unroll
            for (int c = 0; c < CPC; c++)
                for (int m = threadIdx.x; m < K; m += NT) y[(c << LOGL) + m] = from_global(c, m);
        }
        __syncthreads();
        for (int part = 0; part < nsplit; part++) {
            const int rho0 = part << A.logrc;
            if (nsplit > 1) {
                load_residue_factors<LOGL>(gt, ht, A, rho0, a.rootx);
                __syncthreads();
            }
            double2 *Wt = a.W + ((size_t)t * a.ny + j0) * a.nx;
            const int nxl = a.nx;
            auto store = [&](int c, int i, double2 v) {
                if (CPC == 1 || j0 + c < nyl) Wt[(size_t)c * nxl + i] = v;
            };
            if (staged)
                fft_lines<LOGL, CPC, PAD>(sm, from_smem, store, A, plane, lo, K, rho0, tw, tb);
            else
                fft_lines<LOGL, CPC, PAD>(sm, from_global, store, A, plane, lo, K, rho0, tw, tb);
        }
    }
}

}  // namespace prn
}  // namespace pfx
