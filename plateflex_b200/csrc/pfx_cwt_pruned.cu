// pfx_cwt_pruned.cu -- the band-pruned inverse transform engine (default).
//
// What it replaces: the inner loop of cpwt.wlet_transform (cpwt.f90:143-188) --
// wave_function, the multiply by the spectrum, the full-size inverse fourn, the
// normalisation, cfftswap and the crop -- for all 11 azimuths of a scale and up
// to two grids, producing the cropped (nx, ny) coefficient planes directly.
//
// Idea.  The daughter (cpwt_sub.f90:90-94) is  c*[u1(kx) v1(ky) - E*u2(kx) v2(ky)]:
//   * the main Gaussian u1*v1 is centred on (kx0, ky0)/s with standard deviation 1/s in
//     k-space; beyond `cutoff` (default 9.5) sigmas it is < 2.5e-20 of its peak, so only a
//     Ka x Kb block of the spectrum contributes above rounding.  The inverse 2-D FFT of
//     a band-limited spectrum factors into  r = N/L  length-L FFTs per line (L = power of
//     two >= K): output sample i = q*r + rho is element q of the FFT of the K inputs
//     multiplied by exp(2 pi i a (rho+1)/N).  The "+1" is the reference's one-sample
//     shift (cpwt.f90:173-185), so outputs land directly on the cropped grid;
//   * the admissibility term u2*v2 depends on neither scale nor azimuth (the reference
//     uses unscaled wavenumbers there, cpwt_sub.f90:93): its transform Z is computed once
//     per grid with a dense FFT and subtracted as c*E*Z in the last pass.
// Pass 1 transforms the Ka needed spectrum columns along y (outputs j < ny only), pass 2
// transforms every output row along x (outputs i < nx only) and applies the Z correction.
// Both passes run the same shared-memory FFT: the r FFTs of a line are interleaved
// (element e of FFT rho at e*r + rho) so that consecutive lanes touch consecutive 16-byte
// words in every radix stage and the final stage leaves the line in output order.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "pfx_plan.hpp"

namespace pfx {

namespace {

constexpr int MAX_STAGES = 6;
constexpr int NT = 256;  // threads per CTA

struct AxisDesc {
    int N, logN;       // padded length of the axis
    int L, logL;       // FFT length (power of two >= support width)
    int r, logr;       // residues per line, N / L
    int rc, logrc;     // residues handled by one CTA (r / rsplit)
    int nout;          // outputs needed along this axis (nx or ny)
    int nstages;
    int lrad[MAX_STAGES];  // log2 of the radix of each stage, first executed first
    int padshift;          // element padding e + (e >> padshift) when rc < 8 (31 = none)
    int stw_off[MAX_STAGES];  // offset of each stage's twiddles in the shared-memory table
    int stw_total;
};

struct ScaleDesc {
    AxisDesc ax, ay;
    int alo[PFX_NA], ka[PFX_NA];  // support [alo, alo+ka) along x per azimuth (signed wavenumber index)
    int blo[PFX_NA], kb[PFX_NA];
    double ce[PFX_NA];            // norm * exp(-0.5*(kx0^2+ky0^2)): factor of the Z correction
    int kamax;                    // rows of T per transform
    size_t tabu, tabv;            // offsets (doubles) of the weight tables of this scale
    size_t twa, twb;              // offsets (double2) of the stage-1 twiddle tables
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// exp(+2 pi i k / 16), k = 0..7
__device__ __forceinline__ double2 w16(int k)
{
    constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, H = 0.70710678118654752440;
    switch (k) {
        case 0: return make_double2(1., 0.);
        case 1: return make_double2(C1, S1);
        case 2: return make_double2(H, H);
        case 3: return make_double2(S1, C1);
        case 4: return make_double2(0., 1.);
        case 5: return make_double2(-S1, C1);
        case 6: return make_double2(-H, H);
        default: return make_double2(-C1, S1);
    }
}

// In-register inverse DFT (exp(+i...)) of R = 2,4,8,16 points, natural order in and out.
template <int R>
__device__ __forceinline__ void dft(double2 (&v)[R])
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
            double2 t;
            const int k16 = k * (16 / R);
            if (k16 == 0) {
                t = o[k];
            } else if (k16 == 4) {
                t = make_double2(-o[k].y, o[k].x);  // * i
            } else if (k16 == 2) {
                constexpr double H = 0.70710678118654752440;
                t = make_double2(H * (o[k].x - o[k].y), H * (o[k].x + o[k].y));
            } else if (k16 == 6) {
                constexpr double H = 0.70710678118654752440;
                t = make_double2(-H * (o[k].x + o[k].y), H * (o[k].x - o[k].y));
            } else {
                t = cmul(o[k], w16(k16));
            }
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
}

__device__ __forceinline__ int padded(int e, int padshift) { return e + (e >> padshift); }

// Shared-memory tables of one CTA: stage twiddles w_{M*R}^{q0} for every stage after the first, and
// the per-residue factors of the first-stage twiddle recurrence.
struct SmemTables {
    const double2 *stw;  // stage s (>= 1) at stw + A.stw_off[s], M_s entries
    const double2 *g;    // [rc]  w_N^{(L/R1)(rho+1)}: step between the R1 inputs of a first-stage butterfly
    const double2 *h;    // [rc]  w_N^{-L(rho+1)}: correction when the support index wraps modulo L
};

__device__ __forceinline__ void load_stage_twiddles(double2 *stw, const AxisDesc &A, const double2 *__restrict__ root)
{
    int logM = A.lrad[0];
    for (int s = 1; s < A.nstages; s++) {
        const int rs = A.logN - (logM + A.lrad[s]);
        for (int q0 = threadIdx.x; q0 < (1 << logM); q0 += NT) stw[A.stw_off[s] + q0] = __ldg(root + (q0 << rs));
        logM += A.lrad[s];
    }
}

__device__ __forceinline__ void load_residue_factors(double2 *gt, double2 *ht, const AxisDesc &A, int rho0,
                                                     const double2 *__restrict__ root)
{
    for (int rho = threadIdx.x; rho < A.rc; rho += NT) {
        const int e = rho0 + rho + 1;
        gt[rho] = __ldg(root + ((e << (A.logL - A.lrad[0])) & (A.N - 1)));
        ht[rho] = __ldg(root + ((A.N - ((e << A.logL) & (A.N - 1))) & (A.N - 1)));
    }
}

// One radix-R stage (not the first) over the interleaved FFTs of all `cpc` lines held by the CTA.
//   sub-FFT size M = 2^logM, block size M*R; twiddle for input k at offset q0: (w_{M*R}^{q0})^k
template <int LR>
__device__ __forceinline__ void fft_stage(double2 *sm, const AxisDesc &A, int cpc, int plane, int logM,
                                          const double2 *stw)
{
    constexpr int R = 1 << LR;
    const int logbf = A.logL - LR;  // butterflies per FFT
    const int total = cpc << (logbf + A.logrc);
    for (int w = threadIdx.x; w < total; w += NT) {
        const int rho = w & (A.rc - 1);
        const int u = (w >> A.logrc) & ((1 << logbf) - 1);
        const int c = w >> (A.logrc + logbf);
        const int q0 = u & ((1 << logM) - 1);
        const int base = ((u >> logM) << (logM + LR)) + q0;
        double2 *p = sm + c * plane + rho;
        double2 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = p[padded(base + (k << logM), A.padshift) << A.logrc];
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
#pragma unroll
        for (int k = 0; k < R; k++) p[padded(base + (k << logM), A.padshift) << A.logrc] = v[k];
    }
}

// First stage fused with the input twiddles.  load(c, m) returns the weighted input m = a - lo of
// line c (from the staged copy in shared memory, or straight from global memory when r == 1 and every
// input is used exactly once).  The twiddle of input k of a butterfly is exp(2 pi i a_k (rho+1)/N) with
// a_k = lo + ((n_k - lo) mod L), n_k = nlow + k*L/R: one table read tw[nlow*r + rho] for k = 0, then
// t_k = t_{k-1} * g[rho], times h[rho] once when the support index wraps.
template <int LR, class Load>
__device__ __forceinline__ void fft_first_stage(double2 *sm, const Load &load, const AxisDesc &A, int cpc, int plane,
                                                int lo, int K, int rho0, const double2 *__restrict__ tw,
                                                const SmemTables &tb)
{
    constexpr int R = 1 << LR;
    const int logbf = A.logL - LR;
    const int total = 1 << (logbf + A.logrc);
    const int step = 1 << logbf;
    for (int w = threadIdx.x; w < total; w += NT) {
        const int rho = w & (A.rc - 1);
        const int u = w >> A.logrc;
        // natural index of the butterfly's first input: digits of u re-assembled in reverse order
        int nlow = 0, rest = u;
        for (int s = 1; s < A.nstages; s++) {
            const int d = rest & ((1 << A.lrad[s]) - 1);
            rest >>= A.lrad[s];
            nlow = (nlow << A.lrad[s]) | d;
        }
        const int m0 = (nlow - lo) & (A.L - 1);
        const double2 t0 = __ldg(tw + ((size_t)nlow << A.logr) + rho0 + rho);
        const double2 g = tb.g[rho], h = tb.h[rho];
        for (int c = 0; c < cpc; c++) {
            double2 v[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int m = (m0 + k * step) & (A.L - 1);
                v[k] = (m < K) ? load(c, m) : make_double2(0., 0.);
            }
            double2 t = t0;
            v[0] = cmul(v[0], t);
#pragma unroll
            for (int k = 1; k < R; k++) {
                t = cmul(t, g);
                if (((m0 + k * step) & (A.L - 1)) < step) t = cmul(t, h);  // wrapped between k-1 and k
                v[k] = cmul(v[k], t);
            }
            dft<R>(v);
            double2 *p = sm + c * plane + rho;
#pragma unroll
            for (int k = 0; k < R; k++) p[padded((u << LR) + k, A.padshift) << A.logrc] = v[k];
        }
    }
}

// MAXLR = log2 of the largest radix compiled in: 4 (radix 16, ~150 registers) or 3 (radix 8, more
// stages but three resident CTAs per SM).
template <int MAXLR, class Load>
__device__ __forceinline__ void fft_lines(double2 *sm, const Load &load, const AxisDesc &A, int cpc, int plane, int lo,
                                          int K, int rho0, const double2 *__restrict__ tw, const SmemTables &tb)
{
    const int l0 = A.lrad[0];
    if (MAXLR >= 4 && l0 == 4) fft_first_stage<(MAXLR >= 4 ? 4 : 3)>(sm, load, A, cpc, plane, lo, K, rho0, tw, tb);
    else if (l0 == 3) fft_first_stage<3>(sm, load, A, cpc, plane, lo, K, rho0, tw, tb);
    else if (l0 == 2) fft_first_stage<2>(sm, load, A, cpc, plane, lo, K, rho0, tw, tb);
    else fft_first_stage<1>(sm, load, A, cpc, plane, lo, K, rho0, tw, tb);
    int logM = l0;
    for (int s = 1; s < A.nstages; s++) {
        __syncthreads();
        const int l = A.lrad[s];
        const double2 *stw = tb.stw + A.stw_off[s];
        if (MAXLR >= 4 && l == 4) fft_stage<(MAXLR >= 4 ? 4 : 3)>(sm, A, cpc, plane, logM, stw);
        else if (l == 3) fft_stage<3>(sm, A, cpc, plane, logM, stw);
        else if (l == 2) fft_stage<2>(sm, A, cpc, plane, logM, stw);
        else fft_stage<1>(sm, A, cpc, plane, logM, stw);
        logM += l;
    }
    __syncthreads();
}

struct PrunedArgs {
    ScaleDesc sd;
    int nx, ny, NX, NY, ngrids;
    const double2 *spec0, *spec1;  // padded spectra [b][a]
    const double *tabu, *tabv;     // weight tables (all scales)
    const double2 *twa, *twb;      // stage-1 twiddle tables (all scales)
    const double2 *rootx, *rooty;  // exp(2 pi i m / N)
    double2 *T;                    // [t][ka row][j]  intermediate, t = ia + 11*g
    double2 *W;                    // [t][j][i] output planes (main Gaussian term only)
};

// shared memory carve-up, in double2 elements: [data][staged inputs][stage twiddles][g][h]
__device__ __forceinline__ void carve(double2 *smem, const AxisDesc &A, int cpc, int plane, double2 *&y, double2 *&stw,
                                      double2 *&gt, double2 *&ht)
{
    y = smem + cpc * plane;
    stw = y + ((A.r > 1) ? cpc * A.L : 0);
    gt = stw + A.stw_total;
    ht = gt + A.rc;
}

// Pass 1: for transform t = blockIdx.y and spectrum columns a = alo + m (m = blockIdx.x, + gridDim.x, ..):
// the y-direction transform of v1(b)*X[b][a] over the support rows, outputs j < ny, times (c/P)*u1(a).
template <int MAXLR>
__global__ void __launch_bounds__(NT, (MAXLR >= 4 ? 1 : 3)) pruned_pass1_kernel(PrunedArgs a)
{
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.sd;
    const AxisDesc &A = sd.ay;
    const int t = blockIdx.y, ia = t % PFX_NA, g = t / PFX_NA;
    if ((int)blockIdx.x >= sd.ka[ia]) return;
    const int K = sd.kb[ia], lo = sd.blo[ia];
    const int plane = (padded(A.L - 1, A.padshift) + 1) << A.logrc;
    double2 *sm = smem, *y, *stw, *gt, *ht;
    carve(smem, A, 1, plane, y, stw, gt, ht);
    load_stage_twiddles(stw, A, a.rooty);
    const int nsplit = A.r >> A.logrc;
    if (nsplit == 1) load_residue_factors(gt, ht, A, 0, a.rooty);
    const SmemTables tb{stw, gt, ht};
    const double2 *spec = g ? a.spec1 : a.spec0;
    const double *tv = a.tabv + sd.tabv + (size_t)ia * A.L;
    const double2 *tw = a.twb + sd.twb + ((size_t)ia << A.logN);
    const int NXl = a.NX, NYm = a.NY - 1;
    const bool staged = A.r > 1;
    for (int m = blockIdx.x; m < sd.ka[ia]; m += gridDim.x) {
        const int acol = (sd.alo[ia] + m) & (a.NX - 1);
        auto from_global = [&](int, int mb) {
            const double2 x = __ldg(spec + (size_t)((lo + mb) & NYm) * NXl + acol);
            const double wv = tv[mb];
            return make_double2(x.x * wv, x.y * wv);
        };
        auto from_smem = [&](int, int mb) { return y[mb]; };
        if (staged)
            for (int mb = threadIdx.x; mb < K; mb += NT) y[mb] = from_global(0, mb);
        __syncthreads();
        const double wu = a.tabu[sd.tabu + (size_t)ia * sd.ax.L + m];
        double2 *Trow = a.T + ((size_t)t * sd.kamax + m) * a.ny;
        for (int part = 0; part < nsplit; part++) {
            const int rho0 = part << A.logrc;
            if (nsplit > 1) {
                load_residue_factors(gt, ht, A, rho0, a.rooty);
                __syncthreads();
            }
            if (staged)
                fft_lines<MAXLR>(sm, from_smem, A, 1, plane, lo, K, rho0, tw, tb);
            else
                fft_lines<MAXLR>(sm, from_global, A, 1, plane, lo, K, rho0, tw, tb);
            // outputs j = q*r + rho0 + rho, j < ny
            const int nq = (A.nout + A.r - 1) >> A.logr;
            for (int o = threadIdx.x; o < (nq << A.logrc); o += NT) {
                const int rho = o & (A.rc - 1), q = o >> A.logrc;
                const int j = (q << A.logr) + rho0 + rho;
                if (j < A.nout) {
                    const double2 v = sm[(padded(q, A.padshift) << A.logrc) + rho];
                    Trow[j] = make_double2(v.x * wu, v.y * wu);
                }
            }
            __syncthreads();
        }
    }
}

// Pass 2: for transform t = blockIdx.y and output rows j0 .. j0+CPC-1 (row groups blockIdx.x, + gridDim.x,
// ..): the x-direction transform of T[t][:, j], outputs i < nx, into W[t][j][i].  The admissibility
// correction -c*E*Z is applied by the consumers of W (PlaneView::zc).
template <int CPC, int MAXLR>
__global__ void __launch_bounds__(NT, (MAXLR >= 4 ? 1 : 3)) pruned_pass2_kernel(PrunedArgs a)
{
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.sd;
    const AxisDesc &A = sd.ax;
    const int t = blockIdx.y, ia = t % PFX_NA;
    const int K = sd.ka[ia], lo = sd.alo[ia];
    const int plane = (padded(A.L - 1, A.padshift) + 1) << A.logrc;
    double2 *sm = smem, *y, *stw, *gt, *ht;
    carve(smem, A, CPC, plane, y, stw, gt, ht);
    load_stage_twiddles(stw, A, a.rootx);
    const int nsplit = A.r >> A.logrc;
    if (nsplit == 1) load_residue_factors(gt, ht, A, 0, a.rootx);
    const SmemTables tb{stw, gt, ht};
    const double2 *Tt = a.T + (size_t)t * sd.kamax * a.ny;
    const double2 *tw = a.twa + sd.twa + ((size_t)ia << A.logN);
    const int nyl = a.ny, Ll = A.L;
    const bool staged = A.r > 1;
    const int ngroups = (a.ny + CPC - 1) / CPC;
    for (int cg = blockIdx.x; cg < ngroups; cg += gridDim.x) {
        const int j0 = cg * CPC;
        auto from_global = [&](int c, int m) {
            const int j = j0 + c;
            return (j < nyl) ? __ldg(Tt + (size_t)m * nyl + j) : make_double2(0., 0.);
        };
        auto from_smem = [&](int c, int m) { return y[c * Ll + m]; };
        if (staged)
            for (int o = threadIdx.x; o < K * CPC; o += NT) {
                const int c = o % CPC, m = o / CPC;
                y[c * Ll + m] = from_global(c, m);
            }
        __syncthreads();
        for (int part = 0; part < nsplit; part++) {
            const int rho0 = part << A.logrc;
            if (nsplit > 1) {
                load_residue_factors(gt, ht, A, rho0, a.rootx);
                __syncthreads();
            }
            if (staged)
                fft_lines<MAXLR>(sm, from_smem, A, CPC, plane, lo, K, rho0, tw, tb);
            else
                fft_lines<MAXLR>(sm, from_global, A, CPC, plane, lo, K, rho0, tw, tb);
            const int nq = (A.nout + A.r - 1) >> A.logr;
            const int per = nq << A.logrc;
            for (int o = threadIdx.x; o < per * CPC; o += NT) {
                const int c = o / per, oo = o - c * per;
                const int rho = oo & (A.rc - 1), q = oo >> A.logrc;
                const int i = (q << A.logr) + rho0 + rho;
                const int j = j0 + c;
                if (i < A.nout && j < a.ny)
                    a.W[((size_t)t * a.ny + j) * a.nx + i] = sm[c * plane + (padded(q, A.padshift) << A.logrc) + rho];
            }
            __syncthreads();
        }
    }
}

// ---- table builders ------------------------------------------------------------------------
__global__ void root_kernel(int N, double2 *root)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)N, &s, &c);
    root[m] = make_double2(c, s);
}

// weights: tab[ia][m] = scale * exp(-0.5*(s*k(lo+m) - k0c[ia])^2) for m < K, else 0  (cpwt_sub.f90:90-92)
__global__ void weight_kernel(int L, int N, double dk, double s, double scale, const int *lo, const int *K,
                              const double *k0c, double *tab)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = blockIdx.y;
    if (m >= L) return;
    double v = 0.0;
    if (m < K[ia]) {
        const int ai = lo[ia] + m;  // signed index in (-N/2, N/2]
        const double k = (ai >= 0) ? (double)ai * dk : -((double)(-ai) * dk);  // defk, cpwt_sub.f90:47-60
        const double d = s * k - k0c[ia];
        v = scale * exp(-0.5 * (d * d));
    }
    tab[(size_t)ia * L + m] = v;
}

// tw[ia][n0][rho] = root[(a(n0) * (rho+1)) mod N],  a(n0) = lo + ((n0 - lo) mod L)
__global__ void twiddle_kernel(int logN, int logL, const int *lo, const double2 *__restrict__ root, double2 *tw)
{
    const int N = 1 << logN, L = 1 << logL, logr = logN - logL;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = blockIdx.y;
    if (e >= N) return;
    const int rho = e & ((1 << logr) - 1), n0 = e >> logr;
    const int ai = lo[ia] + ((n0 - lo[ia]) & (L - 1));
    const long long idx = ((long long)ai * (rho + 1)) & (N - 1);
    tw[((size_t)ia << logN) + e] = root[idx];
}

// Admissibility term: plane[g][b][a] = exp(-kx^2/2) exp(-ky^2/2) X_g[b][a]   (cpwt_sub.f90:93)
__global__ void zmul_kernel(int NX, int NY, const double2 *__restrict__ s0, const double2 *__restrict__ s1,
                            const double *__restrict__ u2, const double *__restrict__ v2, double2 *planes)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (a >= NX) return;
    const size_t o = (size_t)b * NX + a, P = (size_t)NX * NY;
    const double w = u2[a] * v2[b];
    const double2 x0 = s0[o];
    planes[o] = make_double2(w * x0.x, w * x0.y);
    if (s1) {
        const double2 x1 = s1[o];
        planes[P + o] = make_double2(w * x1.x, w * x1.y);
    }
}

__global__ void zcrop_kernel(int nx, int ny, int NX, size_t P, double scale, const double2 *__restrict__ planes,
                             double2 *zc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, g = blockIdx.z;
    if (i >= nx) return;
    const double2 v = planes[g * P + (size_t)(j + 1) * NX + (i + 1)];
    zc[((size_t)g * ny + j) * nx + i] = make_double2(v.x * scale, v.y * scale);
}

size_t smem_bytes(const AxisDesc &A, int cpc)
{
    const size_t plane = (size_t)((A.L - 1) + ((A.L - 1) >> A.padshift) + 1) * A.rc;
    return (plane * cpc + (A.r > 1 ? (size_t)A.L * cpc : 0) + A.stw_total + 2 * (size_t)A.rc) * sizeof(double2);
}

int ilog2(int v)
{
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

void plan_axis(AxisDesc &A, int N, int K, int nout, size_t smem_budget_elems, int cpc, int maxlr)
{
    A.N = N;
    A.logN = ilog2(N);
    int L = 2;
    while (L < K) L *= 2;
    if (L > N) L = N;
    A.L = L;
    A.logL = ilog2(L);
    A.r = N / L;
    A.logr = A.logN - A.logL;
    A.nout = nout;
    // radix plan: as many radix-16 stages as possible, remainder 8 / 4 / 2, largest radix first
    int rem = A.logL, n = 0;
    int tmp[MAX_STAGES];
    while (rem > 0) {
        int lr = rem >= maxlr ? maxlr : rem;
        if (maxlr == 4) {
            if (rem == 5) lr = 3;    // 32 = 8 * 4
            if (rem == 6) lr = 3;    // 64 = 8 * 8
            if (rem == 9) lr = 3;    // 512 = 8 * 8 * 8
        } else if (rem == 4) {
            lr = 2;                  // 16 = 4 * 4
        }
        tmp[n++] = lr;
        rem -= lr;
    }
    A.nstages = n;
    for (int s = 0; s < n; s++) A.lrad[s] = tmp[s];
    for (int s = n; s < MAX_STAGES; s++) A.lrad[s] = 0;
    // residues per CTA: all of them unless the line does not fit in shared memory
    A.rc = A.r;
    A.logrc = A.logr;
    auto plane_elems = [&](int rc, int padshift) { return (size_t)((L - 1) + ((L - 1) >> padshift) + 1) * rc; };
    A.padshift = (A.rc >= 8) ? 31 : 3;
    {
        int logM = A.lrad[0], off = 0;
        A.stw_off[0] = 0;
        for (int st = 1; st < MAX_STAGES; st++) {
            A.stw_off[st] = off;
            if (st < A.nstages) {
                off += 1 << logM;
                logM += A.lrad[st];
            }
        }
        A.stw_total = off;
    }
    const size_t ystage = ((A.r > 1) ? (size_t)L * cpc : 0) + A.stw_total + 2 * (size_t)A.r;
    while (A.rc > 1 && plane_elems(A.rc, A.padshift) * cpc + ystage > smem_budget_elems) {
        A.rc /= 2;
        A.logrc--;
        A.padshift = (A.rc >= 8) ? 31 : 3;
    }
}

}  // namespace

struct PrunedState {
    std::vector<ScaleDesc> sd;
    double2 *rootx = nullptr, *rooty = nullptr;
    double *tabu = nullptr, *tabv = nullptr;
    double2 *twa = nullptr, *twb = nullptr;
    double2 *T = nullptr, *W = nullptr, *zc = nullptr, *zplanes = nullptr;
    cufftHandle fft_z[2] = {0, 0};
    bool fft_z_ok[2] = {false, false};
    size_t smem1 = 0, smem2 = 0;
    int cpc = 2;
    int maxlr = 3;
    int ctas_per_transform = 64;
    double cur_ce[PFX_NA] = {};  // c*E of the scale whose planes currently sit in W
    std::vector<std::pair<void *, size_t>> owned;
};

static PrunedState *state_of(pfx_plan *pl) { return reinterpret_cast<PrunedState *>(pl->pruned); }

void plan_free_pruned(pfx_plan *pl);

int plan_init_pruned(pfx_plan *pl)
{
    // (re)build everything that depends on the cutoff
    plan_free_pruned(pl);  // a cutoff change rebuilds tables and buffers
    PrunedState *st = new PrunedState();
    pl->pruned = st;
    auto dalloc = [&](void **p, size_t bytes) -> int {
        *p = nullptr;
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("pruned engine: device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return PFX_ERR_OOM;
        }
        st->owned.emplace_back(*p, bytes);
        pl->device_bytes += bytes;
        return PFX_OK;
    };
    const int NX = pl->NX, NY = pl->NY, nx = pl->nx, ny = pl->ny, ns = pl->ns;
    const double pi = 3.141592653589793;
    const double dkx = 2 * pi / ((double)NX * pl->dx), dky = 2 * pi / ((double)NY * pl->dy);
    const size_t P = (size_t)NX * NY;
    int dev_smem = 0;
    PFX_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, pl->device));
    const size_t budget = (size_t)dev_smem / sizeof(double2);

    st->cpc = (NX <= 4096) ? 2 : 1;
    {
        const char *e = getenv("PLATEFLEX_B200_RADIX");
        st->maxlr = (e && atoi(e) == 16) ? 4 : 3;
        const char *c = getenv("PLATEFLEX_B200_CTAS");
        if (c && atoi(c) > 0) st->ctas_per_transform = atoi(c);
    }
    st->sd.resize(ns);
    size_t tabu_n = 0, tabv_n = 0, twa_n = 0, twb_n = 0, kamax_all = 0;
    std::vector<int> h_lo((size_t)ns * 4 * PFX_NA);
    for (int is = 0; is < ns; is++) {
        ScaleDesc &sd = st->sd[is];
        const PfxScaleConsts &c = pl->consts[is];
        int kamax = 0, kbmax = 0;
        for (int ia = 0; ia < PFX_NA; ia++) {
            auto range = [&](double k0c, double dk, int N, int &lo, int &K) {
                int l = (int)std::ceil((k0c - pl->cutoff) / (c.s * dk));
                int h = (int)std::floor((k0c + pl->cutoff) / (c.s * dk));
                l = std::max(l, -N / 2 + 1);
                h = std::min(h, N / 2);
                lo = l;
                K = std::max(h - l + 1, 0);
            };
            range(c.kx0[ia], dkx, NX, sd.alo[ia], sd.ka[ia]);
            range(c.ky0[ia], dky, NY, sd.blo[ia], sd.kb[ia]);
            kamax = std::max(kamax, sd.ka[ia]);
            kbmax = std::max(kbmax, sd.kb[ia]);
            sd.ce[ia] = c.norm * c.e0[ia];
        }
        plan_axis(sd.ax, NX, std::max(kamax, 1), nx, budget, st->cpc, st->maxlr);
        plan_axis(sd.ay, NY, std::max(kbmax, 1), ny, budget, 1, st->maxlr);
        sd.kamax = std::max(kamax, 1);
        sd.tabu = tabu_n;
        sd.tabv = tabv_n;
        sd.twa = twa_n;
        sd.twb = twb_n;
        tabu_n += (size_t)PFX_NA * sd.ax.L;
        tabv_n += (size_t)PFX_NA * sd.ay.L;
        twa_n += (size_t)PFX_NA * NX;
        twb_n += (size_t)PFX_NA * NY;
        kamax_all = std::max(kamax_all, (size_t)sd.kamax);
        st->smem1 = std::max(st->smem1, smem_bytes(sd.ay, 1));
        st->smem2 = std::max(st->smem2, smem_bytes(sd.ax, st->cpc));
    }
    if (st->smem1 > (size_t)dev_smem || st->smem2 > (size_t)dev_smem) {
        set_error("pruned engine: a %d x %d padded line does not fit in shared memory; use the dense engine", NX, NY);
        return PFX_ERR_ARG;
    }

    PFX_CHECK(dalloc((void **)&st->rootx, NX * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->rooty, NY * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->tabu, tabu_n * sizeof(double)));
    PFX_CHECK(dalloc((void **)&st->tabv, tabv_n * sizeof(double)));
    PFX_CHECK(dalloc((void **)&st->twa, twa_n * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->twb, twb_n * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->T, 2 * PFX_NA * kamax_all * ny * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->W, 2 * PFX_NA * (size_t)nx * ny * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->zc, 2 * (size_t)nx * ny * sizeof(double2)));
    PFX_CHECK(dalloc((void **)&st->zplanes, 2 * P * sizeof(double2)));

    root_kernel<<<(NX + 255) / 256, 256, 0, pl->stream>>>(NX, st->rootx);
    root_kernel<<<(NY + 255) / 256, 256, 0, pl->stream>>>(NY, st->rooty);

    // per-scale tables
    int *d_i = nullptr;
    double *d_k0 = nullptr;
    PFX_CHECK(dalloc((void **)&d_i, (size_t)ns * 4 * PFX_NA * sizeof(int)));
    PFX_CHECK(dalloc((void **)&d_k0, (size_t)ns * 2 * PFX_NA * sizeof(double)));
    std::vector<double> h_k0((size_t)ns * 2 * PFX_NA);
    for (int is = 0; is < ns; is++) {
        const ScaleDesc &sd = st->sd[is];
        for (int ia = 0; ia < PFX_NA; ia++) {
            h_lo[((size_t)is * 4 + 0) * PFX_NA + ia] = sd.alo[ia];
            h_lo[((size_t)is * 4 + 1) * PFX_NA + ia] = sd.ka[ia];
            h_lo[((size_t)is * 4 + 2) * PFX_NA + ia] = sd.blo[ia];
            h_lo[((size_t)is * 4 + 3) * PFX_NA + ia] = sd.kb[ia];
            h_k0[((size_t)is * 2 + 0) * PFX_NA + ia] = pl->consts[is].kx0[ia];
            h_k0[((size_t)is * 2 + 1) * PFX_NA + ia] = pl->consts[is].ky0[ia];
        }
    }
    PFX_CUDA(cudaMemcpyAsync(d_i, h_lo.data(), h_lo.size() * sizeof(int), cudaMemcpyHostToDevice, pl->stream));
    PFX_CUDA(cudaMemcpyAsync(d_k0, h_k0.data(), h_k0.size() * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    for (int is = 0; is < ns; is++) {
        const ScaleDesc &sd = st->sd[is];
        const PfxScaleConsts &c = pl->consts[is];
        const int *alo = d_i + ((size_t)is * 4 + 0) * PFX_NA, *ka = alo + PFX_NA, *blo = ka + PFX_NA, *kb = blo + PFX_NA;
        const double *kx0 = d_k0 + ((size_t)is * 2) * PFX_NA, *ky0 = kx0 + PFX_NA;
        // (c/P) folded into the x weights: norm of cpwt_sub.f90:87 and the 1/(nnx*nny) of cpwt.f90:177
        weight_kernel<<<dim3((sd.ax.L + 127) / 128, PFX_NA), 128, 0, pl->stream>>>(sd.ax.L, NX, dkx, c.s,
                                                                                    c.norm / (double)P, alo, ka, kx0,
                                                                                    st->tabu + sd.tabu);
        weight_kernel<<<dim3((sd.ay.L + 127) / 128, PFX_NA), 128, 0, pl->stream>>>(sd.ay.L, NY, dky, c.s, 1.0, blo, kb,
                                                                                    ky0, st->tabv + sd.tabv);
        twiddle_kernel<<<dim3((NX + 255) / 256, PFX_NA), 256, 0, pl->stream>>>(sd.ax.logN, sd.ax.logL, alo, st->rootx,
                                                                                st->twa + sd.twa);
        twiddle_kernel<<<dim3((NY + 255) / 256, PFX_NA), 256, 0, pl->stream>>>(sd.ay.logN, sd.ay.logL, blo, st->rooty,
                                                                                st->twb + sd.twb);
    }
    PFX_CUDA(cudaGetLastError());
    PFX_CUDA(cudaStreamSynchronize(pl->stream));  // h_lo / h_k0 go out of scope

    // dense inverse plans for the admissibility term (1 or 2 planes)
    int n[2] = {NY, NX};
    size_t need = pl->fft_work_bytes;
    for (int g = 0; g < 2; g++) {
        size_t ws = 0;
        PFX_CUFFT(cufftCreate(&st->fft_z[g]));
        st->fft_z_ok[g] = true;
        PFX_CUFFT(cufftSetAutoAllocation(st->fft_z[g], 0));
        PFX_CUFFT(cufftMakePlanMany(st->fft_z[g], 2, n, nullptr, 1, (int)P, nullptr, 1, (int)P, CUFFT_Z2Z, g + 1, &ws));
        need = std::max(need, ws);
    }
    if (need > pl->fft_work_bytes) {
        void *w = nullptr;
        PFX_CHECK(pl->alloc(&w, need));
        pl->d_fft_work = w;
        pl->fft_work_bytes = need;
        PFX_CUFFT(cufftSetWorkArea(pl->fft_fwd, pl->d_fft_work));
    }
    for (int g = 0; g < 2; g++) PFX_CUFFT(cufftSetWorkArea(st->fft_z[g], pl->d_fft_work));

    PFX_CUDA(cudaFuncSetAttribute(pruned_pass1_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem1));
    PFX_CUDA(cudaFuncSetAttribute(pruned_pass1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem1));
    PFX_CUDA(cudaFuncSetAttribute(pruned_pass2_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem2));
    PFX_CUDA(cudaFuncSetAttribute(pruned_pass2_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem2));
    PFX_CUDA(cudaFuncSetAttribute(pruned_pass2_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem2));
    PFX_CUDA(cudaFuncSetAttribute(pruned_pass2_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem2));
    return PFX_OK;
}

void plan_free_pruned(pfx_plan *pl)
{
    PrunedState *st = state_of(pl);
    if (!st) return;
    for (int g = 0; g < 2; g++)
        if (st->fft_z_ok[g]) cufftDestroy(st->fft_z[g]);
    cudaStreamSynchronize(pl->stream);
    for (auto &pb : st->owned) {
        cudaFree(pb.first);
        pl->device_bytes -= pb.second;
    }
    delete st;
    pl->pruned = nullptr;
}

// Once per set of forward spectra: Z = crop(IFFT2(u2*v2*X))/P for each grid.
int pruned_prepare(pfx_plan *pl, int ngrids)
{
    PrunedState *st = state_of(pl);
    const size_t P = (size_t)pl->NX * pl->NY;
    zmul_kernel<<<dim3((pl->NX + 255) / 256, pl->NY), 256, 0, pl->stream>>>(
        pl->NX, pl->NY, pl->d_spec[0], ngrids == 2 ? pl->d_spec[1] : nullptr, pl->d_u2, pl->d_v2, st->zplanes);
    PFX_CUDA(cudaGetLastError());
    cufftHandle h = st->fft_z[ngrids - 1];
    PFX_CUFFT(cufftSetStream(h, pl->stream));
    PFX_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex *)st->zplanes, (cufftDoubleComplex *)st->zplanes, CUFFT_INVERSE));
    zcrop_kernel<<<dim3((pl->nx + 127) / 128, pl->ny, ngrids), 128, 0, pl->stream>>>(pl->nx, pl->ny, pl->NX, P,
                                                                                     1.0 / (double)P, st->zplanes, st->zc);
    PFX_CUDA(cudaGetLastError());
    pl->launches += 3;
    return PFX_OK;
}

int pruned_scale_planes(pfx_plan *pl, int is, int ngrids)
{
    PrunedState *st = state_of(pl);
    if (!st) {
        set_error("pruned engine not initialised");
        return PFX_ERR_STATE;
    }
    PrunedArgs a;
    a.sd = st->sd[is];
    a.nx = pl->nx;
    a.ny = pl->ny;
    a.NX = pl->NX;
    a.NY = pl->NY;
    a.ngrids = ngrids;
    a.spec0 = pl->d_spec[0];
    a.spec1 = pl->d_spec[1];
    a.tabu = st->tabu;
    a.tabv = st->tabv;
    a.twa = st->twa;
    a.twb = st->twb;
    a.rootx = st->rootx;
    a.rooty = st->rooty;
    a.T = st->T;
    a.W = st->W;
    const ScaleDesc &sd = a.sd;
    const size_t sm1 = smem_bytes(sd.ay, 1), sm2 = smem_bytes(sd.ax, st->cpc);
    const int ngroups = (pl->ny + st->cpc - 1) / st->cpc;
    dim3 g1(std::min(sd.kamax, st->ctas_per_transform), PFX_NA * ngrids);
    dim3 g2(std::min(ngroups, st->ctas_per_transform), PFX_NA * ngrids);
    if (st->maxlr >= 4) {
        pruned_pass1_kernel<4><<<g1, NT, sm1, pl->stream>>>(a);
        if (st->cpc == 2)
            pruned_pass2_kernel<2, 4><<<g2, NT, sm2, pl->stream>>>(a);
        else
            pruned_pass2_kernel<1, 4><<<g2, NT, sm2, pl->stream>>>(a);
    } else {
        pruned_pass1_kernel<3><<<g1, NT, sm1, pl->stream>>>(a);
        if (st->cpc == 2)
            pruned_pass2_kernel<2, 3><<<g2, NT, sm2, pl->stream>>>(a);
        else
            pruned_pass2_kernel<1, 3><<<g2, NT, sm2, pl->stream>>>(a);
    }
    PFX_CUDA(cudaGetLastError());
    for (int ia = 0; ia < PFX_NA; ia++) st->cur_ce[ia] = sd.ce[ia];
    pl->launches += 2;
    return PFX_OK;
}

PlaneView pruned_view(const pfx_plan *pl, int g)
{
    const PrunedState *st = reinterpret_cast<const PrunedState *>(pl->pruned);
    const size_t nxy = (size_t)pl->nx * pl->ny;
    PlaneView v;
    v.base = st->W + (size_t)g * PFX_NA * nxy;
    v.plane_stride = nxy;
    v.ld = pl->nx;
    v.off_i = 0;
    v.off_j = 0;
    v.scale = 1.0;
    v.zc = st->zc + (size_t)g * nxy;  // W = main term - c*E*Z (cpwt_sub.f90:93-94)
    for (int ia = 0; ia < PFX_NA; ia++) v.ce[ia] = st->cur_ce[ia];
    return v;
}

}  // namespace pfx
