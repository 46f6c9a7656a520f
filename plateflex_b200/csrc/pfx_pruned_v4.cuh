// pfx_pruned_v4.cuh -- fourth-generation pass kernels of the band-pruned engine: radix-16 butterflies.
//
// Why.  ncu of the second-generation pass 2 at 4096^2 (profiles/r02_ncu_full.txt): the shared-memory data pipe
// is the busiest unit (l1tex data-pipe wavefronts 73 % of peak at L = 512, 85 % at L = 1024) while the FP64 pipe
// sits at 39-47 %.  A radix-8 plan moves every element of a line through shared memory once per stage boundary
// (16 B written + 16 B read), and L = 128 / 256 need three stages, L = 1024 / 2048 four.  With 16 elements per
// thread the plan is  16 x {2,4,8,16}  for L <= 256 (ONE round trip) and  16 x 8 x {4,8,16}  for L <= 2048 (two):
// a quarter to a third fewer shared-memory wavefronts for the same floating-point work
// (3.4 FP64 instructions per element and radix bit either way).
//
// Mapping (same mathematics as pfx_pruned_v2.cuh, see pfx_cwt_pruned.cu):
//   * 128 threads per CTA, a line part is CH = 2048 elements = the L-point FFTs of RC = 2048/L consecutive
//     residues; every thread owns ONE radix-16 first-stage butterfly (u, rho) for the whole kernel, its sixteen
//     support samples stay in registers across the parts of a line (N > 2048), the first-stage twiddles
//     w_N^(a (rho+1)) advance from part to part by one complex multiply;
//   * element e of residue rho at shared-memory slot ((e + (e >> 4)) << LOGRC) + rho when RC < 8 (one slot of
//     padding per 16 elements keeps both the radix-16 first-stage stores, 17 slots apart between lanes, and the
//     8-aligned runs read by the later stages on distinct banks), plain (e << LOGRC) + rho otherwise;
//   * later stages: twiddled radix-R butterflies built from log2(R) levels of 6-FMA radix-2 butterflies
//     (tw_dft), the last one storing straight to global memory; on power-of-two grids (N = 2 nout) the upper
//     half of every FFT is never computed.
#pragma once
#include "pfx_pruned.hpp"
#include "pfx_pruned_kernels.cuh"

namespace pfx {
namespace v4 {

using prn::cadd;
using prn::cmul;
using prn::csub;

constexpr int NT = 128;
constexpr int LOGCH = 11;  // 16 elements per thread
#ifndef PFX_V4_SPLIT_CHAIN
#define PFX_V4_SPLIT_CHAIN 0
#endif
#define PFX_V4_HD __host__ __device__
#ifndef PFX_V4_MULTI_OCC
#define PFX_V4_MULTI_OCC 3  // CTAs per SM the several-parts-per-line kernels are compiled for (168 registers)
#endif

// radix plan: first stage 16, then one stage (L <= 256) or radix 8 followed by the rest (L <= 2048)
__host__ __device__ constexpr int nstages(int logl) { return logl - 4 <= 4 ? 2 : 3; }
__host__ __device__ constexpr int lrad(int logl, int s)
{
    const int rem = logl - 4;
    if (s == 0) return 4;
    if (rem <= 4) return s == 1 ? rem : 0;
    return s == 1 ? 3 : (s == 2 ? rem - 3 : 0);
}
__host__ __device__ constexpr int logm(int logl, int s)
{
    int m = 0;
    for (int i = 0; i < s; i++) m += lrad(logl, i);
    return m;
}
__host__ __device__ constexpr int stw_off(int logl, int s)
{
    int off = 0;
    for (int i = 1; i < s; i++) off += 1 << logm(logl, i);
    return off;
}
__host__ __device__ constexpr int stw_total(int logl) { return stw_off(logl, nstages(logl)); }

template <int LOGL>
struct Geo {
    static_assert(LOGL >= 5 && LOGL <= LOGCH, "v4 kernels: 32 <= L <= 2048");
    static constexpr int LOGRC = LOGCH - LOGL;
    static constexpr int RC = 1 << LOGRC;
    static constexpr int L = 1 << LOGL;
    static constexpr bool PAD = LOGRC < 3;
    static constexpr int NST = nstages(LOGL);
    static constexpr int PLANE = (PAD ? (L + (L >> 4)) : L) << LOGRC;  // double2 elements per line part
    static constexpr int LOGBF = LOGL - 4;   // log2 of first-stage butterflies per FFT
    static constexpr int STEP = 1 << LOGBF;  // distance between the inputs of a first-stage butterfly
    static constexpr int STW = stw_total(LOGL);
};

template <bool PAD>
__host__ __device__ __forceinline__ constexpr int padded(int e)
{
    return PAD ? e + (e >> 4) : e;
}

struct V4Args {
    PrunedArgs p;
    int logr;    // log2(N / L) of the transformed axis
    int nparts;  // N / CH
};

template <int LOGL>
__device__ __forceinline__ void load_stage_twiddles(double2 *stw, int logN, const double2 *__restrict__ root)
{
#pragma unroll
    for (int s = 1; s < nstages(LOGL); s++) {
        const int lm = logm(LOGL, s);
        const int rs = logN - (lm + lrad(LOGL, s));
        for (int q0 = threadIdx.x; q0 < (1 << lm); q0 += NT) stw[stw_off(LOGL, s) + q0] = __ldg(root + (q0 << rs));
    }
}

template <int LOGL>
__host__ __device__ __forceinline__ int digit_reverse(int u)
{
    int nlow = 0;
#pragma unroll
    for (int s = 1; s < nstages(LOGL); s++) {
        const int lr = lrad(LOGL, s);
        nlow = (nlow << lr) | (u & ((1 << lr) - 1));
        u >>= lr;
    }
    return nlow;
}

// a <- a + w b, b <- a - w b in six fused multiply-adds (the difference as 2a - (a + w b)).
__host__ __device__ __forceinline__ void bfly(double2 &a, double2 &b, const double2 w)
{
    const double pr = fma(w.x, b.x, fma(-w.y, b.y, a.x));
    const double pi = fma(w.x, b.y, fma(w.y, b.x, a.y));
    b = make_double2(fma(2.0, a.x, -pr), fma(2.0, a.y, -pi));
    a = make_double2(pr, pi);
}
__host__ __device__ __forceinline__ void bfly_plus(double2 &a, const double2 b, const double2 w)
{
    a = make_double2(fma(w.x, b.x, fma(-w.y, b.y, a.x)), fma(w.x, b.y, fma(w.y, b.x, a.y)));
}
__host__ __device__ __forceinline__ double2 csqr(const double2 w) { return make_double2(fma(w.x, w.x, -(w.y * w.y)), 2.0 * (w.x * w.y)); }

// z * exp(2 pi i / Q)
template <int Q>
__host__ __device__ __forceinline__ double2 mul_root(const double2 z)
{
    if constexpr (Q == 2) return make_double2(-z.x, -z.y);
    else if constexpr (Q == 4) return make_double2(-z.y, z.x);
    else if constexpr (Q == 8) {
        constexpr double H = 0.70710678118654752440;
        return make_double2(H * (z.x - z.y), H * (z.x + z.y));
    } else {
        static_assert(Q == 16, "roots up to the sixteenth");
        constexpr double C = 0.92387953251128675613, S = 0.38268343236508977173;
        return make_double2(fma(C, z.x, -(S * z.y)), fma(C, z.y, S * z.x));
    }
}

__host__ __device__ constexpr int ilog2c(int r)
{
    int l = 0;
    while (r > 1) {
        r >>= 1;
        l++;
    }
    return l;
}

// Twiddled radix-R butterfly, X[k] = sum_j W_R^{jk} w^j v[j] (W_R = exp(+2 pi i / R)), R = 2 .. 16, as log2(R)
// levels of radix-2 butterflies.  pw[i] = w^(2^i).  Splitting j = j' + (R/2) b:  the even outputs are the
// R/2-point problem of v[j'] + w^(R/2) v[j'+R/2] with twiddle w, the odd ones that of the differences with
// twiddle w W_R, whose powers are pw[i] times fixed roots.  Output k is left in v[bitrev(k)]; with HALF only
// X[0 .. R/2) are computed (only the final level differs).
template <int R, bool HALF>
__host__ __device__ __forceinline__ void tw_dft(double2 *v, const double2 *pw)
{
    constexpr int LG = ilog2c(R);
    if constexpr (R == 2) {
        if constexpr (HALF) bfly_plus(v[0], v[1], pw[0]);
        else bfly(v[0], v[1], pw[0]);
    } else {
#pragma unroll
        for (int j = 0; j < R / 2; j++) bfly(v[j], v[j + R / 2], pw[LG - 1]);
        tw_dft<R / 2, HALF>(v, pw);
        double2 pw2[LG - 1];
        pw2[0] = mul_root<R>(pw[0]);
        if constexpr (LG - 1 > 1) pw2[1] = mul_root<R / 2>(pw[1]);
        if constexpr (LG - 1 > 2) pw2[2] = mul_root<R / 4>(pw[2]);
        tw_dft<R / 2, HALF>(v + R / 2, pw2);
    }
}
template <int R>
__host__ __device__ __forceinline__ constexpr int bitrev(int k)
{
    int r = 0;
    for (int b = 0; b < ilog2c(R); b++) r |= ((k >> b) & 1) << (ilog2c(R) - 1 - b);
    return r;
}

// plain in-register inverse DFT of 16 points, natural order in and out
__host__ __device__ __forceinline__ void dft16(double2 (&v)[16])
{
    double2 e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        e[k] = v[2 * k];
        o[k] = v[2 * k + 1];
    }
    prn::dft<8>(e);
    prn::dft<8>(o);
    constexpr double H = 0.70710678118654752440, C = 0.92387953251128675613, S = 0.38268343236508977173;
    // o[k] * exp(2 pi i k / 16)
    const double2 t0 = o[0];
    const double2 t1 = make_double2(fma(C, o[1].x, -(S * o[1].y)), fma(C, o[1].y, S * o[1].x));
    const double2 t2 = make_double2(H * (o[2].x - o[2].y), H * (o[2].x + o[2].y));
    const double2 t3 = make_double2(fma(S, o[3].x, -(C * o[3].y)), fma(S, o[3].y, C * o[3].x));
    const double2 t4 = make_double2(-o[4].y, o[4].x);
    const double2 t5 = make_double2(-fma(S, o[5].x, C * o[5].y), fma(C, o[5].x, -(S * o[5].y)));
    const double2 t6 = make_double2(-H * (o[6].x + o[6].y), H * (o[6].x - o[6].y));
    const double2 t7 = make_double2(-fma(C, o[7].x, S * o[7].y), fma(S, o[7].x, -(C * o[7].y)));
    const double2 t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = cadd(e[k], t[k]);
        v[k + 8] = csub(e[k], t[k]);
    }
}

// First-stage twiddles of one butterfly: t[k] = t0 g^k, times h from position w on (the support index wraps there;
// w = lh when flag, lh + 1 otherwise, never 0).  Two independent chains (k < 8 from t0, k >= 8 from t0 g^8) halve
// the depth of the dependent complex multiplies that every part starts with.
PFX_V4_HD __forceinline__ void first_stage_twiddles(double2 t0, double2 g, double2 h, int lh, bool flag, double2 (&t)[16])
{
    const double2 one = make_double2(1., 0.);
    const double2 ha = flag ? h : one, hb = flag ? one : h;
#if PFX_V4_SPLIT_CHAIN
    const int w = flag ? lh : lh + 1;
    const double2 g8 = csqr(csqr(csqr(g)));
    double2 a = t0, b = cmul(cmul(t0, g8), (w >= 1 && w <= 8) ? h : one);  // w = 0: input 0 already lies past the wrap
    t[0] = a;
    t[8] = b;
#pragma unroll
    for (int k = 1; k < 8; k++) {
        a = cmul(a, g);
        b = cmul(b, g);
        if (k == lh) a = cmul(a, ha);
        if (k == lh + 1) a = cmul(a, hb);
        if (k + 8 == lh) b = cmul(b, ha);
        if (k + 8 == lh + 1) b = cmul(b, hb);
        t[k] = a;
        t[k + 8] = b;
    }
#else
    double2 c = t0;
    t[0] = c;
#pragma unroll
    for (int k = 1; k < 16; k++) {
        c = cmul(c, g);
        if (k == lh) c = cmul(c, ha);
        if (k == lh + 1) c = cmul(c, hb);
        t[k] = c;
    }
#endif
}

// Stage S >= 1 of a part, in place; the last stage stores through store(q, rho, v).
template <int LOGL, int S, bool HALF, class Store>
__device__ __forceinline__ void stage(double2 *sm, const double2 *stw_all, const Store &store)
{
    using G = Geo<LOGL>;
    constexpr int LR = lrad(LOGL, S), R = 1 << LR, LOGM = logm(LOGL, S), M = 1 << LOGM;
    constexpr bool LAST = (S == G::NST - 1);
    constexpr int NBF = (1 << LOGCH) >> LR;  // butterflies per line part
    constexpr int IPT = NBF / NT;
    static_assert(IPT >= 1, "a stage has at least one butterfly per thread");
    const double2 *stw = stw_all + stw_off(LOGL, S);
#pragma unroll
    for (int it = 0; it < IPT; it++) {
        const int b = it * NT + threadIdx.x;
        const int rho = b & (G::RC - 1);
        const int u = b >> G::LOGRC;
        const int q0 = u & (M - 1);
        const int base = ((u >> LOGM) << (LOGM + LR)) + q0;
        double2 pw[LR];
        pw[0] = stw[q0];
#pragma unroll
        for (int i = 1; i < LR; i++) pw[i] = csqr(pw[i - 1]);
        double2 *p = sm + rho;
        double2 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = p[padded<G::PAD>(base + k * M) << G::LOGRC];
        if constexpr (LAST) {
            if (HALF) {
                tw_dft<R, true>(v, pw);
#pragma unroll
                for (int k = 0; k < R / 2; k++) store(base + k * M, rho, v[bitrev<R>(k)]);
            } else {
                tw_dft<R, false>(v, pw);
#pragma unroll
                for (int k = 0; k < R; k++) store(base + k * M, rho, v[bitrev<R>(k)]);
            }
        } else {
            tw_dft<R, false>(v, pw);
#pragma unroll
            for (int k = 0; k < R; k++) p[padded<G::PAD>(base + k * M) << G::LOGRC] = v[bitrev<R>(k)];
        }
    }
}

template <int LOGL, int S, bool HALF, class Store>
__device__ __forceinline__ void run_stages(double2 *sm, const double2 *stw, const Store &store)
{
    if constexpr (S < Geo<LOGL>::NST) {
        __syncthreads();
        stage<LOGL, S, HALF, Store>(sm, stw, store);
        run_stages<LOGL, S + 1, HALF, Store>(sm, stw, store);
    }
}

// Shared driver of both passes.  grp.load(g, m) returns the weighted support sample m of line g;
// grp.store(g, q, rho_global, v) consumes output element q of residue rho_global of that line.
// MULTI: the line has more than one part (N > CH); then the support samples stay in registers across the
// parts and the first-stage twiddles advance by one complex multiply per part.
template <int LOGL, bool MULTI, class Group>
__device__ __forceinline__ void run_lines(double2 *sm, const double2 *stw, const AxisDesc &A, int nparts, int lo, int K,
                                          const double2 *__restrict__ root, Group &grp)
{
    using G = Geo<LOGL>;
    constexpr int STEP = G::STEP, L = G::L;
    const int Nm = A.N - 1;
    const bool half = 2 * A.nout == A.N;  // power-of-two grid: outputs q >= L/2 are never stored
    const int l0 = lo & (L - 1), lh = l0 >> G::LOGBF, ll = l0 & (STEP - 1);

    // this thread's first-stage butterfly (u, rho): fixed for the whole kernel
    const int rho = threadIdx.x & (G::RC - 1), u = threadIdx.x >> G::LOGRC;
    const int nlow = digit_reverse<LOGL>(u);
    const int m0 = (nlow - lo) & (L - 1);  // support index of input 0
    const bool flag = nlow >= ll;          // which of the two candidate positions holds the wrap of the support index
    const int a0 = lo + m0;
    double2 *const out0 = sm + (padded<G::PAD>(u << 4) << G::LOGRC) + rho;
    // t0 = w_N^(a0 (rho+1)); g = w_N^(STEP (rho+1)): ratio between consecutive inputs; h = w_N^(-L (rho+1)): wrap
    const double2 t0i = __ldg(root + (int)(((long long)a0 * (rho + 1)) & Nm));
    const double2 g0i = __ldg(root + (((rho + 1) << G::LOGBF) & Nm));
    const double2 h0i = __ldg(root + ((A.N - (((rho + 1) << LOGL) & Nm)) & Nm));

    double2 x[MULTI ? 16 : 1];
    for (int gi = grp.first(); grp.valid(gi); gi = grp.next(gi)) {
        if (MULTI) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int m = (m0 + k * STEP) & (L - 1);
                x[k] = (m < K) ? grp.load(gi, m) : make_double2(0., 0.);
            }
        }
        double2 t0 = t0i, g = g0i, h = h0i;
        for (int part = 0; part < (MULTI ? nparts : 1); part++) {
            // ---- first stage: twiddle the support samples, radix-16 butterfly into shared memory
            {
                double2 v[16], tw[16];
                first_stage_twiddles(t0, g, h, lh, flag, tw);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    double2 xin;
                    if (MULTI) {
                        xin = x[k];
                    } else {
                        const int m = (m0 + k * STEP) & (L - 1);
                        xin = (m < K) ? grp.load(gi, m) : make_double2(0., 0.);
                    }
                    v[k] = cmul(xin, tw[k]);
                }
                dft16(v);
#pragma unroll
                for (int k = 0; k < 16; k++) out0[k << G::LOGRC] = v[k];
            }
            if (MULTI && part + 1 < nparts) {
                // rho -> rho + RC
                t0 = cmul(t0, __ldg(root + (int)(((long long)a0 * G::RC) & Nm)));
                g = cmul(g, __ldg(root + ((G::RC << G::LOGBF) & Nm)));
                h = cmul(h, __ldg(root + ((A.N - ((G::RC << LOGL) & Nm)) & Nm)));
            }
            const int rho0 = part << G::LOGRC;
            auto store = [&](int q, int rr, double2 v) { grp.store(gi, q, rho0 + rr, v); };
            if (half)
                run_stages<LOGL, 1, true>(sm, stw, store);
            else
                run_stages<LOGL, 1, false>(sm, stw, store);
            __syncthreads();  // the next first stage overwrites the plane
        }
    }
}

// ---- pass 2: x-direction transforms, one output row per group ------------------------------------------------
struct Pass2Group {
    const double2 *Tt;  // T[t][m][j]
    double2 *Wt;        // W[t][j][i]
    int ny, nx, logr;
    __device__ __forceinline__ int first() const { return blockIdx.x; }
    __device__ __forceinline__ bool valid(int g) const { return g < ny; }
    __device__ __forceinline__ int next(int g) const { return g + gridDim.x; }
    __device__ __forceinline__ double2 load(int j, int m) const { return __ldg(Tt + (size_t)m * ny + j); }
    __device__ __forceinline__ void store(int j, int q, int rho, double2 v) const
    {
        const int i = (q << logr) + rho;
        if (i < nx) Wt[(size_t)j * nx + i] = v;
    }
};

template <int LOGL, bool MULTI>
__global__ void __launch_bounds__(NT, MULTI ? PFX_V4_MULTI_OCC : 4) pass2_kernel(V4Args a)
{
    using G = Geo<LOGL>;
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ax;
    const int t = blockIdx.y, ia = t % PFX_NA;
    double2 *stw = smem + G::PLANE;
    load_stage_twiddles<LOGL>(stw, A.logN, a.p.rootx);
    Pass2Group grp;
    grp.Tt = a.p.T + (size_t)t * sd.kamax * a.p.ny;
    grp.Wt = a.p.W + (size_t)t * a.p.ny * a.p.nx;
    grp.ny = a.p.ny;
    grp.nx = a.p.nx;
    grp.logr = a.logr;
    __syncthreads();
    run_lines<LOGL, MULTI>(smem, stw, A, a.nparts, sd.alo[ia], sd.ka[ia], a.p.rootx, grp);
}

// ---- pass 1: y-direction transforms, one spectrum column per group -------------------------------------------
template <bool DCT>
struct Pass1Group {
    const double2 *spec;  // padded spectrum [a][b] (y-wavenumber fastest)
    const double *tv;     // v1 weights of this azimuth [L]
    const double *tu;     // (c/P) u1 weights of this azimuth [Lx]
    double2 *T;           // intermediate (layout: prn_t_column)
    const PrunedArgs *pa;
    const double *dct;         // cosine-transform spectrum D[|a|][|b|] of this grid, or null (pfx_dct.cu)
    const double2 *phx, *phy;  // ... and its phases e^{i pi a/NX}, e^{i pi b/NY}
    int t;
    int NX, NYm, ny, alo, ka, blo, logr;
    __device__ __forceinline__ int first() const { return blockIdx.x; }
    __device__ __forceinline__ bool valid(int g) const { return g < ka; }
    __device__ __forceinline__ int next(int g) const { return g + gridDim.x; }
    __device__ __forceinline__ double2 load(int m, int mb) const
    {
        const double wv = tv[mb];
        if (DCT) {
            const int a = alo + m, b = blo + mb, aa = a < 0 ? -a : a, bb = b < 0 ? -b : b;
            if (2 * aa >= NX || bb >= ny) return make_double2(0., 0.);  // D[nx][.] = D[.][ny] = 0
            const double d = __ldg(dct + (size_t)aa * ny + bb) * wv;
            const double2 ph = __ldg(phy + (b & NYm));
            return make_double2(d * ph.x, d * ph.y);
        }
        const double2 x = __ldg(spec + (size_t)((alo + m) & (NX - 1)) * (NYm + 1) + ((blo + mb) & NYm));
        return make_double2(x.x * wv, x.y * wv);
    }
    __device__ __forceinline__ void store(int m, int q, int rho, double2 v) const
    {
        const int j = (q << logr) + rho;
        if (j < ny) {
            const double wu = tu[m];
            size_t js;
            const size_t o = prn_t_column(*pa, t, m, alo, js);
            if (DCT) {
                const double2 ph = __ldg(phx + ((alo + m) & (NX - 1)));
                const double2 w = make_double2(ph.x * wu, ph.y * wu);
                T[o + (size_t)j * js] = make_double2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
            } else {
                T[o + (size_t)j * js] = make_double2(v.x * wu, v.y * wu);
            }
        }
    }
};

template <int LOGL, bool MULTI, bool DCT>
__global__ void __launch_bounds__(NT, MULTI ? 3 : 4) pass1_kernel(V4Args a)
{
    using G = Geo<LOGL>;
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ay;
    const int t = blockIdx.y, ia = t % PFX_NA, g = t / PFX_NA;
    if ((int)blockIdx.x >= sd.ka[ia]) return;
    double2 *stw = smem + G::PLANE;
    load_stage_twiddles<LOGL>(stw, A.logN, a.p.rooty);
    Pass1Group<DCT> grp;
    grp.spec = g ? a.p.spec1 : a.p.spec0;
    grp.dct = g ? a.p.dct1 : a.p.dct0;
    grp.phx = a.p.phx;
    grp.phy = a.p.phy;
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
    grp.logr = a.logr;
    __syncthreads();
    run_lines<LOGL, MULTI>(smem, stw, A, a.nparts, sd.blo[ia], sd.kb[ia], a.p.rooty, grp);
}

template <int LOGL>
inline size_t smem_bytes()
{
    return ((size_t)Geo<LOGL>::PLANE + Geo<LOGL>::STW) * sizeof(double2);
}

}  // namespace v4

// launchers (pfx_pruned_v4_p1.cu, pfx_pruned_v4_p2.cu); resident receives the CTAs of this kernel per SM
int v4_configure_p1(int logL, int multi, int dct, int *resident);
int v4_configure_p2(int logL, int multi, int *resident);
int v4_launch_p1(const v4::V4Args &a, int logL, int multi, int dct, dim3 grid, cudaStream_t st);
int v4_launch_p2(const v4::V4Args &a, int logL, int multi, dim3 grid, cudaStream_t st);

}  // namespace pfx
