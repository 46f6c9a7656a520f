// pfx_flex.cuh -- the flexural model of src/flex/flex.f90 as a device function,
// with analytic first derivatives with respect to (Te, F, alpha).
//
//   flexfilter_top  flex.f90:74-86     theta = -r / (1 + psi/(drho*g))
//   flexfilter_bot  flex.f90:94-106    phi   = -r * (1 + psi/((rhoc-rhof)*g))
//   decon           flex.f90:115-132   mu_h, mu_w, nu_h, nu_w
//   tr_func         flex.f90:141-169   hg, hh, gg -> admittance, coherence
//   real_xspec_functions flex.f90:178-235
//
// Only cos(alpha) enters the real admittance and the squared-real coherency
// (flex.f90:158-165,231).
#pragma once
#include "pfx_common.cuh"

namespace pfx {

// flex.f90:39-43
constexpr double FLEX_PI = 3.141592653589793;
constexpr double FLEX_G = 9.81;
constexpr double FLEX_EM = 1.e11;
constexpr double FLEX_NU = 0.25;
constexpr double FLEX_GC = 6.67e-6;

#define PFX_HD __host__ __device__

// Per-cell constants: everything that does not depend on the wavenumber or on (Te, F, alpha).
struct FlexCell {
    double drho;  // rhom - rhoc
    double rcf;   // rhoc - rhof, rhof chosen from wd (flex.f90:207-211)
    double r;     // rcf / drho
    double A;     // 0 Bouguer, 1 free air (flex.f90:200-204)
    double wd, zc;
    double idg;   // 1 / ((rhom - rhoc) * g)
    double irg;   // 1 / ((rhoc - rhof) * g)
};

__host__ __device__ inline FlexCell flex_cell(const pfx_flex_conf &c, double rhoc, double zc, double wd)
{
    FlexCell f;
    const double rhof = (wd > 0.) ? c.rhow : c.rhoa;
    f.drho = c.rhom - rhoc;
    f.rcf = rhoc - rhof;
    f.r = f.rcf / f.drho;
    f.A = (c.boug == 1) ? 0. : 1.;
    f.wd = wd;
    f.zc = zc;
    f.idg = 1. / (f.drho * FLEX_G);
    f.irg = 1. / (f.rcf * FLEX_G);
    return f;
}

// Per-(cell, wavenumber) constants.
struct FlexK {
    double k4d;   // k^4 * Em / 12 / (1 - nu^2) * 1e9   => psi = k4d * Te_km^3
    double top;   // 2 pi G * A * (rhoc-rhof) * exp(-k*wd)
    double deep;  // 2 pi G * (rhom-rhoc) * exp(-k*(zc+wd))
};

PFX_HD __forceinline__ FlexK flex_k(const FlexCell &f, double k)
{
    FlexK q;
    const double k2 = k * k;
    q.k4d = (k2 * k2) * (FLEX_EM * 1.e9 / 12. / (1. - FLEX_NU * FLEX_NU));
    q.top = 2. * FLEX_PI * FLEX_GC * (f.A * f.rcf * exp(-k * f.wd));
    q.deep = 2. * FLEX_PI * FLEX_GC * (f.drho * exp(-k * (f.zc + f.wd)));
    return q;
}

// 1/x to ~1 ulp from the hardware seed (rcp.approx.ftz.f64) and two Newton steps, without the
// special-case branches of the IEEE division sequence (the arguments here are finite and non-zero).
PFX_HD __forceinline__ double flex_rcp(double x)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;  // host builds of the model (tests/flex_emul.cu)
#endif
}

// Two wavenumbers side by side: every operation of the model on a P2 is two INDEPENDENT double operations next to
// each other in program order, which is what lets a lone warp keep the FP64 pipe busy through the model's long
// dependency chain (two reciprocals and ~40 dependent operations per wavenumber).  Each lane performs exactly the
// scalar sequence, so results are those of the one-wavenumber form.
struct P2 {
    double a, b;
};
PFX_HD __forceinline__ P2 operator+(P2 x, P2 y) { return {x.a + y.a, x.b + y.b}; }
PFX_HD __forceinline__ P2 operator-(P2 x, P2 y) { return {x.a - y.a, x.b - y.b}; }
PFX_HD __forceinline__ P2 operator*(P2 x, P2 y) { return {x.a * y.a, x.b * y.b}; }
PFX_HD __forceinline__ P2 operator+(P2 x, double y) { return {x.a + y, x.b + y}; }
PFX_HD __forceinline__ P2 operator+(double y, P2 x) { return {y + x.a, y + x.b}; }
PFX_HD __forceinline__ P2 operator-(P2 x, double y) { return {x.a - y, x.b - y}; }
PFX_HD __forceinline__ P2 operator-(double y, P2 x) { return {y - x.a, y - x.b}; }
PFX_HD __forceinline__ P2 operator*(P2 x, double y) { return {x.a * y, x.b * y}; }
PFX_HD __forceinline__ P2 operator*(double y, P2 x) { return {y * x.a, y * x.b}; }
PFX_HD __forceinline__ P2 operator-(P2 x) { return {-x.a, -x.b}; }
PFX_HD __forceinline__ P2 flex_rcp(P2 x)
{
#ifdef __CUDA_ARCH__
    P2 r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r.a) : "d"(x.a));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r.b) : "d"(x.b));
    P2 e = {fma(-x.a, r.a, 1.0), fma(-x.b, r.b, 1.0)};
    r = {fma(r.a, e.a, r.a), fma(r.b, e.b, r.b)};
    e = {fma(-x.a, r.a, 1.0), fma(-x.b, r.b, 1.0)};
    return {fma(r.a, e.a, r.a), fma(r.b, e.b, r.b)};
#else
    return {1.0 / x.a, 1.0 / x.b};
#endif
}

// Admittance and coherence at one wavenumber (T = double) or two (T = P2); when JAC, also d/dTe, d/dF, d/dalpha.
//   k4d, top, deep: the FlexK constants of the wavenumber(s)
//   te3 = Te^3 (km^3), ff = F/(1-F), dff = 1/(1-F)^2, ca = cos(alpha), sa = sin(alpha)
// The seven quotients of the Fortran are formed from two reciprocals (of qq*(qq+r)*(phi-1) and of
// hh*gg); results agree with the IEEE-division form to a few ulp (tested at 1e-12 against the oracle).
// Written in the four amplitudes  A = mu_h, B = mu_w t, C = nu_h, D = nu_w t  (t = f r, flex.f90:155-156):
//   hg = CA + DB + ca (CB + DA),  hh = A^2 + B^2 + 2 ca AB,  gg = C^2 + D^2 + 2 ca CD          (flex.f90:157-161)
// so that the load ratio enters once per amplitude and every derivative is a short sum of products.
template <bool JAC, class T>
PFX_HD __forceinline__ void flex_eval_g(const FlexCell &f, const T k4d, const T top, const T deep, double te,
                                        double te3, double ff, double dff, double ca, double sa, T &admit, T &coh,
                                        T *dadm, T *dcoh)
{
    const double idg = f.idg;
    const T psi = k4d * te3;                         // flex.f90:217-220
    const T qq = 1. + psi * idg;
    const T phi = -f.r * (1. + psi * f.irg);         // flex.f90:103
    // theta = -r/qq (flex.f90:83); 1 - theta = (qq + r)/qq
    const T qr = qq + f.r, pm = phi - 1.;
    const T qqr = qq * qr;
    const T inv3 = flex_rcp(qqr * pm);
    const T qpi = qq * pm * inv3;          // 1/(qq + r)
    const T iqq = qr * pm * inv3;          // 1/qq
    const T A = qq * qpi;                  // mu_h = 1/(1-theta) = qq/(qq+r), flex.f90:123
    const T mu_w = qqr * inv3;             // 1/(phi-1), flex.f90:124
    const T theta = -f.r * iqq;
    const T bh = top + deep * theta;  // flex.f90:125-128 (2 pi G folded into top/deep)
    const T bw = top + deep * phi;
    const T C = bh * A;               // nu_h
    const T nu_w = bw * mu_w;
    const double t = ff * f.r;  // flex.f90:155-156
    const double c2 = 2. * ca;
    const T B = mu_w * t, D = nu_w * t;
    const T AB = A * B, CD = C * D;
    const T X = C * B + D * A;                 // d hg / d ca
    const T hg = C * A + D * B + ca * X;       // Re(hg), flex.f90:157-159
    const T hh = A * A + B * B + c2 * AB;
    const T gg = C * C + D * D + c2 * CD;
    const T ihg = flex_rcp(hh * gg), ihh = gg * ihg;
    admit = hg * ihh;         // flex.f90:163, 231
    coh = (hg * hg) * ihg;    // Re(hg/sqrt(hh)/sqrt(gg))^2, flex.f90:164-165
    if (JAC) {
        // quotient rule with the common factors formed once:
        //   d admit = ihh dhg - (admit ihh) dhh,   d coh = (2 hg ihg) dhg - (coh/hh) dhh - (coh/gg) dgg
        const T a1 = admit * ihh;
        const T h2 = 2. * (hg * ihg), cg = coh * ihh, ch = coh * (hh * ihg);
        // --- d/dpsi of the amplitudes
        const T dtheta = -theta * iqq * idg;
        const T dA = A * A * dtheta;                // d mu_h
        const T dmu_w = mu_w * mu_w * idg;          // -mu_w^2 dphi, dphi = -idg
        const T dC = deep * dtheta * A + bh * dA;   // d nu_h
        const T dnu_w = bw * dmu_w - deep * idg * mu_w;
        const T dB = dmu_w * t, dD = dnu_w * t;
        const T dhg_p = dC * A + C * dA + dD * B + D * dB + ca * (dC * B + C * dB + dD * A + D * dA);
        const T dhh_p = 2. * (A * dA + B * dB) + c2 * (dA * B + A * dB);
        const T dgg_p = 2. * (C * dC + D * dD) + c2 * (dC * D + C * dD);
        const T dpsi = 3. * k4d * (te * te);
        dadm[0] = (ihh * dhg_p - a1 * dhh_p) * dpsi;
        dcoh[0] = (h2 * dhg_p - cg * dhh_p - ch * dgg_p) * dpsi;
        // --- d/dt (dB/dt = mu_w, dD/dt = nu_w), times dt/dF
        const T dhg_t = 2. * (D * mu_w) + ca * (C * mu_w + nu_w * A);
        const T dhh_t = 2. * mu_w * (B + ca * A);
        const T dgg_t = 2. * nu_w * (D + ca * C);
        const double dt = f.r * dff;
        dadm[1] = (ihh * dhg_t - a1 * dhh_t) * dt;
        dcoh[1] = (h2 * dhg_t - cg * dhh_t - ch * dgg_t) * dt;
        // --- d/dcos(alpha), times d cos(alpha)/d alpha
        const double dc = -sa;
        dadm[2] = (ihh * X - a1 * (2. * AB)) * dc;
        dcoh[2] = (h2 * X - cg * (2. * AB) - ch * (2. * CD)) * dc;
    }
}

template <bool JAC>
PFX_HD __forceinline__ void flex_eval(const FlexCell &f, const FlexK &q, double te, double te3, double ff,
                                          double dff, double ca, double sa, double &admit, double &coh, double *dadm,
                                          double *dcoh)
{
    flex_eval_g<JAC, double>(f, q.k4d, q.top, q.deep, te, te3, ff, dff, ca, sa, admit, coh, dadm, dcoh);
}

}  // namespace pfx
