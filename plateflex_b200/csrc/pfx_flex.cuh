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

__device__ __forceinline__ FlexK flex_k(const FlexCell &f, double k)
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
__device__ __forceinline__ double flex_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// Admittance and coherence at one wavenumber; when JAC, also d/dTe, d/dF, d/dalpha.
//   te3 = Te^3 (km^3), ff = F/(1-F), dff = 1/(1-F)^2, ca = cos(alpha), sa = sin(alpha)
// The seven quotients of the Fortran are formed from two reciprocals (of qq*(qq+r)*(phi-1) and of
// hh*gg); results agree with the IEEE-division form to a few ulp (tested at 1e-12 against the oracle).
template <bool JAC>
__device__ __forceinline__ void flex_eval(const FlexCell &f, const FlexK &q, double te, double te3, double ff,
                                          double dff, double ca, double sa, double &admit, double &coh, double *dadm,
                                          double *dcoh)
{
    const double idg = f.idg;
    const double psi = q.k4d * te3;                         // flex.f90:217-220
    const double qq = 1. + psi * idg;
    const double phi = -f.r * (1. + psi * f.irg);           // flex.f90:103
    // theta = -r/qq (flex.f90:83); 1 - theta = (qq + r)/qq
    const double qr = qq + f.r, pm = phi - 1.;
    const double inv3 = flex_rcp(qq * qr * pm);
    const double iqq = qr * pm * inv3;     // 1/qq
    const double mu_h = qq * (qq * pm * inv3);  // 1/(1-theta) = qq/(qq+r), flex.f90:123
    const double mu_w = qq * qr * inv3;    // 1/(phi-1), flex.f90:124
    const double theta = -f.r * iqq;
    const double bh = q.top + q.deep * theta;  // flex.f90:125-128 (2 pi G folded into top/deep)
    const double bw = q.top + q.deep * phi;
    const double nu_h = bh * mu_h;
    const double nu_w = bw * mu_w;
    const double t = ff * f.r;  // flex.f90:155-156
    const double tc = t * ca;
    const double cross_hg = nu_h * mu_w + nu_w * mu_h;
    const double hg = nu_h * mu_h + nu_w * mu_w * (t * t) + cross_hg * tc;  // Re(hg), flex.f90:157-159
    const double hh = mu_h * mu_h + (mu_w * t) * (mu_w * t) + 2. * mu_h * mu_w * tc;
    const double gg = nu_h * nu_h + (nu_w * t) * (nu_w * t) + 2. * nu_h * nu_w * tc;
    const double ihg = flex_rcp(hh * gg), ihh = gg * ihg;
    admit = hg * ihh;         // flex.f90:163, 231
    coh = (hg * hg) * ihg;    // Re(hg/sqrt(hh)/sqrt(gg))^2, flex.f90:164-165
    if (JAC) {
        // --- d/dpsi
        const double dtheta = -theta * iqq * idg;
        const double dphi = -idg;
        const double dmu_h = mu_h * mu_h * dtheta;
        const double dmu_w = -mu_w * mu_w * dphi;
        const double dnu_h = q.deep * dtheta * mu_h + bh * dmu_h;
        const double dnu_w = q.deep * dphi * mu_w + bw * dmu_w;
        const double dhg_p = dnu_h * mu_h + nu_h * dmu_h + (dnu_w * mu_w + nu_w * dmu_w) * (t * t) +
                             (dnu_h * mu_w + nu_h * dmu_w + dnu_w * mu_h + nu_w * dmu_h) * tc;
        const double dhh_p = 2. * (mu_h * dmu_h + mu_w * dmu_w * (t * t) + (dmu_h * mu_w + mu_h * dmu_w) * tc);
        const double dgg_p = 2. * (nu_h * dnu_h + nu_w * dnu_w * (t * t) + (dnu_h * nu_w + nu_h * dnu_w) * tc);
        const double dpsi = 3. * q.k4d * (te * te);
        // --- d/dt
        const double dhg_t = 2. * nu_w * mu_w * t + cross_hg * ca;
        const double dhh_t = 2. * (mu_w * mu_w * t + mu_h * mu_w * ca);
        const double dgg_t = 2. * (nu_w * nu_w * t + nu_h * nu_w * ca);
        const double dt = f.r * dff;
        // --- d/dcos(alpha)
        const double dhg_c = cross_hg * t;
        const double dhh_c = 2. * mu_h * mu_w * t;
        const double dgg_c = 2. * nu_h * nu_w * t;
        const double dc = -sa;
        const double dhg[3] = {dhg_p * dpsi, dhg_t * dt, dhg_c * dc};
        const double dhh[3] = {dhh_p * dpsi, dhh_t * dt, dhh_c * dc};
        const double dgg[3] = {dgg_p * dpsi, dgg_t * dt, dgg_c * dc};
#pragma unroll
        for (int p = 0; p < 3; p++) {
            dadm[p] = (dhg[p] - admit * dhh[p]) * ihh;
            dcoh[p] = (2. * hg * dhg[p] - coh * (dhh[p] * gg + hh * dgg[p])) * ihg;
        }
    }
}

}  // namespace pfx
