/*
 * oracle/pfx_oracle.c -- CPU oracle for the PlateFlex hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the checker and the timed
 * CPU baseline.  The product (plateflex_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" for CWT / jackknife / flexure values (the
 * reference holds no golden vectors for them and its Fortran cannot be built
 * in this image: no gfortran/flang, no meson).  See pfx_oracle_impl.h.
 *
 * Exports (C ABI, column-major arrays as in the Fortran):
 *   ora32_*  float32 storage/arithmetic, float64 FFT twiddle recurrence --
 *            the reference's own precision (cpwt.f90 default-kind REAL)
 *   ora64_*  all float64 (long double twiddle recurrence) -- parity target
 *   ora_flex_* the flexural model of src/flex/flex.f90 (float64 like the source)
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- float32 instantiation ---------------- */
#define REAL float
#define TWID double
#define SUMT float
#define SUF 32
#define XCOS cosf
#define XSIN sinf
#define XEXP expf
#define XLOG logf
#define XSQRT sqrtf
#define XFABS fabsf
#define TWSIN sin
#include "pfx_oracle_impl.h"
#undef REAL
#undef TWID
#undef SUMT
#undef SUF
#undef XCOS
#undef XSIN
#undef XEXP
#undef XLOG
#undef XSQRT
#undef XFABS
#undef TWSIN

/* ---------------- float64 instantiation ---------------- */
#define REAL double
#define TWID long double
#define SUMT long double
#define SUF 64
#define XCOS cos
#define XSIN sin
#define XEXP exp
#define XLOG log
#define XSQRT sqrt
#define XFABS fabs
#define TWSIN sinl
#include "pfx_oracle_impl.h"
#undef REAL
#undef TWID
#undef SUMT
#undef SUF
#undef XCOS
#undef XSIN
#undef XEXP
#undef XLOG
#undef XSQRT
#undef XFABS
#undef TWSIN

/* ======================================================================
 * Flexural model -- src/flex/flex.f90 (all DOUBLE PRECISION there).
 * The Fortran keeps rhoc..boug in module conf_flex (flex.f90:47-48); here
 * they travel in a struct so the oracle is re-entrant.
 * ====================================================================== */
typedef struct {
    double rhoc, rhom, rhof, rhow, rhoa, wd, zc;
    int boug;
} ora_flex_conf;

/* flex.f90:39-43 */
static const double FLEX_PI = 3.141592653589793;
static const double FLEX_G = 9.81;
static const double FLEX_EM = 1.e11;
static const double FLEX_NU = 0.25;
static const double FLEX_GC = 6.67e-6;

/* flex.f90:74-86 */
void ora_flex_flexfilter_top(const ora_flex_conf *c, int ns, const double *psi, double *filt)
{
    for (int i = 0; i < ns; i++)
        filt[i] = -((c->rhoc - c->rhof) / (c->rhom - c->rhoc)) * pow(1. + psi[i] / (c->rhom - c->rhoc) / FLEX_G, -1.);
}

/* flex.f90:94-106 */
void ora_flex_flexfilter_bot(const ora_flex_conf *c, int ns, const double *psi, double *filt)
{
    for (int i = 0; i < ns; i++)
        filt[i] = -((c->rhoc - c->rhof) / (c->rhom - c->rhoc)) * (1. + psi[i] / (c->rhoc - c->rhof) / FLEX_G);
}

/* flex.f90:115-132 */
void ora_flex_decon(const ora_flex_conf *c, int ns, const double *theta, const double *phi, const double *k, double A,
                    double *mu_h, double *mu_w, double *nu_h, double *nu_w)
{
    for (int i = 0; i < ns; i++) {
        mu_h[i] = 1. / (1. - theta[i]);
        mu_w[i] = 1. / (phi[i] - 1.);
        double top = exp(-k[i] * c->wd), deep = exp(-k[i] * (c->zc + c->wd));
        nu_h[i] = 2. * FLEX_PI * FLEX_GC * (A * (c->rhoc - c->rhof) * top + (c->rhom - c->rhoc) * theta[i] * deep);
        nu_h[i] = nu_h[i] / (1. - theta[i]);
        nu_w[i] = 2. * FLEX_PI * FLEX_GC * (A * (c->rhoc - c->rhof) * top + (c->rhom - c->rhoc) * phi[i] * deep);
        nu_w[i] = nu_w[i] / (phi[i] - 1.);
    }
}

/* flex.f90:141-169.  admit is complex (re,im interleaved); coh = Re(coherency)^2. */
void ora_flex_tr_func(const ora_flex_conf *c, int ns, const double *mu_h, const double *mu_w, const double *nu_h,
                      const double *nu_w, double F, double alpha, double *admit_c, double *coh)
{
    double r = (c->rhoc - c->rhof) / (c->rhom - c->rhoc);
    double ff = F / (1. - F);
    double ca = cos(alpha), sa = sin(alpha);
    for (int i = 0; i < ns; i++) {
        double hg_re = nu_h[i] * mu_h[i] + nu_w[i] * mu_w[i] * (ff * ff) * pow(r, 2.) +
                       (nu_h[i] * mu_w[i] + nu_w[i] * mu_h[i]) * ff * r * ca;
        double hg_im = (nu_h[i] * mu_w[i] - nu_w[i] * mu_h[i]) * ff * r * sa;
        double hh = mu_h[i] * mu_h[i] + pow(mu_w[i] * ff * r, 2.) + 2. * mu_h[i] * mu_w[i] * ff * r * ca;
        double gg = nu_h[i] * nu_h[i] + pow(nu_w[i] * ff * r, 2.) + 2. * nu_h[i] * nu_w[i] * ff * r * ca;
        admit_c[2 * i] = hg_re / hh;
        admit_c[2 * i + 1] = hg_im / hh;
        double cohy_re = hg_re / sqrt(hh) / sqrt(gg);
        coh[i] = pow(cohy_re, 2.);
    }
}

/* flex.f90:178-235.  Mutates conf->rhof exactly as the Fortran mutates the
 * module variable (flex.f90:207-211). */
void ora_flex_real_xspec_functions(ora_flex_conf *c, int ns, const double *k, double Te, double F, double alpha,
                                   double *admit, double *coh)
{
    double A = (c->boug == 1) ? 0. : 1.;
    c->rhof = (c->wd > 0.) ? c->rhow : c->rhoa;
    double Te_meters = Te * 1.e3;
    double D = FLEX_EM * (Te_meters * Te_meters * Te_meters) / 12. / (1. - pow(FLEX_NU, 2.));
    double *w = (double *)calloc((size_t)ns * 9, sizeof(double));
    double *psi = w, *theta = w + ns, *phi = w + 2 * ns, *mu_h = w + 3 * ns, *mu_w = w + 4 * ns, *nu_h = w + 5 * ns,
           *nu_w = w + 6 * ns, *cad = w + 7 * ns;
    for (int i = 0; i < ns; i++) psi[i] = D * pow(k[i], 4.);
    ora_flex_flexfilter_top(c, ns, psi, theta);
    ora_flex_flexfilter_bot(c, ns, psi, phi);
    ora_flex_decon(c, ns, theta, phi, k, A, mu_h, mu_w, nu_h, nu_w);
    ora_flex_tr_func(c, ns, mu_h, mu_w, nu_h, nu_w, F, alpha, cad, coh);
    for (int i = 0; i < ns; i++) admit[i] = cad[2 * i];
    free(w);
}
