// tests/flex_emul.cu -- host check of the flexure model of pfx_flex.cuh (test infrastructure, CPU only).
// flex_eval_g is compiled for the host and compared with (1) a literal restatement of flex.f90:74-169, 178-235 in
// its original quotient form, (2) central differences of that restatement for the analytic Jacobian, and (3) itself
// in the two-wavenumber form (P2), whose lanes must reproduce the scalar results bit for bit.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../plateflex_b200/csrc/pfx_flex.cuh"

using namespace pfx;

struct Conf {
    double rhom, rhoc, rhof, wd, zc;
    int boug;
};

// flex.f90 as written: flexfilter_top/bot, decon, tr_func (real admittance, squared-real coherency)
static void literal(const Conf &c, double k, double te_km, double F, double alpha, double &adm, double &coh)
{
    const double pi = 3.141592653589793, g = 9.81, em = 1.e11, nu = 0.25, G = 6.67e-6;
    const double te = te_km * 1.e3;
    const double d = em * te * te * te / 12. / (1. - nu * nu);
    const double psi = d * k * k * k * k;
    const double drho = c.rhom - c.rhoc, rcf = c.rhoc - c.rhof, r = rcf / drho;
    const double theta = -r / (1. + psi / (drho * g));      // flexfilter_top
    const double phi = -r * (1. + psi / (rcf * g));          // flexfilter_bot
    const double A = (c.boug == 1) ? 0. : 1.;
    const double mu_h = 1. / (1. - theta), mu_w = 1. / (phi - 1.);
    const double nu_h = 2. * pi * G * (A * rcf * std::exp(-k * c.wd) + drho * theta * std::exp(-k * (c.zc + c.wd))) / (1. - theta);
    const double nu_w = 2. * pi * G * (A * rcf * std::exp(-k * c.wd) + drho * phi * std::exp(-k * (c.zc + c.wd))) / (phi - 1.);
    const double f = F / (1. - F), rr = f * r, ca = std::cos(alpha);
    const double hg = nu_h * mu_h + nu_w * mu_w * rr * rr + (nu_h * mu_w + nu_w * mu_h) * rr * ca;
    const double hh = mu_h * mu_h + mu_w * mu_w * rr * rr + 2. * mu_h * mu_w * rr * ca;
    const double gg = nu_h * nu_h + nu_w * nu_w * rr * rr + 2. * nu_h * nu_w * rr * ca;
    adm = hg / hh;
    const double q = hg / std::sqrt(hh) / std::sqrt(gg);
    coh = q * q;
}

int main()
{
    int bad = 0;
    const Conf confs[] = {{3200., 2700., 0., 0., 35.e3, 1}, {3200., 2800., 1030., 4500., 20.e3, 0}, {3300., 2670., 1030., 800., 12.e3, 0},
                          {3200., 2700., 0., 0., 35.e3, 0}};
    double worst_v = 0., worst_j = 0.;
    for (const Conf &c : confs) {
        pfx_flex_conf pc;
        std::memset(&pc, 0, sizeof pc);
        pc.rhom = c.rhom;
        pc.rhoc = c.rhoc;
        pc.rhow = 1030.;
        pc.rhoa = 0.;
        pc.boug = c.boug;
        pc.zc = c.zc;
        const FlexCell fc = flex_cell(pc, c.rhoc, c.zc, c.wd);
        for (double te : {2.5, 20., 87., 240.})
            for (double F : {0.001, 0.3, 0.5, 0.93})
                for (double al : {0.2, 1.5707963267948966, 2.7})
                    for (double k = 4e-6; k < 7e-4; k *= 1.9) {
                        const FlexK q = flex_k(fc, k);
                        const double te3 = te * te * te, ff = F / (1. - F), dff = 1. / ((1. - F) * (1. - F));
                        double adm, coh, dA[3], dC[3];
                        flex_eval<true>(fc, q, te, te3, ff, dff, std::cos(al), std::sin(al), adm, coh, dA, dC);
                        double a0, c0;
                        literal(c, k, te, F, al, a0, c0);
                        const double ev = std::fmax(std::fabs(adm - a0) / std::fabs(a0), std::fabs(coh - c0) / std::fmax(c0, 1e-300));
                        worst_v = std::fmax(worst_v, ev);
                        if (!(ev < 1e-11)) {
                            if (bad < 10) std::printf("value te=%g F=%g al=%g k=%g: %.17g %.17g vs %.17g %.17g\n", te, F, al, k, adm, coh, a0, c0);
                            bad++;
                        }
                        // Jacobian vs central differences of the literal form
                        const double h[3] = {1e-5 * te, 1e-6, 1e-6};
                        for (int p = 0; p < 3; p++) {
                            double ap, cp, am, cm;
                            literal(c, k, te + (p == 0) * h[0], F + (p == 1) * h[1], al + (p == 2) * h[2], ap, cp);
                            literal(c, k, te - (p == 0) * h[0], F - (p == 1) * h[1], al - (p == 2) * h[2], am, cm);
                            const double da = (ap - am) / (2 * h[p]), dc = (cp - cm) / (2 * h[p]);
                            const double sa = std::fabs(a0) / (p == 0 ? te : 1.), sc = std::fmax(c0, 1e-12) / (p == 0 ? te : 1.);
                            const double ej = std::fmax(std::fabs(dA[p] - da) / (std::fabs(da) + 1e-4 * sa),
                                                        std::fabs(dC[p] - dc) / (std::fabs(dc) + 1e-4 * sc));
                            worst_j = std::fmax(worst_j, ej);
                            if (!(ej < 2e-4)) {
                                if (bad < 10) std::printf("jac p=%d te=%g F=%g al=%g k=%g: %g %g vs %g %g\n", p, te, F, al, k, dA[p], dC[p], da, dc);
                                bad++;
                            }
                        }
                        // two-wavenumber form: lane a = this k, lane b = 1.3 k, must equal the scalar results
                        const FlexK q2 = flex_k(fc, 1.3 * k);
                        P2 admp, cohp, dAp[3], dCp[3];
                        flex_eval_g<true, P2>(fc, P2{q.k4d, q2.k4d}, P2{q.top, q2.top}, P2{q.deep, q2.deep}, te, te3, ff, dff,
                                              std::cos(al), std::sin(al), admp, cohp, dAp, dCp);
                        double adm2, coh2, dA2[3], dC2[3];
                        flex_eval<true>(fc, q2, te, te3, ff, dff, std::cos(al), std::sin(al), adm2, coh2, dA2, dC2);
                        bool same = admp.a == adm && cohp.a == coh && admp.b == adm2 && cohp.b == coh2;
                        for (int p = 0; p < 3; p++) same = same && dAp[p].a == dA[p] && dCp[p].a == dC[p] && dAp[p].b == dA2[p] && dCp[p].b == dC2[p];
                        if (!same) {
                            if (bad < 10) std::printf("P2 lanes differ from the scalar form at te=%g F=%g k=%g\n", te, F, k);
                            bad++;
                        }
                    }
    }
    std::printf("worst value error %.3e, worst Jacobian error %.3e\n", worst_v, worst_j);
    std::printf(bad ? "flex emulation: %d failures\n" : "flex emulation: ok\n", bad);
    return bad ? 1 : 0;
}
