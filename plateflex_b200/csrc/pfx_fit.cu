// pfx_fit.cu -- K5/K6: batched flexural-model evaluation and the per-cell
// bounded non-linear least-squares fit of (Te, F[, alpha]).
//
// Replaces, for a whole grid in one launch,
//   Project.estimate_grid / estimate_cell   classes.py:1031-1123, 1218-1304
//   estimate.L2_estimate_cell               estimate.py:227-409
//   scipy.optimize.curve_fit(bounds=...)    -> least_squares(method='trf')
// The reference's optimiser is SciPy's Trust-Region-Reflective method (bounded
// problem, exact trust-region solver, x_scale=1, ftol=xtol=gtol=1e-8,
// max_nfev=1000), an un-vendored dependency; its published algorithm
// (Branch, Coleman & Li 1999; More 1977 for the trust-region sub-problem) is
// restated here for n<=3 parameters.  With n<=3 everything the method needs
// from the Jacobian is the n x n matrix J^T J and the gradient J^T f,
// accumulated over the wavelet-scale axis; the SVD of the (augmented,
// Coleman-Li scaled) Jacobian becomes a Jacobi eigen-decomposition of an n x n
// matrix.  The Jacobian is analytic (pfx_flex.cuh) where SciPy uses 2-point
// finite differences.
//
// Three kernels share the optimiser:
//   fit_lane_kernel  (default) one lane per cell in flight, one trust-region trial per loop
//                    trip, cells handed out by a global counter
//   fit_cell_kernel  (PLATEFLEX_B200_FIT=static) one cell per thread per grid-stride trip;
//                    also the fallback when the shared-memory cache does not fit
//   fit_kernel       (PLATEFLEX_B200_FIT=group) G lanes per cell, scales strided over the lanes
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pfx_flex.cuh"

namespace pfx {

namespace {

constexpr double TRF_FTOL = 1e-8, TRF_XTOL = 1e-8, TRF_GTOL = 1e-8;  // least_squares defaults
constexpr int TRF_MAX_NFEV = 1000;                                    // estimate.py:279
constexpr double DBL_EPS = 2.220446049250313e-16;

struct FitArgs {
    pfx_flex_conf conf;
    const double *k;
    int ns, nx, ny;
    const double *adm, *eadm, *coh, *ecoh;
    const double *wd, *rhoc_map, *zc_map;
    const uint8_t *mask;
    int nn, atype;
    int fx, fy;  // number of fitted cells along x, y: len(range(0, nx-nn, nn))
    int mx, my;  // output map shape (nx/nn, ny/nn)
    int row0;    // global index of the first grid row (y) held by the maps (row-sharded fits; else 0)
    int cj0;     // global index of the first output row of this shard (else 0)
    double *o_te, *o_te_std, *o_f, *o_f_std, *o_a, *o_a_std, *o_chi2;
    int32_t *status;
    unsigned long long *counter;  // lane-persistent kernel: next cell to hand out (zeroed before the launch)
    int fetch_min;                // ... lanes of a warp wait until this many of them need a new cell
    // Tail control: a lane gives up a cell that has not terminated after trial_cap model evaluations and appends it
    // to todo[] (todo_count, zeroed before the launch; todo_cap = number of cells); the finisher -- the
    // 8-lanes-per-cell kernel, a fraction of the latency per trial -- then fits exactly those cells from the start.
    // A cell that needs hundreds of trials no longer holds the whole grid for milliseconds at the end.
    int *todo = nullptr;
    unsigned long long *todo_count = nullptr;
    int todo_cap = 0, trial_cap = 0;
    int finisher = 0;  // fit_kernel: take the cells from todo[0 .. min(*todo_count, todo_cap))
    const int *dry_flag = nullptr;  // lane kernel: device word, 1 = no water anywhere and one crustal thickness
};

template <int NP>
struct Normal {  // 0.5*sum r^2, J^T r, J^T J
    double cost;
    double g[NP];
    double B[NP][NP];
};

template <int NP>
__device__ __forceinline__ double vnorm(const double (&a)[NP])
{
    double s = 0.;
#pragma unroll
    for (int i = 0; i < NP; i++) s += a[i] * a[i];
    return sqrt(s);
}

// Cyclic Jacobi eigen-decomposition of a symmetric NP x NP matrix; eigenvalues sorted
// descending (the order of singular values), eigenvectors in the columns of V.
template <int NP>
__device__ void jacobi_eig(double (&A)[NP][NP], double (&lam)[NP], double (&V)[NP][NP])
{
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = 0; j < NP; j++) V[i][j] = (i == j) ? 1. : 0.;
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0.;
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int q = p + 1; q < NP; q++) off += fabs(A[p][q]);
        if (off == 0.) break;
#pragma unroll
        for (int p = 0; p < NP; p++) {
#pragma unroll
            for (int q = p + 1; q < NP; q++) {
                const double apq = A[p][q];
                if (apq == 0.) continue;
                if (apq * apq <= (DBL_EPS * 1e-3) * (DBL_EPS * 1e-3) * fabs(A[p][p] * A[q][q])) {
                    A[p][q] = A[q][p] = 0.;  // negligible against the scaled diagonal
                    continue;
                }
                // t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = h / (2 apq), h = Aqq - App, in one square root and
                // one division:  t = 2 apq sgn(h) / (|h| + sqrt(h^2 + 4 apq^2))
                const double h = A[q][q] - A[p][p], a2 = 2. * apq;
                const double t = (h >= 0. ? a2 : -a2) / (fabs(h) + sqrt(h * h + a2 * a2));
                const double c = rsqrt(1. + t * t), s = t * c;
                A[p][p] -= t * apq;
                A[q][q] += t * apq;
                A[p][q] = A[q][p] = 0.;
#pragma unroll
                for (int r = 0; r < NP; r++) {
                    if (r != p && r != q) {
                        const double arp = A[r][p], arq = A[r][q];
                        A[r][p] = A[p][r] = c * arp - s * arq;
                        A[r][q] = A[q][r] = s * arp + c * arq;
                    }
                    const double vrp = V[r][p], vrq = V[r][q];
                    V[r][p] = c * vrp - s * vrq;
                    V[r][q] = s * vrp + c * vrq;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NP; i++) lam[i] = A[i][i];
    // sort descending (NP <= 3: bubble)
#pragma unroll
    for (int a = 0; a < NP - 1; a++)
#pragma unroll
        for (int b = 0; b < NP - 1 - a; b++)
            if (lam[b] < lam[b + 1]) {
                const double tl = lam[b];
                lam[b] = lam[b + 1];
                lam[b + 1] = tl;
#pragma unroll
                for (int r = 0; r < NP; r++) {
                    const double tv = V[r][b];
                    V[r][b] = V[r][b + 1];
                    V[r][b + 1] = tv;
                }
            }
}

// scipy _lsq/common.py: CL_scaling_vector (all bounds finite here).
template <int NP>
__device__ __forceinline__ void cl_scaling(const double (&x)[NP], const double (&g)[NP], const double (&lb)[NP],
                                           const double (&ub)[NP], double (&v)[NP], double (&dv)[NP])
{
#pragma unroll
    for (int i = 0; i < NP; i++) {
        v[i] = 1.;
        dv[i] = 0.;
        if (g[i] < 0.) {
            v[i] = ub[i] - x[i];
            dv[i] = -1.;
        }
        if (g[i] > 0.) {
            v[i] = x[i] - lb[i];
            dv[i] = 1.;
        }
    }
}

// scipy _lsq/common.py: step_size_to_bound.
template <int NP>
__device__ __forceinline__ double step_to_bound(const double (&x)[NP], const double (&s)[NP], const double (&lb)[NP],
                                                const double (&ub)[NP], int (&hits)[NP])
{
    double steps[NP], mn = INFINITY;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        steps[i] = INFINITY;
        if (s[i] != 0.) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
        mn = fmin(mn, steps[i]);
    }
#pragma unroll
    for (int i = 0; i < NP; i++) hits[i] = (steps[i] == mn) ? ((s[i] > 0.) - (s[i] < 0.)) : 0;
    return mn;
}

// 0.5 * s^T Bh s + g_h . s   (evaluate_quadratic with J_h^T J_h + diag folded into Bh)
template <int NP>
__device__ __forceinline__ double quad_form(const double (&Bh)[NP][NP], const double (&a)[NP], const double (&b)[NP])
{
    double q = 0.;
#pragma unroll
    for (int i = 0; i < NP; i++)
#pragma unroll
        for (int j = 0; j < NP; j++) q += a[i] * Bh[i][j] * b[j];
    return q;
}

template <int NP>
__device__ __forceinline__ double vdot(const double (&a)[NP], const double (&b)[NP])
{
    double s = 0.;
#pragma unroll
    for (int i = 0; i < NP; i++) s += a[i] * b[i];
    return s;
}

// scipy _lsq/common.py: minimize_quadratic_1d (candidate order lb, ub, extremum; first minimum wins).
__device__ __forceinline__ void min_quad_1d(double a, double b, double lo, double hi, double c, double &t, double &y)
{
    t = lo;
    y = lo * (a * lo + b) + c;
    const double yh = hi * (a * hi + b) + c;
    if (yh < y) {
        t = hi;
        y = yh;
    }
    if (a != 0.) {
        const double ex = -0.5 * b / a;
        if (lo < ex && ex < hi) {
            const double ye = ex * (a * ex + b) + c;
            if (ye < y) {
                t = ex;
                y = ye;
            }
        }
    }
}

// scipy _lsq/common.py: solve_lsq_trust_region, expressed with lam = s^2 and suf = V^T g_h.
template <int NP>
__device__ void solve_tr(int m, const double (&lam)[NP], const double (&V)[NP][NP], const double (&suf)[NP],
                         double Delta, double &alpha, double (&p)[NP])
{
    // The optimiser's small algebra is a chain of dependent divisions and square roots executed by every lane of a
    // warp at the pace of the slowest: they are kept to the minimum -- the rank test compares squares (the singular
    // values are the square roots of lam), each secular-equation term takes one reciprocal instead of two
    // quotients, and phi/phi' is formed directly.
    const double lam0 = fmax(lam[0], 0.), lamn = fmax(lam[NP - 1], 0.);
    const double rk = DBL_EPS * m;
    const bool full_rank = (m >= NP) && (lamn > rk * rk * lam0);  // s[NP-1] > EPS * m * s[0]
    double w[NP];
    if (full_rank) {
#pragma unroll
        for (int i = 0; i < NP; i++) w[i] = suf[i] / lam[i];
        if (vnorm<NP>(w) <= Delta) {
#pragma unroll
            for (int i = 0; i < NP; i++) {
                double a = 0.;
#pragma unroll
                for (int j = 0; j < NP; j++) a += V[i][j] * w[j];
                p[i] = -a;
            }
            alpha = 0.;
            return;
        }
    }
    // phi(al) = ||p(al)|| - Delta and ratio = phi / phi'  (phi' = -sum suf^2/(lam+al)^3 / ||p||)
    auto phi_ratio = [&](double al, double &phi, double &ratio, double &pn) {
        double pn2 = 0., d3 = 0.;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const double r = 1. / (lam[i] + al);
            const double q = suf[i] * r;
            pn2 += q * q;
            d3 += (q * q) * r;
        }
        pn = sqrt(pn2);
        phi = pn - Delta;
        ratio = -(phi * pn) / d3;
    };
    const double inv_delta = 1. / Delta;
    double alpha_upper = vnorm<NP>(suf) * inv_delta, alpha_lower = 0.;
    if (full_rank) {
        double phi, ratio, pn;
        phi_ratio(0., phi, ratio, pn);
        alpha_lower = -ratio;
    }
    if (!full_rank && alpha == 0.) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    for (int it = 0; it < 10; it++) {
        if (alpha < alpha_lower || alpha > alpha_upper)
            alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        double phi, ratio, pn;
        phi_ratio(alpha, phi, ratio, pn);
        if (phi < 0.) alpha_upper = alpha;
        alpha_lower = fmax(alpha_lower, alpha - ratio);
        alpha -= pn * ratio * inv_delta;  // (phi + Delta) * ratio / Delta
        if (fabs(phi) < 0.01 * Delta) break;
    }
#pragma unroll
    for (int i = 0; i < NP; i++) w[i] = suf[i] / (lam[i] + alpha);
#pragma unroll
    for (int i = 0; i < NP; i++) {
        double a = 0.;
#pragma unroll
        for (int j = 0; j < NP; j++) a += V[i][j] * w[j];
        p[i] = -a;
    }
    const double sc = Delta / vnorm<NP>(p);
#pragma unroll
    for (int i = 0; i < NP; i++) p[i] *= sc;
}

// scipy _lsq/trf.py: select_step.
template <int NP>
__device__ void select_step(const double (&x)[NP], const double (&Bh)[NP][NP], const double (&g_h)[NP],
                            double (&p)[NP], double (&p_h)[NP], const double (&d)[NP], double Delta,
                            const double (&lb)[NP], const double (&ub)[NP], double theta, double (&step)[NP],
                            double (&step_h)[NP], double &predicted)
{
    bool inb = true;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const double xn = x[i] + p[i];
        inb = inb && (xn >= lb[i]) && (xn <= ub[i]);
    }
    if (inb) {
        const double pv = 0.5 * quad_form<NP>(Bh, p_h, p_h) + vdot<NP>(g_h, p_h);
#pragma unroll
        for (int i = 0; i < NP; i++) {
            step[i] = p[i];
            step_h[i] = p_h[i];
        }
        predicted = -pv;
        return;
    }
    int hits[NP], dummy[NP];
    const double p_stride = step_to_bound<NP>(x, p, lb, ub, hits);
    double r_h[NP], r[NP], xb[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        r_h[i] = hits[i] ? -p_h[i] : p_h[i];
        r[i] = d[i] * r_h[i];
        p[i] *= p_stride;
        p_h[i] *= p_stride;
        xb[i] = x[i] + p[i];
    }
    // intersect_trust_region(p_h, r_h, Delta): positive root of ||p_h + t r_h|| = Delta
    double to_tr;
    {
        const double a = vdot<NP>(r_h, r_h), b = vdot<NP>(p_h, r_h), c = vdot<NP>(p_h, p_h) - Delta * Delta;
        const double disc = sqrt(fmax(b * b - a * c, 0.));
        const double q = -(b + copysign(disc, b));
        const double t1 = q / a, t2 = c / q;
        to_tr = fmax(t1, t2);
    }
    const double to_bound = step_to_bound<NP>(xb, r, lb, ub, dummy);
    double r_stride = fmin(to_bound, to_tr), r_lo, r_hi;
    if (r_stride > 0.) {
        r_lo = (1. - theta) * p_stride / r_stride;
        r_hi = (r_stride == to_bound) ? theta * to_bound : to_tr;
    } else {
        r_lo = 0.;
        r_hi = -1.;
    }
    double r_value = INFINITY;
    if (r_lo <= r_hi) {
        const double a = 0.5 * quad_form<NP>(Bh, r_h, r_h);
        const double b = vdot<NP>(g_h, r_h) + quad_form<NP>(Bh, p_h, r_h);
        const double c = 0.5 * quad_form<NP>(Bh, p_h, p_h) + vdot<NP>(g_h, p_h);
        min_quad_1d(a, b, r_lo, r_hi, c, r_stride, r_value);
#pragma unroll
        for (int i = 0; i < NP; i++) {
            r_h[i] = r_h[i] * r_stride + p_h[i];
            r[i] = r_h[i] * d[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        p[i] *= theta;
        p_h[i] *= theta;
    }
    const double p_value = 0.5 * quad_form<NP>(Bh, p_h, p_h) + vdot<NP>(g_h, p_h);
    double ag_h[NP], ag[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        ag_h[i] = -g_h[i];
        ag[i] = d[i] * ag_h[i];
    }
    const double tr2 = Delta / vnorm<NP>(ag_h);
    const double tb2 = step_to_bound<NP>(x, ag, lb, ub, dummy);
    double ag_stride = (tb2 < tr2) ? theta * tb2 : tr2, ag_value;
    {
        const double a = 0.5 * quad_form<NP>(Bh, ag_h, ag_h), b = vdot<NP>(g_h, ag_h);
        min_quad_1d(a, b, 0., ag_stride, 0., ag_stride, ag_value);
    }
    if (p_value < r_value && p_value < ag_value) {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            step[i] = p[i];
            step_h[i] = p_h[i];
        }
        predicted = -p_value;
    } else if (r_value < p_value && r_value < ag_value) {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            step[i] = r[i];
            step_h[i] = r_h[i];
        }
        predicted = -r_value;
    } else {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            step[i] = ag[i] * ag_stride;
            step_h[i] = ag_h[i] * ag_stride;
        }
        predicted = -ag_value;
    }
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
}

// ---- trf_bounds (scipy/optimize/_lsq/trf.py), n = NP <= 3 -----------------------------------
// evaluate(x, normal) fills cost = 0.5*sum r^2, g = J^T r and B = J^T J at x.  Returns SciPy's
// termination status (1 gtol, 2 ftol, 3 xtol, 4 both), -1 when max_nfev was reached, -3 when the
// residuals are not finite at the starting point; x / cur hold the final point.
template <int NP, class Eval>
__device__ __forceinline__ int trf_minimize(const Eval &evaluate, int m, double (&x)[NP], Normal<NP> &cur, int &nfev_out)
{
    // estimate.py:271,280,290,299
    const double lb[3] = {2., 0.0001, 0.0001};
    const double ub[3] = {250., 0.9999, FLEX_PI - 0.001};
    double lbp[NP], ubp[NP];
    const double x0[3] = {20., 0.5, FLEX_PI / 2.};  // estimate.py:271,290
#pragma unroll
    for (int p = 0; p < NP; p++) {
        x[p] = x0[p];
        lbp[p] = lb[p];
        ubp[p] = ub[p];
    }
    Normal<NP> trial;
    evaluate(x, cur);
    int nfev = 1, status = -1;
    double v[NP], dv[NP];
    cl_scaling<NP>(x, cur.g, lbp, ubp, v, dv);
    double Delta;
    {
        double t[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) t[p] = x[p] / sqrt(v[p]);
        Delta = vnorm<NP>(t);
        if (Delta == 0.) Delta = 1.;
    }
    double alpha = 0.;
    if (!isfinite(cur.cost)) status = -3;  // "Residuals are not finite in the initial point"
    while (status == -1) {
        cl_scaling<NP>(x, cur.g, lbp, ubp, v, dv);
        double g_norm = 0.;
#pragma unroll
        for (int p = 0; p < NP; p++) g_norm = fmax(g_norm, fabs(cur.g[p] * v[p]));
        if (g_norm < TRF_GTOL) status = 1;
        if (status != -1 || nfev >= TRF_MAX_NFEV) break;
        double d[NP], diag_h[NP], g_h[NP], Bh[NP][NP], Bw[NP][NP], lam[NP], V[NP][NP], suf[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            d[p] = sqrt(v[p]);
            diag_h[p] = cur.g[p] * dv[p];
            g_h[p] = d[p] * cur.g[p];
        }
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int q = 0; q < NP; q++) {
                Bh[p][q] = d[p] * d[q] * cur.B[p][q] + ((p == q) ? diag_h[p] : 0.);
                Bw[p][q] = Bh[p][q];
            }
        jacobi_eig<NP>(Bw, lam, V);
#pragma unroll
        for (int p = 0; p < NP; p++) {
            double sacc = 0.;
#pragma unroll
            for (int q = 0; q < NP; q++) sacc += V[q][p] * g_h[q];
            suf[p] = sacc;
        }
        const double theta = fmax(0.995, 1. - g_norm);
        double actual = -1., step_norm = 0.;
        double xn[NP];
        while (actual <= 0. && nfev < TRF_MAX_NFEV) {
            double p_h[NP], pp[NP], step[NP], step_h[NP], predicted;
            solve_tr<NP>(m, lam, V, suf, Delta, alpha, p_h);
#pragma unroll
            for (int p = 0; p < NP; p++) pp[p] = d[p] * p_h[p];
            select_step<NP>(x, Bh, g_h, pp, p_h, d, Delta, lbp, ubp, theta, step, step_h, predicted);
#pragma unroll
            for (int p = 0; p < NP; p++) {  // make_strictly_feasible(x + step, rstep=0)
                double t = x[p] + step[p];
                if (t <= lbp[p]) t = nextafter(lbp[p], ubp[p]);
                else if (t >= ubp[p]) t = nextafter(ubp[p], lbp[p]);
                xn[p] = t;
            }
            evaluate(xn, trial);
            nfev++;
            const double step_h_norm = vnorm<NP>(step_h);
            if (!isfinite(trial.cost)) {
                Delta = 0.25 * step_h_norm;
                continue;
            }
            actual = cur.cost - trial.cost;
            // update_tr_radius
            double ratio, Delta_new = Delta;
            if (predicted > 0.) ratio = actual / predicted;
            else if (predicted == 0. && actual == 0.) ratio = 1.;
            else ratio = 0.;
            if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = 2.0 * Delta;
            step_norm = vnorm<NP>(step);
            // check_termination
            const bool f_ok = (actual < TRF_FTOL * cur.cost) && (ratio > 0.25);
            const bool x_ok = step_norm < TRF_XTOL * (TRF_XTOL + vnorm<NP>(x));
            if (f_ok && x_ok) status = 4;
            else if (f_ok) status = 2;
            else if (x_ok) status = 3;
            if (status != -1) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual > 0.) {
#pragma unroll
            for (int p = 0; p < NP; p++) x[p] = xn[p];
            cur = trial;
        }
    }
    nfev_out = nfev;
    return status;
}

// Covariance (curve_fit, _minpack_py.py:1050-1054: pseudo-inverse of J^T J through its singular values,
// absolute_sigma=True), reduced chi-square (estimate.py:286-287) and the output maps of one cell.
template <int NP>
__device__ __forceinline__ void write_cell(const FitArgs &a, size_t dst, int m, const double (&x)[NP],
                                           const Normal<NP> &cur, int status, int nfev)
{
    double Bc[NP][NP], lam[NP], V[NP][NP], sd[NP];
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
        for (int q = 0; q < NP; q++) Bc[p][q] = cur.B[p][q];
    jacobi_eig<NP>(Bc, lam, V);
    const double s0 = sqrt(fmax(lam[0], 0.));
    const double thr = DBL_EPS * (double)((m > NP) ? m : NP) * s0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        double c = 0.;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const double sq = sqrt(fmax(lam[q], 0.));
            if (sq > thr) c += V[p][q] * V[p][q] / lam[q];
        }
        sd[p] = sqrt(c);
    }
    a.o_te[dst] = x[0];
    a.o_te_std[dst] = sd[0];
    a.o_f[dst] = x[1];
    a.o_f_std[dst] = sd[1];
    if (NP == 3) {
        a.o_a[dst] = x[NP - 1];
        a.o_a_std[dst] = sd[NP - 1];
    }
    a.o_chi2[dst] = 2. * cur.cost / (double)(m - NP);
    // low byte: 1 converged, 2 iteration limit, 3 bad data; upper bits: model evaluations used (SciPy's nfev)
    if (a.status) a.status[dst] = ((status > 0) ? 1 : ((status == -3) ? 3 : 2)) | (nfev << 8);
}

__device__ __forceinline__ void write_bad_cell(const FitArgs &a, size_t dst, bool with_alpha)
{
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    a.o_te[dst] = a.o_te_std[dst] = a.o_f[dst] = a.o_f_std[dst] = a.o_chi2[dst] = nanv;
    if (with_alpha) a.o_a[dst] = a.o_a_std[dst] = nanv;
    if (a.status) a.status[dst] = 3;
}

// accumulate one residual row (r, jr[NP]) into acc = [sum r^2, J^T r (NP), upper triangle of J^T J]
template <int NP>
__device__ __forceinline__ void accumulate(double (&acc)[1 + NP + NP * (NP + 1) / 2], double r, const double (&jr)[NP])
{
    acc[0] += r * r;
    int q = 1 + NP;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        acc[1 + p] += jr[p] * r;
#pragma unroll
        for (int p2 = p; p2 < NP; p2++) acc[q++] += jr[p] * jr[p2];
    }
}

template <int NP>
__device__ __forceinline__ void unpack(const double (&acc)[1 + NP + NP * (NP + 1) / 2], Normal<NP> &nm)
{
    nm.cost = 0.5 * acc[0];
    int q = 1 + NP;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        nm.g[p] = acc[1 + p];
#pragma unroll
        for (int p2 = p; p2 < NP; p2++) {
            nm.B[p][p2] = acc[q];
            nm.B[p2][p] = acc[q];
            q++;
        }
    }
}

__device__ __forceinline__ double rcp_nr(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// ---- thread-per-cell fit kernel ----------------------------------------------------------------
// One thread per grid cell, consecutive threads on consecutive x so that every load of the
// (nx, ny, ns) maps is a coalesced 256-byte warp access; the scale axis is walked serially and
// the maps are re-read (from L1/L2) at every model evaluation, only the depth attenuation factors
// exp(-k*(zc+wd)), exp(-k*wd) are kept per thread in shared memory.  No lane ever idles on the
// small dense algebra of the optimiser, which made this 3x cheaper than one warp per cell.
constexpr int FIT_NT = 128;
constexpr int FIT_NT_DRY = 64, FIT_CTAS_DRY = 5;  // the dry-land form of the lane kernel (see fit_lane_kernel)
#ifndef FIT_PAIR
#define FIT_PAIR 1  // lane kernel: two scales side by side in the model loop (P2 of pfx_flex.cuh)
#endif
#ifndef FIT_UNROLL
#define FIT_UNROLL (FIT_PAIR ? 1 : 2)  // unrolling of the model loop over scales (pairs of scales with FIT_PAIR)
#endif
constexpr int FIT_UNROLL_N = FIT_UNROLL;
#ifndef FIT_MIN_BLOCKS
#define FIT_MIN_BLOCKS 3
#endif

// CACHE: the cell's observations (value and 1/sigma per scale) are copied to shared memory once and every model
// evaluation reads them from there instead of from L2 (4 x ns more doubles per thread, two CTAs per SM).
template <int NP, bool CACHE>
__global__ void __launch_bounds__(FIT_NT, CACHE ? 2 : FIT_MIN_BLOCKS) fit_cell_kernel(FitArgs a)
{
    extern __shared__ double fsh[];
    // Depth attenuation per (cell, scale), one array per thread: the deep term 2 pi G drho e^{-k(zc+wd)} for
    // Bouguer data (the top term vanishes); for free air with one crustal thickness e^{-k wd}, from which both
    // terms follow with the per-scale table e^{-k zc}; for free air with a zc grid both terms (two arrays).
    const bool freeair = a.conf.boug != 1;
    const bool two = freeair && a.zc_map != nullptr;  // both terms stored
    const bool shared_zc = freeair && !two;           // e^{-k wd} stored
    double *k4 = fsh;                              // [ns] k^4 * Em*1e9/12/(1-nu^2)
    double *ezc = fsh + a.ns;                      // [ns] e^{-k zc}  (uniform zc)
    double *sdeep = fsh + 2 * a.ns;                // [ns][FIT_NT]  e^{-k wd}, or the deep term with a zc grid
    double *stop = sdeep + (size_t)a.ns * FIT_NT;  // [ns][FIT_NT]  (zc grid and free air only, else absent)
    double *sobs = stop + (two ? (size_t)a.ns * FIT_NT : 0);  // CACHE: [4][ns][FIT_NT]
    for (int s = threadIdx.x; s < a.ns; s += FIT_NT) {
        const double k2 = a.k[s] * a.k[s];
        k4[s] = (k2 * k2) * (FLEX_EM * 1.e9 / 12. / (1. - FLEX_NU * FLEX_NU));
        ezc[s] = exp(-a.k[s] * a.conf.zc);
    }
    __syncthreads();
    const long long ncells = (long long)a.fx * a.fy;
    const size_t nxy = (size_t)a.nx * a.ny;
    const bool use_adm = a.atype != PFX_ATYPE_COH, use_coh = a.atype != PFX_ATYPE_ADMIT;
    const int m = (use_adm ? a.ns : 0) + (use_coh ? a.ns : 0);
    const int tid = threadIdx.x;
    for (long long cell = (long long)blockIdx.x * FIT_NT + tid; cell < ncells; cell += (long long)gridDim.x * FIT_NT) {
        const int ci = (int)(cell % a.fx), cj = (int)(cell / a.fx);
        const size_t src = (size_t)((a.cj0 + cj) * a.nn - a.row0) * a.nx + (size_t)ci * a.nn;
        const size_t dst = (size_t)cj * a.mx + ci;
        if (a.mask && a.mask[src]) continue;  // classes.py:1228-1231 (outputs stay zero)
        const double rhoc = a.rhoc_map ? a.rhoc_map[src] : a.conf.rhoc;  // classes.py:1087-1091
        const double zc = a.zc_map ? a.zc_map[src] : a.conf.zc;
        const FlexCell fc = flex_cell(a.conf, rhoc, zc, a.wd[src]);
        bool bad = false;
        for (int s = 0; s < a.ns; s++) {
            if (shared_zc) {
                sdeep[s * FIT_NT + tid] = (fc.wd == 0.) ? 1. : exp(-a.k[s] * fc.wd);
            } else {
                const FlexK q = flex_k(fc, a.k[s]);
                sdeep[s * FIT_NT + tid] = q.deep;
                if (two) stop[s * FIT_NT + tid] = q.top;
            }
            const size_t o = src + nxy * s;
            if (use_adm) {
                const double y = __ldg(a.adm + o), e = __ldg(a.eadm + o);
                bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                if (CACHE) {
                    sobs[(0 * a.ns + s) * FIT_NT + tid] = y;
                    sobs[(1 * a.ns + s) * FIT_NT + tid] = rcp_nr(e);
                }
            }
            if (use_coh) {
                const double y = __ldg(a.coh + o), e = __ldg(a.ecoh + o);
                bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                if (CACHE) {
                    sobs[(2 * a.ns + s) * FIT_NT + tid] = y;
                    sobs[(3 * a.ns + s) * FIT_NT + tid] = rcp_nr(e);
                }
            }
        }
        if (bad) {  // SciPy raises ValueError here; we flag the cell
            write_bad_cell(a, dst, NP == 3);
            continue;
        }
        const double cdeep = 2. * FLEX_PI * FLEX_GC * fc.drho, ctop = 2. * FLEX_PI * FLEX_GC * (fc.A * fc.rcf);
        auto evaluate = [&](const double(&x)[NP], Normal<NP> &nm) {
            const double te = x[0], F = x[1];
            const double alpha = (NP == 3) ? x[NP - 1] : FLEX_PI / 2.;  // estimate.py:292
            double sa, ca;
            sincos(alpha, &sa, &ca);
            const double te3 = te * te * te, rf = 1. / (1. - F), ff = F * rf, dff = rf * rf;
            double acc[1 + NP + NP * (NP + 1) / 2];
#pragma unroll
            for (int q = 0; q < 1 + NP + NP * (NP + 1) / 2; q++) acc[q] = 0.;
#pragma unroll FIT_UNROLL_N
            for (int s = 0; s < a.ns; s++) {
                FlexK fk;
                fk.k4d = k4[s];
                fk.deep = sdeep[s * FIT_NT + tid];
                fk.top = 0.;
                if (freeair) {
                    if (two) {
                        fk.top = stop[s * FIT_NT + tid];
                    } else {
                        fk.top = ctop * fk.deep;
                        fk.deep = cdeep * (ezc[s] * fk.deep);
                    }
                }
                double adm, coh, dA[3], dC[3];
                flex_eval<true>(fc, fk, te, te3, ff, dff, ca, sa, adm, coh, dA, dC);
                const size_t o = src + nxy * s;
                if (use_adm) {
                    // curve_fit: transform = 1/sigma
                    const double w = CACHE ? sobs[(1 * a.ns + s) * FIT_NT + tid] : rcp_nr(__ldg(a.eadm + o));
                    const double r = (adm - (CACHE ? sobs[(0 * a.ns + s) * FIT_NT + tid] : __ldg(a.adm + o))) * w;
                    double jr[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++) jr[p] = dA[p] * w;
                    accumulate<NP>(acc, r, jr);
                }
                if (use_coh) {
                    const double w = CACHE ? sobs[(3 * a.ns + s) * FIT_NT + tid] : rcp_nr(__ldg(a.ecoh + o));
                    const double r = (coh - (CACHE ? sobs[(2 * a.ns + s) * FIT_NT + tid] : __ldg(a.coh + o))) * w;
                    double jr[NP];
#pragma unroll
                    for (int p = 0; p < NP; p++) jr[p] = dC[p] * w;
                    accumulate<NP>(acc, r, jr);
                }
            }
            unpack<NP>(acc, nm);
        };
        double x[NP];
        Normal<NP> cur;
        int nfev = 0;
        const int status = trf_minimize<NP>(evaluate, m, x, cur, nfev);
        write_cell<NP>(a, dst, m, x, cur, status, nfev);
    }
}

// ---- lane-persistent fit kernel (default) --------------------------------------------------------
// The number of model evaluations differs a lot between cells (mean 10, maximum several hundred at the
// BASELINE sizes).  With one cell per thread per loop trip a warp advances at the pace of its slowest lane
// and the grid at the pace of its unluckiest thread.  Here every lane keeps its own optimiser state and the
// loop body is ONE trust-region trial (scaling, eigen-decomposition, step, model evaluation, radius update) of
// whatever cell the lane currently holds; a lane whose cell has terminated writes it out and takes the next
// cell from a global counter (one atomic per warp and trip).  The arithmetic of a cell is the sequence of
// trf_minimize exactly -- the outer-iteration quantities are recomputed from (x, J^T J, J^T r) on a retry
// instead of being kept, which gives the same values -- so results do not depend on the schedule.
template <int NP, bool DRY>
__global__ void __launch_bounds__(DRY ? FIT_NT_DRY : FIT_NT, DRY ? FIT_CTAS_DRY : 2) fit_lane_kernel(FitArgs a)
{
    // DRY: no cell of this launch lies under water and the crustal thickness is one number, so the depth factors of
    // a scale are the same for every cell, e^{-k zc} (the table ezc), and no per-lane array of them is kept: 640
    // instead of 800 bytes of shared memory per lane, ten warps per SM instead of eight -- the kernel is bound by
    // the latency of the optimiser's dependent chains, so resident warps are what it needs.  a.dry_flag (device, set
    // by wet_kernel before the launch) says which of the two variants of this launch does the work.
    constexpr int LNT = DRY ? FIT_NT_DRY : FIT_NT;
    if ((*a.dry_flag != 0) != DRY) return;
    extern __shared__ double fsh[];
    const bool freeair = a.conf.boug != 1;
    const bool two = freeair && a.zc_map != nullptr;  // both attenuation terms stored
    const bool shared_zc = freeair && !two;           // e^{-k wd} stored
    double *k4 = fsh;                              // [ns] k^4 * Em*1e9/12/(1-nu^2)
    double *ezc = fsh + a.ns;                      // [ns] e^{-k zc}  (uniform zc)
    double *sdeep = fsh + 2 * a.ns;                // [ns][LNT]  (absent when DRY)
    double *stop = sdeep + (DRY ? 0 : (size_t)a.ns * LNT);  // [ns][LNT]  (zc grid and free air only, else absent)
    double *sobs = stop + (two ? (size_t)a.ns * LNT : 0);  // [4][ns][LNT] = adm, 1/eadm, coh, 1/ecoh
    for (int s = threadIdx.x; s < a.ns; s += LNT) {
        const double k2 = a.k[s] * a.k[s];
        k4[s] = (k2 * k2) * (FLEX_EM * 1.e9 / 12. / (1. - FLEX_NU * FLEX_NU));
        ezc[s] = exp(-a.k[s] * a.conf.zc);
    }
    __syncthreads();
    const long long ncells = (long long)a.fx * a.fy;
    const size_t nxy = (size_t)a.nx * a.ny;
    const bool use_adm = a.atype != PFX_ATYPE_COH, use_coh = a.atype != PFX_ATYPE_ADMIT;
    const int m = (use_adm ? a.ns : 0) + (use_coh ? a.ns : 0);
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    // estimate.py:271,280,290,299
    const double lb3[3] = {2., 0.0001, 0.0001}, ub3[3] = {250., 0.9999, FLEX_PI - 0.001};
    const double x03[3] = {20., 0.5, FLEX_PI / 2.};
    double lbp[NP], ubp[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        lbp[p] = lb3[p];
        ubp[p] = ub3[p];
    }

    // ---- per-lane state of the cell in flight
    bool have = false, more = true, init = false;
    size_t dst = 0;
    int mycell = 0;
    FlexCell fc = flex_cell(a.conf, a.conf.rhoc, a.conf.zc, 0.);
    double cdeep = 0., ctop = 0.;
    double x[NP];
    Normal<NP> cur;
    double Delta = 0., alpha = 0.;
    int nfev = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) x[p] = x03[p];

    auto evaluate = [&](const double(&xe)[NP], Normal<NP> &nm) {
        const double te = xe[0], F = xe[1];
        const double al = (NP == 3) ? xe[NP - 1] : FLEX_PI / 2.;  // estimate.py:292
        double sa, ca;
        sincos(al, &sa, &ca);
        const double te3 = te * te * te, rf = 1. / (1. - F), ff = F * rf, dff = rf * rf;
        double acc[1 + NP + NP * (NP + 1) / 2];
#pragma unroll
        for (int q = 0; q < 1 + NP + NP * (NP + 1) / 2; q++) acc[q] = 0.;
        // the depth factors of scale s (see the fetch below for the three storage forms)
        auto depth = [&](int s, double &top, double &deep) {
            if (DRY) {  // the values the general form stores for wd = 0, in the same order of operations
                if (freeair) {
                    top = ctop * 1.;
                    deep = cdeep * (ezc[s] * 1.);
                } else {
                    top = 0.;
                    deep = 2. * FLEX_PI * FLEX_GC * (fc.drho * ezc[s]);  // flex_k with wd = 0, zc = conf.zc
                }
                return;
            }
            deep = sdeep[s * LNT + tid];
            top = 0.;
            if (freeair) {
                if (two) {
                    top = stop[s * LNT + tid];
                } else {
                    top = ctop * deep;
                    deep = cdeep * (ezc[s] * deep);
                }
            }
        };
        // one residual row per observable of scale s, in the order admittance, coherence
        auto rows = [&](int s, double adm, double coh, const double(&dA)[3], const double(&dC)[3]) {
            if (use_adm) {
                const double w = sobs[(1 * a.ns + s) * LNT + tid];  // curve_fit: transform = 1/sigma
                const double r = (adm - sobs[(0 * a.ns + s) * LNT + tid]) * w;
                double jr[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) jr[p] = dA[p] * w;
                accumulate<NP>(acc, r, jr);
            }
            if (use_coh) {
                const double w = sobs[(3 * a.ns + s) * LNT + tid];
                const double r = (coh - sobs[(2 * a.ns + s) * LNT + tid]) * w;
                double jr[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) jr[p] = dC[p] * w;
                accumulate<NP>(acc, r, jr);
            }
        };
#if FIT_PAIR
        // two scales side by side (pfx_flex.cuh, P2): independent chains for the FP64 pipe; rows accumulated in
        // scale order, so the sums are those of the one-scale loop
        int s = 0;
#pragma unroll FIT_UNROLL_N
        for (; s + 1 < a.ns; s += 2) {
            P2 k4d = {k4[s], k4[s + 1]}, top, deep;
            depth(s, top.a, deep.a);
            depth(s + 1, top.b, deep.b);
            P2 adm, coh, dA[3], dC[3];
            flex_eval_g<true, P2>(fc, k4d, top, deep, te, te3, ff, dff, ca, sa, adm, coh, dA, dC);
            const double dAa[3] = {dA[0].a, dA[1].a, dA[2].a}, dCa[3] = {dC[0].a, dC[1].a, dC[2].a};
            const double dAb[3] = {dA[0].b, dA[1].b, dA[2].b}, dCb[3] = {dC[0].b, dC[1].b, dC[2].b};
            rows(s, adm.a, coh.a, dAa, dCa);
            rows(s + 1, adm.b, coh.b, dAb, dCb);
        }
        if (s < a.ns) {
            double top, deep, adm, coh, dA[3], dC[3];
            depth(s, top, deep);
            flex_eval_g<true, double>(fc, k4[s], top, deep, te, te3, ff, dff, ca, sa, adm, coh, dA, dC);
            rows(s, adm, coh, dA, dC);
        }
#else
#pragma unroll FIT_UNROLL_N
        for (int s = 0; s < a.ns; s++) {
            double top, deep, adm, coh, dA[3], dC[3];
            depth(s, top, deep);
            flex_eval_g<true, double>(fc, k4[s], top, deep, te, te3, ff, dff, ca, sa, adm, coh, dA, dC);
            rows(s, adm, coh, dA, dC);
        }
#endif
        unpack<NP>(acc, nm);
    };

    for (;;) {
        // ---- lanes without a cell take the next one (warp-aggregated atomic: consecutive cells for the lanes
        // that fetch together, so their loads share sectors)
        // (a fetch stalls the whole warp for the latency of ~4 ns loads, so idle lanes wait for company unless
        // nothing else is running in the warp)
        const unsigned want = __ballot_sync(0xffffffffu, !have && more), busy = __ballot_sync(0xffffffffu, have);
        if (!have && more && (__popc(want) >= a.fetch_min || busy == 0u)) {
            const unsigned need = __activemask();
            const int leader = __ffs(need) - 1, rank = __popc(need & ((1u << lane) - 1u));
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(need, base, leader);
            const long long cell = (long long)base + rank;
            if (cell >= ncells) {
                more = false;
            } else {
                const int ci = (int)(cell % a.fx), cj = (int)(cell / a.fx);
                const size_t src = (size_t)((a.cj0 + cj) * a.nn - a.row0) * a.nx + (size_t)ci * a.nn;
                dst = (size_t)cj * a.mx + ci;
                mycell = (int)cell;
                if (!(a.mask && a.mask[src])) {  // classes.py:1228-1231 (outputs of masked cells stay zero)
                    const double rhoc = a.rhoc_map ? a.rhoc_map[src] : a.conf.rhoc;  // classes.py:1087-1091
                    const double zc = a.zc_map ? a.zc_map[src] : a.conf.zc;
                    fc = flex_cell(a.conf, rhoc, zc, a.wd[src]);
                    cdeep = 2. * FLEX_PI * FLEX_GC * fc.drho;
                    ctop = 2. * FLEX_PI * FLEX_GC * (fc.A * fc.rcf);
                    bool bad = false;
                    for (int s = 0; s < a.ns; s++) {
                        if (DRY) {
                        } else if (shared_zc) {
                            sdeep[s * LNT + tid] = (fc.wd == 0.) ? 1. : exp(-a.k[s] * fc.wd);
                        } else {
                            const FlexK q = flex_k(fc, a.k[s]);
                            sdeep[s * LNT + tid] = q.deep;
                            if (two) stop[s * LNT + tid] = q.top;
                        }
                        const size_t o = src + nxy * s;
                        if (use_adm) {
                            const double y = __ldg(a.adm + o), e = __ldg(a.eadm + o);
                            bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                            sobs[(0 * a.ns + s) * LNT + tid] = y;
                            sobs[(1 * a.ns + s) * LNT + tid] = rcp_nr(e);
                        }
                        if (use_coh) {
                            const double y = __ldg(a.coh + o), e = __ldg(a.ecoh + o);
                            bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                            sobs[(2 * a.ns + s) * LNT + tid] = y;
                            sobs[(3 * a.ns + s) * LNT + tid] = rcp_nr(e);
                        }
                    }
                    if (bad) {  // SciPy raises ValueError here; we flag the cell
                        write_bad_cell(a, dst, NP == 3);
                    } else {
                        have = true;
                        init = true;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, have || more)) break;

        // ---- one trust-region trial of the lane's cell (trf_minimize, flattened)
        bool act = have;
        double xn[NP], step[NP], step_h[NP], predicted = 0.;
#pragma unroll
        for (int p = 0; p < NP; p++) {
            xn[p] = x03[p];  // the starting point when the cell is new
            step[p] = step_h[p] = 0.;
        }
        if (act && !init) {
            double v[NP], dv[NP];
            cl_scaling<NP>(x, cur.g, lbp, ubp, v, dv);
            double g_norm = 0.;
#pragma unroll
            for (int p = 0; p < NP; p++) g_norm = fmax(g_norm, fabs(cur.g[p] * v[p]));
            int status = -1;
            if (g_norm < TRF_GTOL) status = 1;
            bool handed_over = false;
            // A cell that has not terminated after trial_cap model evaluations goes to the finisher -- always, whatever
            // the state of the work counter, so that WHICH cells the finisher fits (its sums run in another order, the
            // results differ in the last digits) depends on the cell alone: the maps are reproducible from run to run
            // and identical between a row-sharded fit and the fit of the whole grid.  The list holds every cell.
            if (status == -1 && a.todo && nfev == a.trial_cap) {
                const unsigned long long slot = atomicAdd(a.todo_count, 1ull);
                if (slot < (unsigned long long)a.todo_cap) {
                    a.todo[slot] = mycell;
                    handed_over = true;
                }
            }
            if (handed_over) {
                have = false;
                act = false;
            } else if (status != -1 || nfev >= TRF_MAX_NFEV) {
                write_cell<NP>(a, dst, m, x, cur, status, nfev);
                have = false;
                act = false;
            } else {
                double d[NP], diag_h[NP], g_h[NP], Bh[NP][NP], Bw[NP][NP], lam[NP], V[NP][NP], suf[NP];
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    d[p] = sqrt(v[p]);
                    diag_h[p] = cur.g[p] * dv[p];
                    g_h[p] = d[p] * cur.g[p];
                }
#pragma unroll
                for (int p = 0; p < NP; p++)
#pragma unroll
                    for (int q = 0; q < NP; q++) {
                        Bh[p][q] = d[p] * d[q] * cur.B[p][q] + ((p == q) ? diag_h[p] : 0.);
                        Bw[p][q] = Bh[p][q];
                    }
                jacobi_eig<NP>(Bw, lam, V);
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    double sacc = 0.;
#pragma unroll
                    for (int q = 0; q < NP; q++) sacc += V[q][p] * g_h[q];
                    suf[p] = sacc;
                }
                const double theta = fmax(0.995, 1. - g_norm);
                double p_h[NP], pp[NP];
                solve_tr<NP>(m, lam, V, suf, Delta, alpha, p_h);
#pragma unroll
                for (int p = 0; p < NP; p++) pp[p] = d[p] * p_h[p];
                select_step<NP>(x, Bh, g_h, pp, p_h, d, Delta, lbp, ubp, theta, step, step_h, predicted);
#pragma unroll
                for (int p = 0; p < NP; p++) {  // make_strictly_feasible(x + step, rstep=0)
                    double t = x[p] + step[p];
                    if (t <= lbp[p]) t = nextafter(lbp[p], ubp[p]);
                    else if (t >= ubp[p]) t = nextafter(ubp[p], lbp[p]);
                    xn[p] = t;
                }
            }
        }
        Normal<NP> trial;
        if (act) evaluate(xn, trial);
        if (act) {
            if (init) {
                init = false;
#pragma unroll
                for (int p = 0; p < NP; p++) x[p] = x03[p];
                cur = trial;
                nfev = 1;
                alpha = 0.;
                double v[NP], dv[NP], t[NP];
                cl_scaling<NP>(x, cur.g, lbp, ubp, v, dv);
#pragma unroll
                for (int p = 0; p < NP; p++) t[p] = x[p] / sqrt(v[p]);
                Delta = vnorm<NP>(t);
                if (Delta == 0.) Delta = 1.;
                if (!isfinite(cur.cost)) {  // "Residuals are not finite in the initial point"
                    write_cell<NP>(a, dst, m, x, cur, -3, nfev);
                    have = false;
                }
            } else {
                nfev++;
                const double step_h_norm = vnorm<NP>(step_h);
                if (!isfinite(trial.cost)) {
                    Delta = 0.25 * step_h_norm;
                } else {
                    const double actual = cur.cost - trial.cost;
                    // update_tr_radius
                    double ratio, Delta_new = Delta;
                    if (predicted > 0.) ratio = actual / predicted;
                    else if (predicted == 0. && actual == 0.) ratio = 1.;
                    else ratio = 0.;
                    if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
                    else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = 2.0 * Delta;
                    const double step_norm = vnorm<NP>(step);
                    // check_termination
                    const bool f_ok = (actual < TRF_FTOL * cur.cost) && (ratio > 0.25);
                    const bool x_ok = step_norm < TRF_XTOL * (TRF_XTOL + vnorm<NP>(x));
                    int status = -1;
                    if (f_ok && x_ok) status = 4;
                    else if (f_ok) status = 2;
                    else if (x_ok) status = 3;
                    if (status == -1) {
                        alpha *= Delta / Delta_new;
                        Delta = Delta_new;
                    }
                    if (actual > 0.) {
#pragma unroll
                        for (int p = 0; p < NP; p++) x[p] = xn[p];
                        cur = trial;
                    }
                    if (status != -1) {
                        write_cell<NP>(a, dst, m, x, cur, status, nfev);
                        have = false;
                    }
                }
            }
        }
    }
}

// dry-land flag of a launch (see fit_lane_kernel<NP, DRY>): starts at 1, cleared by any fitted cell under water
__global__ void wet_kernel(FitArgs a, int *flag)
{
    const long long ncells = (long long)a.fx * a.fy;
    bool wet = false;
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < ncells; cell += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(cell % a.fx), cj = (int)(cell / a.fx);
        const size_t src = (size_t)((a.cj0 + cj) * a.nn - a.row0) * a.nx + (size_t)ci * a.nn;
        wet = wet || !(a.wd[src] == 0.);
    }
    if (wet) *flag = 0;
}

// ---- the fit kernel ---------------------------------------------------------------
// G lanes per cell, NSL scales per lane (ns <= G*NSL), NP fitted parameters.
template <int G, int NSL, int NP>
__global__ void __launch_bounds__(128) fit_kernel(FitArgs a)
{
    const int lane = threadIdx.x & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    long long ncells = (long long)a.fx * a.fy;
    if (a.finisher) {
        const unsigned long long n = *a.todo_count;
        ncells = (long long)(n < (unsigned long long)a.todo_cap ? n : (unsigned long long)a.todo_cap);
    }
    const size_t nxy = (size_t)a.nx * a.ny;
    const bool use_adm = a.atype != PFX_ATYPE_COH, use_coh = a.atype != PFX_ATYPE_ADMIT;
    const int m = (use_adm ? a.ns : 0) + (use_coh ? a.ns : 0);

    for (long long item = group; item < ncells; item += ngroups) {
        const long long cell = a.finisher ? (long long)a.todo[item] : item;
        const int ci = (int)(cell % a.fx), cj = (int)(cell / a.fx);
        const int i = ci * a.nn, j = (a.cj0 + cj) * a.nn - a.row0;
        const size_t src = (size_t)j * a.nx + i;
        const size_t dst = (size_t)cj * a.mx + ci;
        if (a.mask && a.mask[src]) continue;  // classes.py:1228-1231 (outputs stay zero)

        const double rhoc = a.rhoc_map ? a.rhoc_map[src] : a.conf.rhoc;  // classes.py:1087-1091
        const double zc = a.zc_map ? a.zc_map[src] : a.conf.zc;
        const FlexCell fc = flex_cell(a.conf, rhoc, zc, a.wd[src]);

        // per-lane data: scales lane, lane+G, ...
        FlexK fk[NSL];
        double ya[NSL], wa[NSL], yc[NSL], wc[NSL];
        bool bad = false;
#pragma unroll
        for (int t = 0; t < NSL; t++) {
            const int s = lane + t * G;
            ya[t] = wa[t] = yc[t] = wc[t] = 0.;
            fk[t].k4d = fk[t].top = fk[t].deep = 0.;
            if (s < a.ns) {
                fk[t] = flex_k(fc, a.k[s]);
                const size_t o = src + nxy * s;
                if (use_adm) {
                    const double y = a.adm[o], e = a.eadm[o];
                    bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                    ya[t] = y;
                    wa[t] = 1. / e;  // curve_fit: transform = 1/sigma
                }
                if (use_coh) {
                    const double y = a.coh[o], e = a.ecoh[o];
                    bad = bad || !isfinite(y) || !isfinite(e) || e == 0.;
                    yc[t] = y;
                    wc[t] = 1. / e;
                }
            }
        }
        bad = __any_sync(gmask, bad);
        if (bad) {  // SciPy raises ValueError here; we flag the cell
            if (lane == 0) {
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                a.o_te[dst] = a.o_te_std[dst] = a.o_f[dst] = a.o_f_std[dst] = a.o_chi2[dst] = nanv;
                if (NP == 3) a.o_a[dst] = a.o_a_std[dst] = nanv;
                if (a.status) a.status[dst] = 3;
            }
            continue;
        }

        // residuals r = (model - y)/sigma; accumulate cost, J^T r, J^T J over the group's scales
        auto evaluate = [&](const double(&x)[NP], Normal<NP> &nm) {
            const double te = x[0], F = x[1];
            const double alpha = (NP == 3) ? x[NP - 1] : FLEX_PI / 2.;  // estimate.py:292
            double sa, ca;
            sincos(alpha, &sa, &ca);
            const double te3 = te * te * te, rf = 1. / (1. - F), ff = F * rf, dff = rf * rf;
            double acc[1 + NP + NP * (NP + 1) / 2];
#pragma unroll
            for (int q = 0; q < 1 + NP + NP * (NP + 1) / 2; q++) acc[q] = 0.;
#pragma unroll
            for (int t = 0; t < NSL; t++) {
                if (lane + t * G < a.ns) {
                    double adm, coh, dA[3], dC[3];
                    flex_eval<true>(fc, fk[t], te, te3, ff, dff, ca, sa, adm, coh, dA, dC);
                    if (use_adm) {
                        const double r = (adm - ya[t]) * wa[t];
                        double jr[NP];
#pragma unroll
                        for (int p = 0; p < NP; p++) jr[p] = dA[p] * wa[t];
                        acc[0] += r * r;
                        int q = 1 + NP;
#pragma unroll
                        for (int p = 0; p < NP; p++) {
                            acc[1 + p] += jr[p] * r;
#pragma unroll
                            for (int p2 = p; p2 < NP; p2++) acc[q++] += jr[p] * jr[p2];
                        }
                    }
                    if (use_coh) {
                        const double r = (coh - yc[t]) * wc[t];
                        double jr[NP];
#pragma unroll
                        for (int p = 0; p < NP; p++) jr[p] = dC[p] * wc[t];
                        acc[0] += r * r;
                        int q = 1 + NP;
#pragma unroll
                        for (int p = 0; p < NP; p++) {
                            acc[1 + p] += jr[p] * r;
#pragma unroll
                            for (int p2 = p; p2 < NP; p2++) acc[q++] += jr[p] * jr[p2];
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 1 + NP + NP * (NP + 1) / 2; q++) acc[q] = group_sum<G>(acc[q], gmask);
            nm.cost = 0.5 * acc[0];
            int q = 1 + NP;
#pragma unroll
            for (int p = 0; p < NP; p++) {
                nm.g[p] = acc[1 + p];
#pragma unroll
                for (int p2 = p; p2 < NP; p2++) {
                    nm.B[p][p2] = acc[q];
                    nm.B[p2][p] = acc[q];
                    q++;
                }
            }
        };

        double x[NP];
        Normal<NP> cur;
        int nfev = 0;
        const int status = trf_minimize<NP>(evaluate, m, x, cur, nfev);

        if (lane == 0) write_cell<NP>(a, dst, m, x, cur, status, nfev);
    }
}

// Batched model curves (flex.real_xspec_functions for many parameter sets).
__global__ void xspec_kernel(pfx_flex_conf conf, const double *__restrict__ k, int ns, int ncurves,
                             const double *__restrict__ te, const double *__restrict__ f,
                             const double *__restrict__ alpha, const double *__restrict__ wd,
                             double *__restrict__ admit, double *__restrict__ coh, double *__restrict__ jac_admit,
                             double *__restrict__ jac_coh)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ncurves * ns) return;
    const int c = (int)(t / ns), s = (int)(t % ns);
    const FlexCell fc = flex_cell(conf, conf.rhoc, conf.zc, wd ? wd[c] : 0.);
    const FlexK fk = flex_k(fc, k[s]);
    const double T = te[c], F = f[c];
    double sa, ca;
    sincos(alpha[c], &sa, &ca);
    double a, co, dA[3], dC[3];
    flex_eval<true>(fc, fk, T, T * T * T, F / (1. - F), 1. / ((1. - F) * (1. - F)), ca, sa, a, co, dA, dC);
    admit[t] = a;
    coh[t] = co;
    if (jac_admit)
        for (int p = 0; p < 3; p++) jac_admit[((size_t)c * 3 + p) * ns + s] = dA[p];
    if (jac_coh)
        for (int p = 0; p < 3; p++) jac_coh[((size_t)c * 3 + p) * ns + s] = dC[p];
}

template <int G, int NSL>
int launch_fit_np(const FitArgs &a, int alph, cudaStream_t stream, int nsm)
{
    const long long ncells = a.finisher ? (long long)a.todo_cap : (long long)a.fx * a.fy;
    const int threads = 128;
    long long blocks = (ncells * G + threads - 1) / threads;
    const long long cap = (long long)nsm * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (alph)
        fit_kernel<G, NSL, 3><<<(unsigned)blocks, threads, 0, stream>>>(a);
    else
        fit_kernel<G, NSL, 2><<<(unsigned)blocks, threads, 0, stream>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

void launch_wet_kernel(const FitArgs &a, int *flag, int blocks, cudaStream_t st) { wet_kernel<<<blocks, 256, 0, st>>>(a, flag); }

int launch_fit_cell(const FitArgs &a, int alph, cudaStream_t stream, int nsm)
{
    const long long ncells = (long long)a.fx * a.fy;
    const size_t smem0 =
        (2 * (size_t)a.ns + ((a.zc_map && a.conf.boug != 1) ? 2 : 1) * (size_t)a.ns * FIT_NT) * sizeof(double);
    const size_t smem1 = smem0 + 4 * (size_t)a.ns * FIT_NT * sizeof(double);
    const char *ce = getenv("PLATEFLEX_B200_FIT_CACHE");
    const bool cache = (ce ? atoi(ce) != 0 : true) && smem1 <= 113 * 1024;  // two CTAs per SM must still fit
    const size_t smem = cache ? smem1 : smem0;
    long long blocks = (ncells + FIT_NT - 1) / FIT_NT;
    const long long cap = (long long)nsm * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    auto launch = [&](auto kernel) -> int {
        PFX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<(unsigned)blocks, FIT_NT, smem, stream>>>(a);
        return PFX_OK;
    };
    // default: the lane-persistent kernel (needs the shared-memory cache and a work counter)
    const char *me = getenv("PLATEFLEX_B200_FIT");
    const bool lanes = cache && a.counter && !(me && !strcmp(me, "static"));
    if (lanes) {
        // two CTAs per SM, each lane pulls cells until the counter runs out; the dry-land form (five CTAs of 64
        // lanes, no depth array) is launched next to it and the device flag decides which of the two works
        const size_t smem_dry = (2 * (size_t)a.ns + 4 * (size_t)a.ns * FIT_NT_DRY) * sizeof(double);
        const long long wet_blocks = std::min<long long>((ncells + FIT_NT - 1) / FIT_NT, 2LL * nsm);
        const long long dry_blocks = std::min<long long>((ncells + FIT_NT_DRY - 1) / FIT_NT_DRY, (long long)FIT_CTAS_DRY * nsm);
        auto launch2 = [&](auto wet, auto dry) -> int {
            PFX_CUDA(cudaFuncSetAttribute(wet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PFX_CUDA(cudaFuncSetAttribute(dry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dry));
            wet<<<(unsigned)wet_blocks, FIT_NT, smem, stream>>>(a);
            dry<<<(unsigned)dry_blocks, FIT_NT_DRY, smem_dry, stream>>>(a);
            return PFX_OK;
        };
        PFX_CHECK(alph ? launch2(fit_lane_kernel<3, false>, fit_lane_kernel<3, true>)
                       : launch2(fit_lane_kernel<2, false>, fit_lane_kernel<2, true>));
    } else if (alph) {
        PFX_CHECK(cache ? launch(fit_cell_kernel<3, true>) : launch(fit_cell_kernel<3, false>));
    } else {
        PFX_CHECK(cache ? launch(fit_cell_kernel<2, true>) : launch(fit_cell_kernel<2, false>));
    }
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace

}  // namespace pfx

using namespace pfx;

// Per-call scratch of the lane-persistent fit kernel, allocated and freed in stream order (cudaMallocAsync): the
// work counter, the count and the list of the cells handed to the finisher.  Nothing is shared between calls, so
// concurrent fits on different streams cannot touch each other's counters.
namespace {
struct FitScratch {
    unsigned long long *words = nullptr;  // [0] work counter, [1] todo count, [2] dry-land flag (as int), [3] unused
    int *todo = nullptr;
    int todo_cap = 0;
};

int fit_scratch_alloc(FitScratch &sc, long long ncells, cudaStream_t st)
{
    long long cap = ncells;  // every cell may exceed the trial cap: the hand-over must never depend on a full list
    if (cap < 1) cap = 1;
    void *p = nullptr;
    const size_t bytes = 32 + (size_t)cap * sizeof(int);
    {   // keep freed scratch in the device's pool across synchronisation points (the default threshold of 0 hands
        // it back to the driver at every sync, and the next call pays milliseconds to map it again)
        static std::mutex mu;
        static bool done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (dev >= 0 && dev < 64 && !done[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            done[dev] = true;
        }
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("fit scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;
    }
    sc.words = static_cast<unsigned long long *>(p);
    sc.todo = reinterpret_cast<int *>(sc.words + 4);
    sc.todo_cap = (int)cap;
    PFX_CUDA(cudaMemsetAsync(sc.words, 0, 32, st));
    return PFX_OK;
}
}  // namespace

extern "C" {

int pfx_l2_fit_rows_shape(int ny_total, int row0, int nrows, int nn, int *first_out_row, int *out_rows)
{
    PFX_ARG(ny_total >= 1 && row0 >= 0 && nrows >= 0 && row0 + nrows <= ny_total && nn >= 1, "bad row range");
    const int cj0 = (row0 + nn - 1) / nn;
    int cj1 = (row0 + nrows + nn - 1) / nn;  // first output row whose grid row lies beyond this shard
    if (cj1 > ny_total / nn) cj1 = ny_total / nn;
    if (first_out_row) *first_out_row = cj0;
    if (out_rows) *out_rows = cj1 > cj0 ? cj1 - cj0 : 0;
    return PFX_OK;
}

int pfx_l2_fit_dev(int device, void *stream, const pfx_flex_conf *conf, const double *k_dev, int ns, int nx, int ny,
                   const double *wl_admit, const double *ewl_admit, const double *wl_coh, const double *ewl_coh,
                   const double *wd, const double *rhoc_map, const double *zc_map, const uint8_t *mask, int nn, int alph,
                   int atype, double *out_te, double *out_te_std, double *out_f, double *out_f_std, double *out_alpha,
                   double *out_alpha_std, double *out_chi2, int32_t *status)
{
    return pfx_l2_fit_rows_dev(device, stream, conf, k_dev, ns, nx, ny, 0, ny, wl_admit, ewl_admit, wl_coh, ewl_coh, wd,
                               rhoc_map, zc_map, mask, nn, alph, atype, out_te, out_te_std, out_f, out_f_std, out_alpha,
                               out_alpha_std, out_chi2, status);
}

int pfx_l2_fit_rows_dev(int device, void *stream, const pfx_flex_conf *conf, const double *k_dev, int ns, int nx,
                        int ny_total, int row0, int nrows, const double *wl_admit, const double *ewl_admit,
                        const double *wl_coh, const double *ewl_coh, const double *wd, const double *rhoc_map,
                        const double *zc_map, const uint8_t *mask, int nn, int alph, int atype, double *out_te,
                        double *out_te_std, double *out_f, double *out_f_std, double *out_alpha, double *out_alpha_std,
                        double *out_chi2, int32_t *status)
{
    const int ny = nrows;
    NvtxRange nvtx_range("pfx:l2_fit");
    PFX_ARG(ny_total >= 1 && row0 >= 0 && nrows >= 1 && row0 + nrows <= ny_total, "bad row range");
    PFX_ARG(conf && k_dev, "conf or k is null");
    PFX_ARG(ns >= 1 && ns <= 64, "ns must be in [1, 64]");
    PFX_ARG(nx >= 1 && ny >= 1 && nn >= 1, "bad grid shape or decimator");
    PFX_ARG(atype == PFX_ATYPE_ADMIT || atype == PFX_ATYPE_COH || atype == PFX_ATYPE_JOINT, "bad atype");
    PFX_ARG(wl_admit && ewl_admit && wl_coh && ewl_coh && wd, "null input map");
    PFX_ARG(out_te && out_te_std && out_f && out_f_std && out_chi2, "null output map");
    PFX_ARG(!alph || (out_alpha && out_alpha_std), "alph set but alpha outputs are null");
    const int np = alph ? 3 : 2;
    const int m = ((atype == PFX_ATYPE_JOINT) ? 2 : 1) * ns;
    PFX_ARG(m > np, "need more data points than parameters");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    FitArgs a;
    a.conf = *conf;
    a.k = k_dev;
    a.ns = ns;
    a.nx = nx;
    a.ny = ny;
    a.adm = wl_admit;
    a.eadm = ewl_admit;
    a.coh = wl_coh;
    a.ecoh = ewl_coh;
    a.wd = wd;
    a.rhoc_map = rhoc_map;
    a.zc_map = zc_map;
    a.mask = mask;
    a.nn = nn;
    a.atype = atype;
    a.fx = (nx > nn) ? (nx - nn + nn - 1) / nn : 0;  // len(range(0, nx-nn, nn)), classes.py:1218
    a.mx = nx / nn;                                  // classes.py:1189
    a.row0 = row0;
    pfx_l2_fit_rows_shape(ny_total, row0, nrows, nn, &a.cj0, &a.my);
    {
        const int fy_total = (ny_total > nn) ? (ny_total - nn + nn - 1) / nn : 0;  // fitted output rows of the grid
        const int hi = (a.cj0 + a.my < fy_total) ? a.cj0 + a.my : fy_total;
        a.fy = hi > a.cj0 ? hi - a.cj0 : 0;
    }
    a.o_te = out_te;
    a.o_te_std = out_te_std;
    a.o_f = out_f;
    a.o_f_std = out_f_std;
    a.o_a = out_alpha;
    a.o_a_std = out_alpha_std;
    a.o_chi2 = out_chi2;
    a.status = status;
    a.counter = nullptr;
    {
        const char *fb = getenv("PLATEFLEX_B200_FIT_BATCH");
        a.fetch_min = (fb && atoi(fb) >= 1 && atoi(fb) <= 32) ? atoi(fb) : 24;
    }
    const size_t mo = (size_t)a.mx * a.my;
    int rc = PFX_OK;
    auto zero = [&](void *p, size_t bytes) {
        if (p && rc == PFX_OK && cudaMemsetAsync(p, 0, bytes, st) != cudaSuccess) {
            set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = PFX_ERR_CUDA;
        }
    };
    zero(out_te, mo * 8);
    zero(out_te_std, mo * 8);
    zero(out_f, mo * 8);
    zero(out_f_std, mo * 8);
    zero(out_chi2, mo * 8);
    if (alph) {
        zero(out_alpha, mo * 8);
        zero(out_alpha_std, mo * 8);
    }
    zero(status, mo * 4);
    FitScratch scratch;
    const char *mode = getenv("PLATEFLEX_B200_FIT");
    const bool lanes_wanted = !(mode && (!strcmp(mode, "group") || !strcmp(mode, "static")));
    if (rc == PFX_OK && a.fx > 0 && a.fy > 0 && lanes_wanted) {
        rc = fit_scratch_alloc(scratch, (long long)a.fx * a.fy, st);
        if (rc == PFX_OK) {
            a.counter = scratch.words;
            // dry-land form of the lane kernel: one crustal thickness and no fitted cell under water
            // (PLATEFLEX_B200_FIT_DRY=off: always the general form)
            int *flag = reinterpret_cast<int *>(scratch.words + 2);
            a.dry_flag = flag;
            const char *de = getenv("PLATEFLEX_B200_FIT_DRY");
            if (!zc_map && !(de && !strcmp(de, "off"))) {
                PFX_CUDA(cudaMemsetAsync(flag, 1, 1, st));  // little-endian int 1
                int nsm_ = 148;
                cudaDeviceGetAttribute(&nsm_, cudaDevAttrMultiProcessorCount, device);
                pfx::launch_wet_kernel(a, flag, nsm_ * 4, st);
                PFX_CUDA(cudaGetLastError());
            }
            // PLATEFLEX_B200_FIT_CAP: model evaluations after which a lane hands its cell to the finisher (0: never)
            const char *ce = getenv("PLATEFLEX_B200_FIT_CAP");
            const int capv = ce ? atoi(ce) : (alph ? 0 : 64);  // the 3-parameter fit needs > 64 trials too often to pay
            if (capv > 0) {
                a.todo = scratch.todo;
                a.todo_count = scratch.words + 1;
                a.todo_cap = scratch.todo_cap;
                a.trial_cap = capv;
            }
        }
    }
    if (rc == PFX_OK && a.fx > 0 && a.fy > 0) {
        // range(0, nx-nn, nn) never reaches index mx*nn or beyond (classes.py:1218 vs 1189)
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        auto group_fit = [&](const FitArgs &g) -> int {  // cooperative variant: 8 or 16 lanes per cell
            if (ns <= 8) return launch_fit_np<8, 1>(g, alph, st, nsm);
            if (ns <= 16) return launch_fit_np<8, 2>(g, alph, st, nsm);
            if (ns <= 24) return launch_fit_np<8, 3>(g, alph, st, nsm);
            if (ns <= 32) return launch_fit_np<8, 4>(g, alph, st, nsm);
            if (ns <= 48) return launch_fit_np<16, 3>(g, alph, st, nsm);
            return launch_fit_np<16, 4>(g, alph, st, nsm);
        };
        if (mode && !strcmp(mode, "group")) {
            rc = group_fit(a);
        } else {
            rc = launch_fit_cell(a, alph, st, nsm);
            if (rc == PFX_OK && a.counter && a.todo) {  // the cells the lanes gave up after trial_cap evaluations
                FitArgs fin = a;
                fin.finisher = 1;
                rc = group_fit(fin);
            }
        }
    }
    if (scratch.words) cudaFreeAsync(scratch.words, st);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

int pfx_l2_fit_host(int device, const pfx_flex_conf *conf, const double *k, int ns, int nx, int ny,
                    const double *wl_admit, const double *ewl_admit, const double *wl_coh, const double *ewl_coh,
                    const double *wd, const double *rhoc_map, const double *zc_map, const uint8_t *mask, int nn, int alph,
                    int atype, double *out_te, double *out_te_std, double *out_f, double *out_f_std, double *out_alpha,
                    double *out_alpha_std, double *out_chi2, int32_t *status)
{
    PFX_ARG(k && nx >= 1 && ny >= 1 && ns >= 1 && nn >= 1, "bad arguments");
    PFX_ARG(wl_admit && ewl_admit && wl_coh && ewl_coh && wd, "null input map");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    const size_t nxy = (size_t)nx * ny, cube = nxy * ns, mo = (size_t)(nx / nn) * (ny / nn);
    std::vector<void *> bufs;
    int rc = PFX_OK;
    auto up = [&](const void *h, size_t bytes) -> void * {
        if (!h || rc != PFX_OK) return nullptr;
        void *d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 8) != cudaSuccess) {
            cudaGetLastError();
            set_error("device allocation of %zu bytes failed", bytes);
            rc = PFX_ERR_OOM;
            return nullptr;
        }
        bufs.push_back(d);
        if (staged_h2d(d, h, bytes, 0) != PFX_OK) rc = PFX_ERR_CUDA;  // pageable numpy maps: pinned chunks + thread team
        return d;
    };
    auto dn = [&](void *h, size_t bytes) -> void * {
        if (!h || rc != PFX_OK) return nullptr;
        void *d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 8) != cudaSuccess) {
            cudaGetLastError();
            set_error("device allocation of %zu bytes failed", bytes);
            rc = PFX_ERR_OOM;
            return nullptr;
        }
        bufs.push_back(d);
        return d;
    };
    const double *d_k = (const double *)up(k, ns * 8);
    const double *d_a = (const double *)up(wl_admit, cube * 8), *d_ea = (const double *)up(ewl_admit, cube * 8);
    const double *d_c = (const double *)up(wl_coh, cube * 8), *d_ec = (const double *)up(ewl_coh, cube * 8);
    const double *d_wd = (const double *)up(wd, nxy * 8);
    const double *d_rc = (const double *)up(rhoc_map, nxy * 8), *d_zc = (const double *)up(zc_map, nxy * 8);
    const uint8_t *d_m = (const uint8_t *)up(mask, nxy);
    double *o[7];
    double *h[7] = {out_te, out_te_std, out_f, out_f_std, out_alpha, out_alpha_std, out_chi2};
    for (int i = 0; i < 7; i++) o[i] = (double *)dn(h[i], mo * 8);
    int32_t *d_st = (int32_t *)dn(status, mo * 4);
    if (rc == PFX_OK)
        rc = pfx_l2_fit_dev(device, nullptr, conf, d_k, ns, nx, ny, d_a, d_ea, d_c, d_ec, d_wd, d_rc, d_zc, d_m, nn, alph,
                            atype, o[0], o[1], o[2], o[3], o[4], o[5], o[6], d_st);
    if (rc == PFX_OK) {
        for (int i = 0; i < 7; i++)
            if (h[i]) cudaMemcpyAsync(h[i], o[i], mo * 8, cudaMemcpyDeviceToHost, 0);
        if (status) cudaMemcpyAsync(status, d_st, mo * 4, cudaMemcpyDeviceToHost, 0);
        cudaError_t e = cudaStreamSynchronize(0);
        if (e != cudaSuccess) {
            set_error("fit failed: %s", cudaGetErrorString(e));
            rc = PFX_ERR_CUDA;
        }
    }
    for (void *p : bufs) cudaFree(p);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

int pfx_flex_xspec_host(int device, const pfx_flex_conf *conf, const double *k, int ns, int ncurves, const double *te,
                        const double *f, const double *alpha, const double *wd, double *admit, double *coh,
                        double *jac_admit, double *jac_coh)
{
    PFX_ARG(conf && k && te && f && alpha && admit && coh, "null pointer");
    PFX_ARG(ns >= 1 && ncurves >= 1, "ns, ncurves must be >= 1");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    const size_t n = (size_t)ns * ncurves;
    double *d = nullptr;
    const size_t total = ns + 4 * (size_t)ncurves + 8 * n;
    PFX_CUDA(cudaMalloc(&d, total * 8));
    double *d_k = d, *d_te = d_k + ns, *d_f = d_te + ncurves, *d_al = d_f + ncurves, *d_wd = d_al + ncurves,
           *d_adm = d_wd + ncurves, *d_coh = d_adm + n, *d_ja = d_coh + n, *d_jc = d_ja + 3 * n;
    int rc = PFX_OK;
    cudaMemcpyAsync(d_k, k, ns * 8, cudaMemcpyHostToDevice, 0);
    cudaMemcpyAsync(d_te, te, ncurves * 8, cudaMemcpyHostToDevice, 0);
    cudaMemcpyAsync(d_f, f, ncurves * 8, cudaMemcpyHostToDevice, 0);
    cudaMemcpyAsync(d_al, alpha, ncurves * 8, cudaMemcpyHostToDevice, 0);
    if (wd) cudaMemcpyAsync(d_wd, wd, ncurves * 8, cudaMemcpyHostToDevice, 0);
    xspec_kernel<<<(unsigned)((n + 127) / 128), 128>>>(*conf, d_k, ns, ncurves, d_te, d_f, d_al, wd ? d_wd : nullptr,
                                                       d_adm, d_coh, jac_admit ? d_ja : nullptr,
                                                       jac_coh ? d_jc : nullptr);
    cudaMemcpyAsync(admit, d_adm, n * 8, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(coh, d_coh, n * 8, cudaMemcpyDeviceToHost, 0);
    if (jac_admit) cudaMemcpyAsync(jac_admit, d_ja, 3 * n * 8, cudaMemcpyDeviceToHost, 0);
    if (jac_coh) cudaMemcpyAsync(jac_coh, d_jc, 3 * n * 8, cudaMemcpyDeviceToHost, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("flex_xspec failed: %s", cudaGetErrorString(e));
        rc = PFX_ERR_CUDA;
    }
    cudaFree(d);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"

// ---- the helper routines numpy.f2py also exports from flex.f90 / cpwt_sub.f90 -------------------------------
// flexfilter_top (flex.f90:74-86), flexfilter_bot (:94-106), decon (:115-132), tr_func (:141-169), defk
// (cpwt_sub.f90:41-62) and wave_function (cpwt_sub.f90:73-99).  The package itself never calls them; they exist so
// that a script using plateflex.flex.flex.decon(...) etc. keeps working.  Written as the Fortran writes them
// (IEEE divisions, no fused reciprocals): these are convenience entry points, not the fit's inner loop.
namespace pfx {
namespace {

enum { FH_TOP = 0, FH_BOT = 1, FH_DECON = 2, FH_TRFUNC = 3 };

__global__ void flex_helper_kernel(int op, pfx_flex_conf c, double rhof, double wd, int ns, const double *in0,
                                   const double *in1, const double *in2, const double *in3, double p0, double p1,
                                   double *o0, double *o1, double *o2, double *o3)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const double rcf = c.rhoc - rhof, drho = c.rhom - c.rhoc;
    if (op == FH_TOP) {
        o0[s] = -(rcf / drho) / (1. + in0[s] / drho / FLEX_G);
    } else if (op == FH_BOT) {
        o0[s] = -(rcf / drho) * (1. + in0[s] / rcf / FLEX_G);
    } else if (op == FH_DECON) {  // in0 = theta, in1 = phi, in2 = k, p0 = A
        const double theta = in0[s], phi = in1[s], k = in2[s];
        o0[s] = 1. / (1. - theta);
        o1[s] = 1. / (phi - 1.);
        const double top = p0 * rcf * exp(-k * wd), deep = drho * exp(-k * (c.zc + wd));
        o2[s] = 2. * FLEX_PI * FLEX_GC * (top + deep * theta) / (1. - theta);
        o3[s] = 2. * FLEX_PI * FLEX_GC * (top + deep * phi) / (phi - 1.);
    } else {  // tr_func: in0..3 = mu_h, mu_w, nu_h, nu_w; p0 = F, p1 = alpha -> o0/o1 = admit (re, im), o2 = coh
        const double mu_h = in0[s], mu_w = in1[s], nu_h = in2[s], nu_w = in3[s];
        const double r = rcf / drho, ff = p0 / (1. - p0), t = ff * r;
        const double hgr = nu_h * mu_h + nu_w * mu_w * (ff * ff) * (r * r) + (nu_h * mu_w + nu_w * mu_h) * t * cos(p1);
        const double hgi = (nu_h * mu_w - nu_w * mu_h) * t * sin(p1);
        const double hh = mu_h * mu_h + (mu_w * t) * (mu_w * t) + 2. * mu_h * mu_w * t * cos(p1);
        const double gg = nu_h * nu_h + (nu_w * t) * (nu_w * t) + 2. * nu_h * nu_w * t * cos(p1);
        o0[s] = hgr / hh;
        o1[s] = hgi / hh;
        const double cr = hgr / sqrt(hh) / sqrt(gg);
        o2[s] = cr * cr;
    }
}

// cpwt_sub.f90:41-62 and :73-99 in float64; wave is (nx, ny) complex, x fastest (the Fortran layout)
__global__ void defk_kernel(int n, double d, double *k)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dk = 2. * FLEX_PI / ((double)n * d);
    k[i] = (i <= n / 2) ? (double)i * dk : -((double)(n - i) * dk);
}

__global__ void wave_function_kernel(int nx, int ny, const double *kx, const double *ky, double k0, double scale,
                                     double angle, double2 *wave)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const double kx0 = k0 * cos(angle), ky0 = k0 * sin(angle), norm = scale / sqrt(FLEX_PI);
    const double a = scale * kx[i] - kx0, b = scale * ky[j] - ky0;
    const double main = exp(-0.5 * (a * a)) * exp(-0.5 * (b * b));
    const double adm = exp(-0.5 * ((kx[i] * kx[i] + ky[j] * ky[j]) + (kx0 * kx0 + ky0 * ky0)));
    wave[(size_t)j * nx + i] = make_double2(norm * (main - adm), 0.);
}

}  // namespace
}  // namespace pfx

extern "C" {

int pfx_flex_helper_host(int device, const pfx_flex_conf *conf, double rhof, double wd, int op, int ns, const double *in0,
                         const double *in1, const double *in2, const double *in3, double p0, double p1, double *o0,
                         double *o1, double *o2, double *o3)
{
    PFX_ARG(conf && ns >= 1 && op >= 0 && op <= 3 && in0 && o0, "bad arguments");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    PFX_CUDA(cudaMalloc(&d, (size_t)8 * ns * sizeof(double)));
    const double *hin[4] = {in0, in1, in2, in3};
    double *hout[4] = {o0, o1, o2, o3};
    for (int q = 0; q < 4; q++)
        if (hin[q]) cudaMemcpyAsync(d + (size_t)q * ns, hin[q], ns * sizeof(double), cudaMemcpyHostToDevice, 0);
    flex_helper_kernel<<<(ns + 127) / 128, 128>>>(op, *conf, rhof, wd, ns, d, d + ns, d + 2 * (size_t)ns, d + 3 * (size_t)ns,
                                                  p0, p1, d + 4 * (size_t)ns, d + 5 * (size_t)ns, d + 6 * (size_t)ns,
                                                  d + 7 * (size_t)ns);
    for (int q = 0; q < 4; q++)
        if (hout[q]) cudaMemcpyAsync(hout[q], d + (size_t)(4 + q) * ns, ns * sizeof(double), cudaMemcpyDeviceToHost, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        set_error("flex helper failed: %s", cudaGetErrorString(e));
        return PFX_ERR_CUDA;
    }
    return PFX_OK;
}

int pfx_defk_host(int device, int nx, int ny, double dx, double dy, double *kx, double *ky)
{
    PFX_ARG(nx >= 1 && ny >= 1 && dx > 0 && dy > 0 && kx && ky, "bad arguments");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    PFX_CUDA(cudaMalloc(&d, (size_t)(nx + ny) * sizeof(double)));
    defk_kernel<<<(nx + 127) / 128, 128>>>(nx, dx, d);
    defk_kernel<<<(ny + 127) / 128, 128>>>(ny, dy, d + nx);
    cudaMemcpyAsync(kx, d, nx * sizeof(double), cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(ky, d + nx, ny * sizeof(double), cudaMemcpyDeviceToHost, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        set_error("defk failed: %s", cudaGetErrorString(e));
        return PFX_ERR_CUDA;
    }
    return PFX_OK;
}

int pfx_wave_function_host(int device, int nx, int ny, const double *kx, const double *ky, double k0, double scale,
                           double angle, double *wave)
{
    PFX_ARG(nx >= 1 && ny >= 1 && kx && ky && wave, "bad arguments");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) PFX_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    const size_t n = (size_t)nx * ny;
    PFX_CUDA(cudaMalloc(&d, (nx + ny + 2 + 2 * n) * sizeof(double)));
    cudaMemcpyAsync(d, kx, nx * sizeof(double), cudaMemcpyHostToDevice, 0);
    cudaMemcpyAsync(d + nx, ky, ny * sizeof(double), cudaMemcpyHostToDevice, 0);
    double2 *w = reinterpret_cast<double2 *>(d + ((nx + ny + 1) & ~1));
    wave_function_kernel<<<dim3((nx + 127) / 128, ny), 128>>>(nx, ny, d, d + nx, k0, scale, angle, w);
    cudaMemcpyAsync(wave, w, n * sizeof(double2), cudaMemcpyDeviceToHost, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        set_error("wave_function failed: %s", cudaGetErrorString(e));
        return PFX_ERR_CUDA;
    }
    return PFX_OK;
}

}  // extern "C"
