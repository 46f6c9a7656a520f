/*
 * oracle/pfx_oracle_impl.h -- precision-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement of the reference's
 * Fortran (src/cpwt/cpwt.f90, src/cpwt/cpwt_sub.f90) used as the parity checker
 * for the CUDA path.  Nothing in plateflex_b200/ may include, link or call it.
 *
 * It is included twice by pfx_oracle.c:
 *   REAL=float,  TWID=double       SUF=32   reference-faithful precision
 *                                           (default-kind REAL/COMPLEX, cpwt.f90:36,41,72-74)
 *   REAL=double, TWID=long double  SUF=64   the <=1e-9 parity target
 *
 * PARITY STATUS: "parity unpinned" for the numeric CWT/jackknife path -- the
 * reference ships no tests or golden vectors for it, and no Fortran compiler
 * exists in this image, so the restatement cannot be run against the Fortran.
 * It is pinned only (a) structurally, routine by routine, against the cited
 * lines, (b) by an independent numpy.fft restatement in tests/, and (c) by the
 * _lam2k known-answer vectors (classes.py:1787-1796) on the host side.
 *
 * All arrays use the reference's column-major (Fortran) layout:
 *   grid(ix,iy)            -> grid[ix + nx*iy]
 *   cube(ix,iy,ia,is)      -> cube[ix + nx*(iy + ny*(ia + na*is))]
 * complex values are interleaved (re, im) pairs of REAL.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(ora, SUF), CAT(_, name))

#ifndef PFX_ORACLE_NA
#define PFX_ORACLE_NA 11 /* cpwt.f90:37 */
#endif

/* cpwt_sub.f90:26-33 -- smallest power of two >= n (never below 2). */
int FN(npow2)(int n)
{
    int n2 = 1;
    for (;;) {
        n2 *= 2;
        if (n2 >= n) return n2;
    }
}

/* cpwt_sub.f90:41-62 -- wavenumber axes; the Nyquist bin keeps a POSITIVE
 * wavenumber, the upper half mirrors the lower half with a minus sign. */
void FN(defk)(int nx, int ny, REAL dx, REAL dy, REAL *kx, REAL *ky)
{
    const REAL pi = (REAL)3.141592653589793;
    REAL dkx = 2 * pi / ((REAL)nx * dx);
    REAL dky = 2 * pi / ((REAL)ny * dy);
    for (int ix = 1; ix <= nx / 2 + 1; ix++) kx[ix - 1] = (REAL)(ix - 1) * dkx;
    for (int iy = 1; iy <= ny / 2 + 1; iy++) ky[iy - 1] = (REAL)(iy - 1) * dky;
    for (int ix = nx / 2 + 2; ix <= nx; ix++) kx[ix - 1] = -kx[nx - ix + 2 - 1];
    for (int iy = ny / 2 + 2; iy <= ny; iy++) ky[iy - 1] = -ky[ny - iy + 2 - 1];
}

/* cpwt_sub.f90:73-99 -- fan-Morlet daughter in k-space.  Note the second
 * (admissibility) exponential uses UNSCALED wavenumbers (line 93). */
void FN(wave_function)(int nx, int ny, const REAL *kx, const REAL *ky, REAL k0,
                       REAL s, REAL a, REAL *daughter /* complex nx*ny */)
{
    const REAL pi = (REAL)3.141592653589793;
    REAL kx0 = k0 * XCOS(a);
    REAL ky0 = k0 * XSIN(a);
    REAL norm = s / XSQRT(pi);
    for (int iy = 0; iy < ny; iy++) {
        for (int ix = 0; ix < nx; ix++) {
            REAL tx = s * kx[ix] - kx0;
            REAL ty = s * ky[iy] - ky0;
            REAL expntx = (REAL)-0.5 * (tx * tx);
            REAL expnty = (REAL)-0.5 * (ty * ty);
            REAL d = norm * XEXP(expntx) * XEXP(expnty);
            REAL expnt = (REAL)-0.5 * ((kx[ix] * kx[ix] + ky[iy] * ky[iy]) + (kx0 * kx0 + ky0 * ky0));
            d = d - norm * XEXP(expnt);
            size_t o = 2 * ((size_t)ix + (size_t)nx * iy);
            daughter[o] = d;
            daughter[o + 1] = 0;
        }
    }
}

/* cpwt_sub.f90:435-449 -- even reflection of the (nx,ny) corner of a
 * (nnx,nny)=(2nx,2ny) array into the other three quadrants. */
void FN(mirror_data)(REAL *dat, int nx, int ny, int nnx, int nny)
{
    (void)nny;
    for (int iy = 1; iy <= ny; iy++) {
        for (int ix = 1; ix <= nx; ix++) {
#define D_(i, j) dat[((i)-1) + (size_t)nnx * ((j)-1)]
            D_(ix + nx, iy + ny) = D_(nx - ix + 1, ny - iy + 1);
            D_(ix + nx, iy) = D_(nx - ix + 1, iy);
            D_(ix, iy + ny) = D_(ix, ny - iy + 1);
#undef D_
        }
    }
}

/* cpwt_sub.f90:457-469 -- full index reversal of a complex (nx,ny) array. */
void FN(cfftswap)(REAL *spec, int nx, int ny)
{
    size_t n = (size_t)nx * ny;
    REAL *tmp = (REAL *)malloc(2 * n * sizeof(REAL));
    memcpy(tmp, spec, 2 * n * sizeof(REAL));
    for (int iy = 0; iy < ny; iy++) {
        for (int ix = 0; ix < nx; ix++) {
            size_t src = 2 * ((size_t)ix + (size_t)nx * iy);
            size_t dst = 2 * ((size_t)(nx - 1 - ix) + (size_t)nx * (ny - 1 - iy));
            spec[dst] = tmp[src];
            spec[dst + 1] = tmp[src + 1];
        }
    }
    free(tmp);
}

/* cpwt_sub.f90:522-594 -- in-place radix-2 N-dimensional FFT (the classic
 * bit-reversal + Danielson-Lanczos scheme).  data holds interleaved complex
 * values, first dimension fastest, 1-based indexing in the text below.  The
 * trigonometric recurrence runs in TWID precision and is truncated to REAL at
 * each butterfly (sngl() at cpwt_sub.f90:576-577).  isign=+1 gives
 * sum x*exp(+i...), isign=-1 gives sum x*exp(-i...); neither normalises. */
void FN(fourn)(REAL *data0, const int *nn, int ndim, int isign)
{
    REAL *data = data0 - 1; /* 1-based view */
    long ntot = 1;
    for (int idim = 0; idim < ndim; idim++) ntot *= nn[idim];
    long nprev = 1;
    for (int idim = 0; idim < ndim; idim++) {
        long n = nn[idim];
        long nrem = ntot / (n * nprev);
        long ip1 = 2 * nprev;
        long ip2 = ip1 * n;
        long ip3 = ip2 * nrem;
        long i2rev = 1;
        for (long i2 = 1; i2 <= ip2; i2 += ip1) { /* bit reversal */
            if (i2 < i2rev) {
                for (long i1 = i2; i1 <= i2 + ip1 - 2; i1 += 2) {
                    for (long i3 = i1; i3 <= ip3; i3 += ip2) {
                        long i3rev = i2rev + i3 - i2;
                        REAL tr = data[i3], ti = data[i3 + 1];
                        data[i3] = data[i3rev];
                        data[i3 + 1] = data[i3rev + 1];
                        data[i3rev] = tr;
                        data[i3rev + 1] = ti;
                    }
                }
            }
            long ibit = ip2 / 2;
            while (ibit >= ip1 && i2rev > ibit) {
                i2rev -= ibit;
                ibit /= 2;
            }
            i2rev += ibit;
        }
        long ifp1 = ip1;
        while (ifp1 < ip2) { /* butterflies */
            long ifp2 = 2 * ifp1;
            TWID theta = (TWID)isign * (TWID)6.28318530717959L / (TWID)(ifp2 / ip1);
            TWID sh = TWSIN((TWID)0.5 * theta);
            TWID wpr = (TWID)-2.0 * sh * sh;
            TWID wpi = TWSIN(theta);
            TWID wr = 1, wi = 0;
            for (long i3 = 1; i3 <= ifp1; i3 += ip1) {
                REAL wrs = (REAL)wr, wis = (REAL)wi;
                for (long i1 = i3; i1 <= i3 + ip1 - 2; i1 += 2) {
                    for (long i2 = i1; i2 <= ip3; i2 += ifp2) {
                        long k1 = i2, k2 = k1 + ifp1;
                        REAL tr = wrs * data[k2] - wis * data[k2 + 1];
                        REAL ti = wrs * data[k2 + 1] + wis * data[k2];
                        data[k2] = data[k1] - tr;
                        data[k2 + 1] = data[k1 + 1] - ti;
                        data[k1] = data[k1] + tr;
                        data[k1 + 1] = data[k1 + 1] + ti;
                    }
                }
                TWID wtemp = wr;
                wr = wr * wpr - wi * wpi + wr;
                wi = wi * wpr + wtemp * wpi + wi;
            }
            ifp1 = ifp2;
        }
        nprev = n * nprev;
    }
}

/* cpwt.f90:94-127 -- mirror, de-mean, zero-pad and forward-transform a grid.
 * Outputs the padded spectrum ft (complex nnx*nny) and the wavenumber axes. */
static void FN(prepare_spectrum)(const REAL *grid, int nx, int ny, REAL dx, REAL dy, int nnx, int nny,
                                 REAL *ft, REAL *kx, REAL *ky)
{
    int mx = 2 * nx, my = 2 * ny;
    REAL *mir = (REAL *)calloc((size_t)mx * my, sizeof(REAL));
    for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++) mir[ix + (size_t)mx * iy] = grid[ix + (size_t)nx * iy];
    FN(mirror_data)(mir, nx, ny, mx, my);
    /* sum(grid_mirror)/2./2./nx/ny (cpwt.f90:111): left-to-right accumulation, in REAL for
     * the float32 restatement (what the Fortran SUM does) and in long double for float64. */
    SUMT sum = 0;
    for (size_t i = 0; i < (size_t)mx * my; i++) sum += mir[i];
    REAL mean = (REAL)sum / (REAL)2 / (REAL)2 / (REAL)nx / (REAL)ny;
    memset(ft, 0, 2 * (size_t)nnx * nny * sizeof(REAL));
    for (int iy = 0; iy < my; iy++)
        for (int ix = 0; ix < mx; ix++) ft[2 * (ix + (size_t)nnx * iy)] = mir[ix + (size_t)mx * iy] - mean;
    free(mir);
    int nn[2] = {nnx, nny};
    FN(fourn)(ft, nn, 2, 1); /* cpwt.f90:122 */
    FN(defk)(nnx, nny, dx, dy, kx, ky);
}

/* cpwt.f90:157-185 -- one (scale, azimuth) plane of the transform written
 * into out(nx,ny) (complex, leading dimension nx). work = 2 complex planes. */
static void FN(one_plane)(const REAL *ft, const REAL *kx, const REAL *ky, int nx, int ny, int nnx, int nny,
                          REAL k0, REAL scales, REAL angle, REAL *daughter, REAL *out)
{
    size_t P = (size_t)nnx * nny;
    FN(wave_function)(nnx, nny, kx, ky, k0, scales, angle, daughter);
    for (size_t i = 0; i < P; i++) { /* daughter*CONJG(ft_grid), cpwt.f90:169 */
        REAL dr = daughter[2 * i], di = daughter[2 * i + 1];
        REAL fr = ft[2 * i], fi = -ft[2 * i + 1];
        daughter[2 * i] = dr * fr - di * fi;
        daughter[2 * i + 1] = dr * fi + di * fr;
    }
    int nn[2] = {nnx, nny};
    FN(fourn)(daughter, nn, 2, -1); /* cpwt.f90:173 */
    REAL fnorm = (REAL)((float)(nnx * nny)); /* FLOAT(nnx*nny), cpwt.f90:177 */
    for (size_t i = 0; i < 2 * P; i++) daughter[i] = daughter[i] / fnorm;
    FN(cfftswap)(daughter, nnx, nny); /* cpwt.f90:181 */
    for (int iy = 0; iy < ny; iy++) /* cpwt.f90:185 */
        memcpy(out + 2 * (size_t)nx * iy, daughter + 2 * (size_t)nnx * iy, 2 * (size_t)nx * sizeof(REAL));
}

/* cpwt.f90:65-191 -- wlet_transform.  grid(nx,ny) column-major; kf in rad/m;
 * dx,dy in km; wt(nx,ny,na,ns) complex.  nthreads>1 spreads the independent
 * (scale, azimuth) planes over OpenMP threads (the reference is serial; the
 * arithmetic per plane is unchanged). */
void FN(wlet_transform)(const REAL *grid, int nx, int ny, REAL dx, REAL dy, const REAL *kf, int ns, REAL k0,
                        REAL *wt, int nthreads)
{
    const int na = PFX_ORACLE_NA;
    const REAL pi = (REAL)3.141592653589793;
    int nnx = FN(npow2)(2 * nx), nny = FN(npow2)(2 * ny);
    size_t P = (size_t)nnx * nny;
    REAL *ft = (REAL *)malloc(2 * P * sizeof(REAL));
    REAL *kx = (REAL *)malloc(nnx * sizeof(REAL));
    REAL *ky = (REAL *)malloc(nny * sizeof(REAL));
    FN(prepare_spectrum)(grid, nx, ny, dx, dy, nnx, nny, ft, kx, ky);
    REAL da = (REAL)2. * XSQRT((REAL)-2. * XLOG((REAL)0.75)) / k0; /* cpwt.f90:131 */
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        REAL *daughter = (REAL *)malloc(2 * P * sizeof(REAL));
#pragma omp for collapse(2) schedule(dynamic)
        for (int is = 0; is < ns; is++) {
            for (int ia = 0; ia < na; ia++) {
                REAL kk = kf[is] * (REAL)1.e3; /* cpwt.f90:148 */
                REAL scales = k0 / kk;         /* cpwt.f90:150 */
                REAL angle = (REAL)ia * da - pi / (REAL)2.e0; /* cpwt.f90:161 */
                REAL *out = wt + 2 * (size_t)nx * ny * ((size_t)ia + (size_t)na * is);
                FN(one_plane)(ft, kx, ky, nx, ny, nnx, nny, k0, scales, angle, daughter, out);
            }
        }
        free(daughter);
    }
    free(ft);
    free(kx);
    free(ky);
}

/* A bounded SAMPLE of wlet_transform for timing the CPU path on grids where one full scale takes minutes
 * (bench.py --impl reference at 4096^2): the azimuth planes [ia0, ia1) of ONE scale, same arithmetic per plane
 * as above; wt is (nx, ny, ia1-ia0).  *t_forward receives the seconds spent in mirror + demean + forward fourn
 * (cpwt.f90:94-122), which the full job pays once per grid rather than once per sample. */
void FN(wlet_planes)(const REAL *grid, int nx, int ny, REAL dx, REAL dy, REAL kscale, REAL k0, int ia0, int ia1,
                     REAL *wt, int nthreads, double *t_forward)
{
    const REAL pi = (REAL)3.141592653589793;
    int nnx = FN(npow2)(2 * nx), nny = FN(npow2)(2 * ny);
    size_t P = (size_t)nnx * nny;
    REAL *ft = (REAL *)malloc(2 * P * sizeof(REAL));
    REAL *kx = (REAL *)malloc(nnx * sizeof(REAL));
    REAL *ky = (REAL *)malloc(nny * sizeof(REAL));
    double t0 = omp_get_wtime();
    FN(prepare_spectrum)(grid, nx, ny, dx, dy, nnx, nny, ft, kx, ky);
    if (t_forward) *t_forward = omp_get_wtime() - t0;
    REAL da = (REAL)2. * XSQRT((REAL)-2. * XLOG((REAL)0.75)) / k0; /* cpwt.f90:131 */
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        REAL *daughter = (REAL *)malloc(2 * P * sizeof(REAL));
#pragma omp for schedule(dynamic)
        for (int ia = ia0; ia < ia1; ia++) {
            REAL kk = kscale * (REAL)1.e3;
            REAL scales = k0 / kk;
            REAL angle = (REAL)ia * da - pi / (REAL)2.e0;
            FN(one_plane)(ft, kx, ky, nx, ny, nnx, nny, k0, scales, angle, daughter, wt + 2 * (size_t)nx * ny * (size_t)(ia - ia0));
        }
        free(daughter);
    }
    free(ft);
    free(kx);
    free(ky);
}

/* cpwt_sub.f90:107-160 -- running-prefix jackknife of a real (nx,ny,na,ns)
 * array over the azimuth axis. */
void FN(jackknife_1d_sgram)(const REAL *ss, int nx, int ny, int na, int ns, REAL *ess)
{
    size_t nxy = (size_t)nx * ny;
    REAL *ssj = (REAL *)malloc((na - 1) * sizeof(REAL));
    REAL *ssj2 = (REAL *)malloc(na * sizeof(REAL));
    for (int is = 0; is < ns; is++) {
        for (size_t c = 0; c < nxy; c++) {
            for (int ia = 0; ia < na; ia++) {
                int ik = 0;
                REAL p1 = 0;
                for (int ia2 = 0; ia2 < na; ia2++) {
                    if (ia2 == ia) continue;
                    p1 = p1 + ss[c + nxy * ((size_t)ia2 + (size_t)na * is)];
                    ssj[ik++] = (p1 != 0) ? p1 : (REAL)0;
                }
                REAL s = 0;
                for (int i = 0; i < na - 1; i++) s += ssj[i];
                ssj2[ia] = s / (REAL)(na - 1);
            }
            REAL mss = 0;
            for (int ia = 0; ia < na; ia++) mss += ssj2[ia];
            mss = mss / (REAL)na;
            REAL dss = 0;
            for (int ia = 0; ia < na; ia++) dss = dss + (ssj2[ia] - mss) * (ssj2[ia] - mss);
            dss = XSQRT(dss * (REAL)(na - 1) / (REAL)na);
            if (dss != dss) dss = 0;
            ess[c + nxy * is] = dss;
        }
    }
    free(ssj);
    free(ssj2);
}

/* cpwt_sub.f90:298-366 -- running-prefix jackknife of admittance Re(xp)/p1 and
 * coherence Re(xp)^2/(p1*p2). sxp complex, ss1/ss2 real, all (nx,ny,na,ns). */
void FN(jackknife_admit_coh)(const REAL *sxp, const REAL *ss1, const REAL *ss2, int nx, int ny, int na, int ns,
                             REAL *sad, REAL *sco)
{
    size_t nxy = (size_t)nx * ny;
    REAL *co = (REAL *)malloc((na - 1) * sizeof(REAL)), *ad = (REAL *)malloc((na - 1) * sizeof(REAL));
    REAL *co2 = (REAL *)malloc(na * sizeof(REAL)), *ad2 = (REAL *)malloc(na * sizeof(REAL));
    for (int is = 0; is < ns; is++) {
        for (size_t c = 0; c < nxy; c++) {
            for (int ia = 0; ia < na; ia++) {
                int ik = 0;
                REAL p1 = 0, p2 = 0, xr = 0, xi = 0;
                for (int ia2 = 0; ia2 < na; ia2++) {
                    if (ia2 == ia) continue;
                    size_t o = c + nxy * ((size_t)ia2 + (size_t)na * is);
                    p1 = p1 + ss1[o];
                    p2 = p2 + ss2[o];
                    xr = xr + sxp[2 * o];
                    xi = xi + sxp[2 * o + 1];
                    if (p1 != 0 && p2 != 0) {
                        co[ik] = xr * xr / (p1 * p2);
                        ad[ik] = xr / p1;
                    } else {
                        co[ik] = 0;
                        ad[ik] = 0;
                    }
                    ik++;
                }
                (void)xi;
                REAL sc = 0, sa = 0;
                for (int i = 0; i < na - 1; i++) sc += co[i];
                for (int i = 0; i < na - 1; i++) sa += ad[i];
                co2[ia] = sc / (REAL)(na - 1);
                ad2[ia] = sa / (REAL)(na - 1);
            }
            REAL mco = 0, mad = 0;
            for (int ia = 0; ia < na; ia++) mco += co2[ia];
            for (int ia = 0; ia < na; ia++) mad += ad2[ia];
            mco = mco / (REAL)na;
            mad = mad / (REAL)na;
            REAL dco = 0, dad = 0;
            for (int ia = 0; ia < na; ia++) {
                dco = dco + (co2[ia] - mco) * (co2[ia] - mco);
                dad = dad + (ad2[ia] - mad) * (ad2[ia] - mad);
            }
            dco = XSQRT(dco * (REAL)(na - 1) / (REAL)na);
            dad = XSQRT(dad * (REAL)(na - 1) / (REAL)na);
            if (dco != dco) dco = 0;
            if (dad != dad) dad = 0;
            sco[c + nxy * is] = dco;
            sad[c + nxy * is] = dad;
        }
    }
    free(co);
    free(ad);
    free(co2);
    free(ad2);
}

/* cpwt.f90:200-241 -- wlet_scalogram from a materialised cube. */
void FN(wlet_scalogram)(const REAL *wl, int nx, int ny, int ns, REAL *wl_sg, REAL *ewl_sg)
{
    const int na = PFX_ORACLE_NA;
    size_t nxy = (size_t)nx * ny;
    REAL *sg = (REAL *)malloc(nxy * na * ns * sizeof(REAL));
    for (size_t i = 0; i < nxy * ns; i++) wl_sg[i] = 0;
    for (int is = 0; is < ns; is++) {
        for (int ia = 0; ia < na; ia++) {
            size_t o = nxy * ((size_t)ia + (size_t)na * is);
            for (size_t c = 0; c < nxy; c++) {
                REAL re = wl[2 * (o + c)], im = wl[2 * (o + c) + 1];
                /* CABS(w*CONJG(w)): the product is (re*re+im*im, 0) */
                REAL pr = re * re + im * im;
                sg[o + c] = XFABS(pr);
                wl_sg[c + nxy * is] = wl_sg[c + nxy * is] + sg[o + c];
            }
        }
    }
    for (size_t i = 0; i < nxy * ns; i++) wl_sg[i] = wl_sg[i] / (REAL)na;
    FN(jackknife_1d_sgram)(sg, nx, ny, na, ns, ewl_sg);
    free(sg);
}

/* cpwt.f90:250-293 -- wlet_xscalogram from two cubes; wl_xsg complex. */
void FN(wlet_xscalogram)(const REAL *w1, const REAL *w2, int nx, int ny, int ns, REAL *wl_xsg, REAL *ewl_xsg)
{
    const int na = PFX_ORACLE_NA;
    size_t nxy = (size_t)nx * ny, ncube = nxy * na * ns;
    REAL *xre = (REAL *)malloc(ncube * sizeof(REAL)), *xim = (REAL *)malloc(ncube * sizeof(REAL));
    REAL *ere = (REAL *)malloc(nxy * ns * sizeof(REAL)), *eim = (REAL *)malloc(nxy * ns * sizeof(REAL));
    for (size_t i = 0; i < 2 * nxy * ns; i++) wl_xsg[i] = 0;
    for (int is = 0; is < ns; is++) {
        for (int ia = 0; ia < na; ia++) {
            size_t o = nxy * ((size_t)ia + (size_t)na * is);
            for (size_t c = 0; c < nxy; c++) {
                REAL a1 = w1[2 * (o + c)], b1 = w1[2 * (o + c) + 1];
                REAL a2 = w2[2 * (o + c)], b2 = -w2[2 * (o + c) + 1]; /* CONJG */
                xre[o + c] = a1 * a2 - b1 * b2;
                xim[o + c] = a1 * b2 + b1 * a2;
                wl_xsg[2 * (c + nxy * is)] += xre[o + c];
                wl_xsg[2 * (c + nxy * is) + 1] += xim[o + c];
            }
        }
    }
    for (size_t i = 0; i < 2 * nxy * ns; i++) wl_xsg[i] = wl_xsg[i] / (REAL)na;
    FN(jackknife_1d_sgram)(xre, nx, ny, na, ns, ere);
    FN(jackknife_1d_sgram)(xim, nx, ny, na, ns, eim);
    for (size_t i = 0; i < nxy * ns; i++) ewl_xsg[i] = XSQRT(ere[i] * ere[i] + eim[i] * eim[i]);
    free(xre);
    free(xim);
    free(ere);
    free(eim);
}

/* cpwt.f90:303-364 -- wlet_admit_coh from two cubes.  The reference never
 * zero-initialises its accumulators (cpwt.f90:318-319); zero is what fresh
 * pages give it in practice and what we define here. */
void FN(wlet_admit_coh)(const REAL *w1, const REAL *w2, int nx, int ny, int ns, REAL *wl_admit, REAL *ewl_admit,
                        REAL *wl_coh, REAL *ewl_coh)
{
    const int na = PFX_ORACLE_NA;
    size_t nxy = (size_t)nx * ny, ncube = nxy * na * ns;
    REAL *sg1 = (REAL *)malloc(ncube * sizeof(REAL)), *sg2 = (REAL *)malloc(ncube * sizeof(REAL));
    REAL *xsg = (REAL *)malloc(2 * ncube * sizeof(REAL));
    REAL *m1 = (REAL *)calloc(nxy * ns, sizeof(REAL)), *m2 = (REAL *)calloc(nxy * ns, sizeof(REAL));
    REAL *mx = (REAL *)calloc(2 * nxy * ns, sizeof(REAL));
    for (int is = 0; is < ns; is++) {
        for (int ia = 0; ia < na; ia++) {
            size_t o = nxy * ((size_t)ia + (size_t)na * is);
            for (size_t c = 0; c < nxy; c++) {
                REAL a1 = w1[2 * (o + c)], b1 = w1[2 * (o + c) + 1];
                REAL a2 = w2[2 * (o + c)], b2 = w2[2 * (o + c) + 1];
                sg1[o + c] = XFABS(a1 * a1 + b1 * b1);
                sg2[o + c] = XFABS(a2 * a2 + b2 * b2);
                REAL cb2 = -b2;
                xsg[2 * (o + c)] = a1 * a2 - b1 * cb2;
                xsg[2 * (o + c) + 1] = a1 * cb2 + b1 * a2;
                size_t q = c + nxy * is;
                m1[q] = m1[q] + sg1[o + c];
                m2[q] = m2[q] + sg2[o + c];
                mx[2 * q] = mx[2 * q] + xsg[2 * (o + c)];
                mx[2 * q + 1] = mx[2 * q + 1] + xsg[2 * (o + c) + 1];
            }
        }
    }
    for (size_t q = 0; q < nxy * ns; q++) {
        REAL s1 = m1[q] / (REAL)na, s2 = m2[q] / (REAL)na, xr = mx[2 * q] / (REAL)na;
        wl_admit[q] = xr / s1;              /* cpwt.f90:354 */
        wl_coh[q] = (xr * xr) / (s1 * s2);  /* cpwt.f90:355 */
    }
    FN(jackknife_admit_coh)(xsg, sg1, sg2, nx, ny, na, ns, ewl_admit, ewl_coh);
    free(sg1);
    free(sg2);
    free(xsg);
    free(m1);
    free(m2);
    free(mx);
}

/* Convenience for sizes where the cubes do not fit: the chain
 * wlet_transform(topo), wlet_transform(grav), wlet_admit_coh evaluated one
 * scale at a time (each scale is independent in every routine above, so the
 * arithmetic per output element is identical).  OpenMP over azimuth planes. */
void FN(admit_coh_from_grids)(const REAL *topo, const REAL *grav, int nx, int ny, REAL dx, REAL dy, const REAL *kf,
                              int ns, REAL k0, REAL *wl_admit, REAL *ewl_admit, REAL *wl_coh, REAL *ewl_coh,
                              int nthreads)
{
    const int na = PFX_ORACLE_NA;
    size_t nxy = (size_t)nx * ny;
    REAL *c1 = (REAL *)malloc(2 * nxy * na * sizeof(REAL)), *c2 = (REAL *)malloc(2 * nxy * na * sizeof(REAL));
    for (int is = 0; is < ns; is++) {
        FN(wlet_transform)(topo, nx, ny, dx, dy, kf + is, 1, k0, c1, nthreads);
        FN(wlet_transform)(grav, nx, ny, dx, dy, kf + is, 1, k0, c2, nthreads);
        FN(wlet_admit_coh)(c1, c2, nx, ny, 1, wl_admit + nxy * is, ewl_admit + nxy * is, wl_coh + nxy * is,
                           ewl_coh + nxy * is);
    }
    free(c1);
    free(c2);
}

#undef FN
#undef CAT
#undef CAT_
