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
#include <cstring>

#include "pfx_dct.hpp"
#include "pfx_plan.hpp"
#include "pfx_pruned.hpp"
#include "pfx_pruned_v2.cuh"
#include "pfx_pruned_v3.cuh"
#include "pfx_pruned_v4.cuh"

namespace pfx {

namespace {

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

// Admissibility term on the y-fastest layout (cpwt_sub.f90:93).  exp(-kx^2/2) exp(-ky^2/2) is real and even and
// the spectrum of a (real) grid is Hermitian, so the transform of the product is real.  One plane per grid: packing
// both grids into one complex transform (X_0 + i X_1) would halve the work but ties the rounding error of each
// plane to the amplitude of the larger one -- with grids in very different units (topography in metres, gravity
// in m/s^2: seven orders of magnitude) that eats the 1e-9 budget of the smaller one.
__global__ void zmul_kernel(int NX, int NY, const double2 *__restrict__ s0, const double *__restrict__ u2,
                            const double *__restrict__ v2, double2 *plane)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= NY) return;
    const size_t o = (size_t)a * NY + b;
    const double w = u2[a] * v2[b];
    const double2 x = s0[o];
    plane[o] = make_double2(w * x.x, w * x.y);
}

// zc[j][i].{x|y} = scale * Re plane[i+1][j+1] = Z_g(i, j) (comp = g): crop with the one-sample shift, transposed to
// the x-fastest layout of the coefficient planes through a 32 x 32 shared-memory tile; `clear` zeroes the other
// component (single-grid calls)
__global__ void __launch_bounds__(256)
    zcrop_kernel(int nx, int ny, int NY, double scale, const double2 *__restrict__ plane, double2 *zc, int comp, int clear)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        if (i < nx && j < ny) tile[r][threadIdx.x] = plane[(size_t)(i + 1) * NY + (j + 1)].x;
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int i = i0 + threadIdx.x, j = j0 + c;
        if (i < nx && j < ny) {
            double *z = reinterpret_cast<double *>(zc + (size_t)j * nx + i);
            z[comp] = tile[threadIdx.x][c] * scale;
            if (clear) z[comp ^ 1] = 0.0;
        }
    }
}

int ilog2(int v)
{
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

// Choose the FFT length of an axis for a support of K samples and fit the line in shared memory.
void plan_axis(AxisDesc &A, int N, int K, int nout, size_t smem_budget_bytes, int cpc)
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
    A.rc = A.r;
    A.logrc = A.logr;
    A.pad = (A.rc < 8 && A.logL >= 8) ? 1 : 0;
    // residues per pass over shared memory: all of them unless the line does not fit
    while (A.rc > 1 && prn_smem_bytes(A, cpc) > smem_budget_bytes) {
        A.rc /= 2;
        A.logrc--;
        A.pad = (A.rc < 8 && A.logL >= 8) ? 1 : 0;
    }
}

}  // namespace

struct PrunedState {
    std::vector<ScaleDesc> sd;
    double2 *rootx = nullptr, *rooty = nullptr;
    double *tabu = nullptr, *tabv = nullptr;
    double2 *twa = nullptr, *twb = nullptr;
    double2 *T = nullptr, *Tbuf[2] = {nullptr, nullptr}, *W = nullptr, *Wbuf[2] = {nullptr, nullptr}, *zc = nullptr, *zplanes = nullptr;
    cufftHandle fft_z = 0;
    bool fft_z_ok = false;
    size_t smem1 = 0, smem2 = 0;
    int cpc = 2;
    int ctas_per_transform = 128;
    // second-generation kernels (pfx_pruned_v2.cuh), chosen per scale and axis: variant < 0 = first generation
    std::vector<int> v2x, v2y, v2x_ctas, v2y_ctas;
    // third-generation pass 2 (pfx_pruned_v3.cuh): line buffers in the ring (0 = not used for this scale)
    std::vector<int> v3x_nbuf, v3x_ctas, v3y_nbuf, v3y_ctas;
    // 128-thread form of the second-generation pass 2 (namespace v2h): CTAs resident per SM, 0 = not used
    std::vector<int> v2hx_ctas;
    // fourth-generation (radix-16, 128-thread) kernels of pfx_pruned_v4.cuh: CTAs resident on the device, 0 = not used
    std::vector<int> v4x_ctas, v4y_ctas;
    DctState *dct = nullptr;  // cosine-transform form of the forward spectra and of Z (pfx_dct.cu), or null
    int v4_waves_p2 = 2;  // pass-2 grid as a multiple of the resident CTAs (PLATEFLEX_B200_V4_WAVES; 2 measured best)
    int v3_occ = 3;  // CTAs per SM the v3 kernels are compiled for (PLATEFLEX_B200_V3_OCC = 2 | 3)
    int nsm = 148;
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
    size_t budget = (size_t)dev_smem;

    st->cpc = (NX <= 4096) ? 2 : 1;
    {
        const char *b = getenv("PLATEFLEX_B200_SMEM_BUDGET");  // experiment knob: shared memory per CTA, bytes
        if (b && atol(b) > 0) budget = std::min(budget, (size_t)atol(b));
        const char *c2 = getenv("PLATEFLEX_B200_CPC");
        if (c2 && (atoi(c2) == 1 || atoi(c2) == 2)) st->cpc = atoi(c2);
    }
    {
        const char *c = getenv("PLATEFLEX_B200_CTAS");
        if (c && atoi(c) > 0) st->ctas_per_transform = atoi(c);
    }
    st->sd.resize(ns);
    st->v2x.assign(ns, -1);
    st->v2y.assign(ns, -1);
    st->v2x_ctas.assign(ns, 0);
    st->v2y_ctas.assign(ns, 0);
    st->v3x_nbuf.assign(ns, 0);
    st->v3x_ctas.assign(ns, 0);
    st->v3y_nbuf.assign(ns, 0);
    st->v3y_ctas.assign(ns, 0);
    st->v2hx_ctas.assign(ns, 0);
    st->v4x_ctas.assign(ns, 0);
    st->v4y_ctas.assign(ns, 0);
    // PLATEFLEX_B200_V4 = "on" | "p1" | "p2" | "off": radix-16 pass kernels for both passes, one of them, or none.
    // Default "p2": measured on B200 (profiles/README.md) pass 2 gains 7-8 % (L = 128 and 1024: 16-18 %), pass 1
    // does not (its lines are short and few; the second generation keeps three CTAs of 256 threads per SM there).
    const char *v4_env = getenv("PLATEFLEX_B200_V4");
    const bool v4_p1 = v4_env && (!strcmp(v4_env, "on") || !strcmp(v4_env, "p1"));
    const bool v4_p2 = !v4_env || !strcmp(v4_env, "on") || !strcmp(v4_env, "p2");
    {
        const char *w = getenv("PLATEFLEX_B200_V4_WAVES");
        if (w && atoi(w) > 0) st->v4_waves_p2 = atoi(w);
    }
    // PLATEFLEX_B200_V2H = "on": pass 2 with 128-thread CTAs (parts of 1024 elements) where the FFT length allows
    const char *v2h_env = getenv("PLATEFLEX_B200_V2H");
    const bool v2h_on = v2h_env && !strcmp(v2h_env, "on");
    // PLATEFLEX_B200_V3 = "on" | "1" | "2" (ring depth): third-generation pass kernels with bulk-copy staged support
    // samples.  Opt-in: they need the intermediate transposed, and the scattered stores that costs pass 1 outweigh
    // what pass 2 gains on B200 (measured, profiles/README.md), so the second generation stays the default.
    const char *v3_env = getenv("PLATEFLEX_B200_V3");
    const bool v3_on = v3_env && (!strcmp(v3_env, "on") || v3_env[0] == '1' || v3_env[0] == '2');
    const int v3_depth = (v3_env && (v3_env[0] == '1' || v3_env[0] == '2')) ? v3_env[0] - '0' : 0;
    const char *v3p1_env = getenv("PLATEFLEX_B200_V3_P1");  // "off": second-generation pass 1
    const bool v3p1_on = v3_on && !(v3p1_env && !strcmp(v3p1_env, "off"));
    {
        const char *o = getenv("PLATEFLEX_B200_V3_OCC");
        if (o && (o[0] == '2' || o[0] == '3')) st->v3_occ = o[0] - '0';
    }
    PFX_CUDA(cudaDeviceGetAttribute(&st->nsm, cudaDevAttrMultiProcessorCount, pl->device));
    // PLATEFLEX_B200_V2 = "off" | "<lines per CTA 1|2>,<prefetch 0|1>"  (default "1,0")
    bool v2_on = true;
    int v2_cpc = 1, v2_pre = 0;
    {
        const char *e = getenv("PLATEFLEX_B200_V2");
        if (e && !strcmp(e, "off")) v2_on = false;
        else if (e && strlen(e) >= 3) {
            v2_cpc = (e[0] == '2') ? 2 : 1;
            v2_pre = (e[2] == '1') ? 1 : 0;
        }
    }
    size_t tabu_n = 0, tabv_n = 0, twa_n = 0, twb_n = 0, t_elems = 0;
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
        plan_axis(sd.ax, NX, std::max(kamax, 1), nx, budget, st->cpc);
        plan_axis(sd.ay, NY, std::max(kbmax, 1), ny, budget, 1);
        // second-generation kernels: padded axis of at least 2048 samples and 32 <= L <= 2048
        auto pick_v2 = [&](AxisDesc &A, int cpc, int pre) -> int {
            if (!v2_on || A.N < 2048 || A.logL > 11) return -1;
            if (A.logL < 5) {
                A.L = 32;
                A.logL = 5;
                A.r = A.N / A.L;
                A.logr = A.logN - A.logL;
                A.rc = A.r;
                A.logrc = A.logr;
                A.pad = 0;
            }
            const int multi = A.N > 2048;
            if (multi || pre) cpc = 1;
            return cpc | (pre << 4) | (multi << 5);
        };
        st->v2x[is] = pick_v2(sd.ax, v2_cpc, v2_pre);
        st->v2y[is] = pick_v2(sd.ay, v2_cpc, v2_pre);
        sd.kamax = std::max(kamax, 1);
        sd.tabu = tabu_n;
        sd.tabv = tabv_n;
        sd.twa = twa_n;
        sd.twb = twb_n;
        tabu_n += (size_t)PFX_NA * sd.ax.L;
        tabv_n += (size_t)PFX_NA * sd.ay.L;
        twa_n += (size_t)PFX_NA * NX;
        twb_n += (size_t)PFX_NA * NY;
        // v3 pass 2 (same conditions as v2, single line per CTA): T is stored transposed, one line of LP slots per row
        if (v3_on && st->v2x[is] >= 0) st->v3x_nbuf[is] = 1;  // ring depth settled below (occupancy)
        if (v3p1_on && st->v2y[is] >= 0) st->v3y_nbuf[is] = 1;
        t_elems = std::max(t_elems, st->v3x_nbuf[is] ? (size_t)ny * v3_line_slots(sd.ax.logL) : (size_t)sd.kamax * ny);
        if (st->v2y[is] < 0) st->smem1 = std::max(st->smem1, prn_smem_bytes(sd.ay, 1));
        if (st->v2x[is] < 0) st->smem2 = std::max(st->smem2, prn_smem_bytes(sd.ax, st->cpc));
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
    for (int b = 0; b < 2; b++) PFX_CHECK(dalloc((void **)&st->Tbuf[b], 2 * PFX_NA * t_elems * sizeof(double2)));
    st->T = st->Tbuf[0];
    for (int b = 0; b < 2; b++) PFX_CHECK(dalloc((void **)&st->Wbuf[b], 2 * PFX_NA * (size_t)nx * ny * sizeof(double2)));
    st->W = st->Wbuf[0];
    PFX_CHECK(dalloc((void **)&st->zc, (size_t)nx * ny * sizeof(double2)));
    // Power-of-two grids: spectra and Z through cosine transforms of the unmirrored grid, provided every scale's
    // pass 1 is a kernel that can read that form (second / fourth generation)
    bool use_dct = dct_applicable(pl);
    for (int is = 0; is < ns; is++) use_dct = use_dct && st->v2y[is] >= 0 && !st->v3y_nbuf[is];
    if (use_dct) {
        PFX_CHECK(dct_init(pl, &st->dct, st->owned));
    } else {
        PFX_CHECK(dalloc((void **)&st->zplanes, P * sizeof(double2)));
        for (int g = 0; g < 2; g++)
            if (!pl->d_spec[g]) PFX_CHECK(pl->alloc((void **)&pl->d_spec[g], P * sizeof(double2)));
    }

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

    // dense inverse plan for the admissibility term (one plane serves both grids), y-fastest layout
    int n[2] = {NX, NY};
    size_t need = pl->fft_work_bytes;
    if (!st->dct) {
        size_t ws = 0;
        PFX_CUFFT(cufftCreate(&st->fft_z));
        st->fft_z_ok = true;
        PFX_CUFFT(cufftSetAutoAllocation(st->fft_z, 0));
        PFX_CUFFT(cufftMakePlanMany(st->fft_z, 2, n, nullptr, 1, (int)P, nullptr, 1, (int)P, CUFFT_Z2Z, 1, &ws));
        need = std::max(need, ws);
    }
    if (need > pl->fft_work_bytes) {
        void *w = nullptr;
        PFX_CHECK(pl->alloc(&w, need));
        pl->d_fft_work = w;
        pl->fft_work_bytes = need;
        PFX_CUFFT(cufftSetWorkArea(pl->fft_fwd, pl->d_fft_work));
    }
    if (st->fft_z_ok) PFX_CUFFT(cufftSetWorkArea(st->fft_z, pl->d_fft_work));

    for (int is = 0; is < ns; is++) {
        const ScaleDesc &sd = st->sd[is];
        int res = 0;
        if (st->v3y_nbuf[is]) {
            const int multi = (st->v2y[is] >> 5) & 1, half = 2 * ny == NY;
            int nbuf = v3_depth ? v3_depth : 2;
            PFX_CHECK(v3_configure_p1(sd.ay.logL, multi, half, st->v3_occ, nbuf, &res));
            if (!v3_depth && res < st->v3_occ) {
                int res1 = 0;
                PFX_CHECK(v3_configure_p1(sd.ay.logL, multi, half, st->v3_occ, 1, &res1));
                if (res1 > res) {
                    nbuf = 1;
                    res = res1;
                } else {
                    PFX_CHECK(v3_configure_p1(sd.ay.logL, multi, half, st->v3_occ, 2, &res));
                }
            }
            st->v3y_nbuf[is] = nbuf;
            st->v3y_ctas[is] = std::max(res, 1) * st->nsm;
        } else if (st->v2y[is] >= 0) {
            PFX_CHECK(v2_configure_p1(sd.ay.logL, st->v2y[is], &res));
            st->v2y_ctas[is] = std::max(res, 1) * st->nsm;
        } else {
            PFX_CHECK(prn_configure_pass1(sd.ay.logL, sd.ay.pad, st->smem1));
        }
        if (st->dct && st->v2y[is] >= 0) st->v2y[is] = (st->v2y[is] & ~0x1f) | 1 | (1 << 6);  // single line, no prefetch
        if (v4_p1 && !st->v3y_nbuf[is] && st->v2y[is] >= 0) {
            PFX_CHECK(v4_configure_p1(sd.ay.logL, sd.ay.N > 2048, st->dct != nullptr, &res));
            st->v4y_ctas[is] = std::max(res, 1) * st->nsm;
        }
        if (v4_p2 && !st->v3x_nbuf[is] && st->v2x[is] >= 0) {
            PFX_CHECK(v4_configure_p2(sd.ax.logL, sd.ax.N > 2048, &res));
            st->v4x_ctas[is] = std::max(res, 1) * st->nsm;
        } else if (v2h_on && !st->v3x_nbuf[is] && st->v2x[is] >= 0 && sd.ax.logL <= 10) {
            PFX_CHECK(v2h_configure_p2(sd.ax.logL, sd.ax.N > 1024, &res));
            st->v2hx_ctas[is] = std::max(res, 1) * st->nsm;
        } else if (st->v3x_nbuf[is]) {
            // two line buffers when three CTAs per SM still fit, else one (refilled behind the first stage)
            const int multi = (st->v2x[is] >> 5) & 1, half = 2 * nx == NX;
            int nbuf = v3_depth ? v3_depth : 2;
            PFX_CHECK(v3_configure_p2(sd.ax.logL, multi, half, st->v3_occ, nbuf, &res));
            if (!v3_depth && res < st->v3_occ) {
                int res1 = 0;
                PFX_CHECK(v3_configure_p2(sd.ax.logL, multi, half, st->v3_occ, 1, &res1));
                if (res1 > res) {
                    nbuf = 1;
                    res = res1;
                } else {
                    PFX_CHECK(v3_configure_p2(sd.ax.logL, multi, half, st->v3_occ, 2, &res));
                }
            }
            st->v3x_nbuf[is] = nbuf;
            st->v3x_ctas[is] = std::max(res, 1) * st->nsm;
        } else if (st->v2x[is] >= 0) {
            PFX_CHECK(v2_configure_p2(sd.ax.logL, st->v2x[is], &res));
            st->v2x_ctas[is] = std::max(res, 1) * st->nsm;
        } else {
            PFX_CHECK(prn_configure_pass2(sd.ax.logL, sd.ax.pad, st->cpc, st->smem2));
        }
    }
    return PFX_OK;
}

void plan_free_pruned(pfx_plan *pl)
{
    PrunedState *st = state_of(pl);
    if (!st) return;
    if (st->fft_z_ok) cufftDestroy(st->fft_z);
    cudaStreamSynchronize(pl->stream);
    dct_free(st->dct);
    for (auto &pb : st->owned) {
        cudaFree(pb.first);
        pl->device_bytes -= pb.second;
    }
    delete st;
    pl->pruned = nullptr;
}

// Once per set of forward spectra: (Z_0, Z_1) = crop(IFFT2(u2*v2*(X_0 + i X_1)))/P.
int pruned_forward(pfx_plan *pl, const double *grid_dev, int slot)
{
    PrunedState *st = state_of(pl);
    if (st && st->dct) return dct_forward(pl, st->dct, grid_dev, slot);
    return forward_spectrum(pl, grid_dev, slot);
}

int pruned_prepare(pfx_plan *pl, int ngrids)
{
    PrunedState *st = state_of(pl);
    if (st->dct) return dct_prepare(pl, st->dct, ngrids, st->zc);
    const size_t P = (size_t)pl->NX * pl->NY;
    ProfScope prof(pl, PFX_KC_PREPARE, pl->stream);
    cufftHandle h = st->fft_z;
    PFX_CUFFT(cufftSetStream(h, pl->stream));
    for (int g = 0; g < ngrids; g++) {
        zmul_kernel<<<dim3((pl->NY + 255) / 256, pl->NX), 256, 0, pl->stream>>>(pl->NX, pl->NY, pl->d_spec[g], pl->d_u2,
                                                                                pl->d_v2, st->zplanes);
        PFX_CUDA(cudaGetLastError());
        PFX_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex *)st->zplanes, (cufftDoubleComplex *)st->zplanes, CUFFT_INVERSE));
        zcrop_kernel<<<dim3((pl->nx + 31) / 32, (pl->ny + 31) / 32), dim3(32, 8), 0, pl->stream>>>(
            pl->nx, pl->ny, pl->NY, 1.0 / (double)P, st->zplanes, st->zc, g, ngrids == 1);
        PFX_CUDA(cudaGetLastError());
    }
    pl->launches += 3 * ngrids;
    return PFX_OK;
}

namespace {
// kernel arguments of scale `is` with intermediate buffer tb
int scale_args(pfx_plan *pl, PrunedState *st, int is, int ngrids, int tb, PrunedArgs &a)
{
    if (!st) {
        set_error("pruned engine not initialised");
        return PFX_ERR_STATE;
    }
    a.sd = st->sd[is];
    a.nx = pl->nx;
    a.ny = pl->ny;
    a.NX = pl->NX;
    a.NY = pl->NY;
    a.ngrids = ngrids;
    a.spec0 = pl->d_spec[0];
    a.spec1 = pl->d_spec[1];
    if (st->dct) {
        a.dct0 = st->dct->D;
        a.dct1 = st->dct->D + (size_t)pl->nx * pl->ny;
        a.phx = st->dct->phx;
        a.phy = st->dct->phy;
    }
    a.tabu = st->tabu;
    a.tabv = st->tabv;
    a.twa = st->twa;
    a.twb = st->twb;
    a.rootx = st->rootx;
    a.rooty = st->rooty;
    a.T = st->Tbuf[tb & 1];
    a.W = st->W;
    if (st->v3x_nbuf[is]) {
        a.t_tr = 1;
        a.t_lp = v3_line_slots(a.sd.ax.logL);
        a.t_pad = v3_pad(a.sd.ax.logL);
    }
    return PFX_OK;
}
}  // namespace

int pruned_pass1(pfx_plan *pl, int is, int ngrids, int tb, cudaStream_t stream)
{
    PrunedState *st = state_of(pl);
    PrunedArgs a;
    PFX_CHECK(scale_args(pl, st, is, ngrids, tb, a));
    const ScaleDesc &sd = a.sd;
    const size_t sm1 = prn_smem_bytes(sd.ay, 1);
    dim3 g1(std::min(sd.kamax, st->ctas_per_transform), PFX_NA * ngrids);
    const int ntr = PFX_NA * ngrids;
    const char *ctas_env = getenv("PLATEFLEX_B200_CTAS");
    NvtxRange r("pfx:pass1");
    ProfScope prof(pl, PFX_KC_PASS1, stream);
    if (st->v4y_ctas[is]) {
        v4::V4Args va{a, sd.ay.logr, std::max(sd.ay.N >> 11, 1)};
        const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, 3 * st->v4y_ctas[is] / ntr);
        PFX_CHECK(v4_launch_p1(va, sd.ay.logL, sd.ay.N > 2048, st->dct != nullptr, dim3(std::min(sd.kamax, ctas), ntr), stream));
    } else if (st->v3y_nbuf[is]) {
        const int multi = (st->v2y[is] >> 5) & 1;
        v3::V3Args va{a, sd.ay.logr, std::max(sd.ay.N >> 11, 1), st->v3y_nbuf[is]};
        // a few waves: the number of columns differs between azimuths
        const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, 3 * st->v3y_ctas[is] / ntr);
        PFX_CHECK(v3_launch_p1(va, sd.ay.logL, multi, 2 * pl->ny == pl->NY, st->v3_occ, dim3(std::min(sd.kamax, ctas), ntr), stream));
    } else if (st->v2y[is] >= 0) {
        v2::V2Args va{a, sd.ay.logr, std::max(sd.ay.N >> 11, 1)};
        const int cpc = st->v2y[is] & 15, groups = (sd.kamax + cpc - 1) / cpc;
        // a few waves: the number of columns differs between azimuths
        const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, 3 * st->v2y_ctas[is] / ntr);
        PFX_CHECK(v2_launch_p1(va, sd.ay.logL, st->v2y[is], dim3(std::min(groups, ctas), ntr), stream));
    } else {
        PFX_CHECK(prn_launch_pass1(a, g1, sm1, stream));
    }
    pl->launches += 1;
    return PFX_OK;
}

int pruned_pass2(pfx_plan *pl, int is, int ngrids, int tb, cudaStream_t stream)
{
    PrunedState *st = state_of(pl);
    PrunedArgs a;
    PFX_CHECK(scale_args(pl, st, is, ngrids, tb, a));
    const ScaleDesc &sd = a.sd;
    const size_t sm2 = prn_smem_bytes(sd.ax, st->cpc);
    const int ngroups = (pl->ny + st->cpc - 1) / st->cpc;
    dim3 g2(std::min(ngroups, st->ctas_per_transform), PFX_NA * ngrids);
    const int ntr = PFX_NA * ngrids;
    const char *ctas_env = getenv("PLATEFLEX_B200_CTAS");
    {
        NvtxRange r("pfx:pass2");
        ProfScope prof(pl, PFX_KC_PASS2, stream);
        if (st->v4x_ctas[is]) {
            v4::V4Args va{a, sd.ax.logr, std::max(sd.ax.N >> 11, 1)};
            const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, st->v4_waves_p2 * st->v4x_ctas[is] / ntr);
            PFX_CHECK(v4_launch_p2(va, sd.ax.logL, sd.ax.N > 2048, dim3(std::min(pl->ny, ctas), ntr), stream));
        } else if (st->v3x_nbuf[is]) {
            const int multi = (st->v2x[is] >> 5) & 1;
            v3::V3Args va{a, sd.ax.logr, std::max(sd.ax.N >> 11, 1), st->v3x_nbuf[is]};
            // every transform has the same number of rows: one wave of persistent CTAs
            const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, st->v3x_ctas[is] / ntr);
            PFX_CHECK(v3_launch_p2(va, sd.ax.logL, multi, 2 * pl->nx == pl->NX, st->v3_occ, dim3(std::min(pl->ny, ctas), ntr), stream));
        } else if (st->v2hx_ctas[is]) {
            v2h::V2Args va{a, sd.ax.logr, std::max(sd.ax.N >> 10, 1)};
            const int ctas = ctas_env ? st->ctas_per_transform : std::max(1, st->v2hx_ctas[is] / ntr);
            PFX_CHECK(v2h_launch_p2(va, sd.ax.logL, sd.ax.N > 1024, dim3(std::min(pl->ny, ctas), ntr), stream));
        } else if (st->v2x[is] >= 0) {
            v2::V2Args va{a, sd.ax.logr, std::max(sd.ax.N >> 11, 1)};
            const int cpc = st->v2x[is] & 15, groups = (pl->ny + cpc - 1) / cpc;
            // every transform has the same number of rows: one wave of persistent CTAs when a line is a single
            // part; with several parts per line (N > 2048, two CTAs per SM) 192 CTAs per transform measured 7 %
            // faster than one wave (the rows of a CTA stay closer together, the tail is finer)
            const bool multi = (st->v2x[is] >> 5) & 1;
            const int ctas = ctas_env ? st->ctas_per_transform : (multi ? 192 : std::max(1, st->v2x_ctas[is] / ntr));
            PFX_CHECK(v2_launch_p2(va, sd.ax.logL, st->v2x[is], dim3(std::min(groups, ctas), ntr), stream));
        } else {
            PFX_CHECK(prn_launch_pass2(a, st->cpc, g2, sm2, stream));
        }
    }
    for (int ia = 0; ia < PFX_NA; ia++) st->cur_ce[ia] = sd.ce[ia];
    pl->launches += 1;
    return PFX_OK;
}

int pruned_scale_planes(pfx_plan *pl, int is, int ngrids)
{
    PFX_CHECK(pruned_pass1(pl, is, ngrids, 0, pl->stream));
    return pruned_pass2(pl, is, ngrids, 0, pl->stream);
}

// Relative cost model of one scale (arbitrary units ~ microseconds on B200), from the measured launch times of the
// 4096^2 configuration (profiles/r02_launch_summary.txt, radix-16 pass 2): pass 2 scales with rows x line length
// times a factor per FFT length (1.96 ms at L = 32 ... 4.58 ms at L = 2048 for 4096 x 8192 x 22), pass 1 with
// columns x line length x log2(FFT length), the reduction with the number of cells.  Used to balance scales over
// ranks (sharding.deal_scales / balanced_cuts).
double pruned_scale_cost(const pfx_plan *pl, int is)
{
    const PrunedState *st = reinterpret_cast<const PrunedState *>(pl->pruned);
    const ScaleDesc &sd = st->sd[is];
    static const double p2_ms[7] = {1.96, 2.00, 2.17, 2.69, 2.93, 3.41, 4.58};  // logL = 5 .. 11
    const int l = sd.ax.logL;
    const double per = l < 5 ? p2_ms[0] : (l > 11 ? p2_ms[6] * (double)l / 11.0 : p2_ms[l - 5]);
    const double p2 = per * 1e3 * ((double)pl->ny * pl->NX) / (4096.0 * 8192.0);
    const double p1 = 1.4e-5 * (double)sd.kamax * pl->NY * sd.ay.logL;
    return p1 + p2 + 0.8e-4 * (double)pl->nx * pl->ny;
}

int pruned_num_buffers(const pfx_plan *pl) { return pl->pruned ? 2 : 1; }

void pruned_select_buffer(pfx_plan *pl, int b) { state_of(pl)->W = state_of(pl)->Wbuf[b & 1]; }

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
    v.zc = st->zc;  // W = main term - c*E*Z_g (cpwt_sub.f90:93-94), Z real: zc[j][i] = (Z_0, Z_1)
    v.zsel = g;
    for (int ia = 0; ia < PFX_NA; ia++) v.ce[ia] = st->cur_ce[ia];
    return v;
}

}  // namespace pfx
