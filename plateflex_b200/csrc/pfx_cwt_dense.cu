// pfx_cwt_dense.cu -- K1 (mirror / de-mean / pad + forward FFT) and the dense
// CWT engine: K2 daughter-times-spectrum for all 11 azimuths of a scale, then
// one batched cuFFT Z2Z inverse over the 11 (or 22) planes.
//
// Device layout: every padded 2-D array is stored x-fastest, i.e. element
// (i, j) of the padded grid / (a, b) of the spectrum sits at [j*NX + i] /
// [b*NX + a] -- the reference's own column-major layout (cpwt.f90:115).
#include "pfx_plan.hpp"

namespace pfx {

namespace {

// ---- deterministic two-stage sum (for sum(grid_mirror), cpwt.f90:111) -------------
constexpr int SUM_BLOCKS = 296;  // 2 per SM
constexpr int SUM_THREADS = 256;

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(SUM_THREADS) sum_partial_kernel(const double *__restrict__ g, size_t n, double *partial)
{
    double v = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v += g[i];
    v = block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

__global__ void __launch_bounds__(SUM_THREADS) sum_final_kernel(double *partial, int np, double inv_n)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) v += partial[i];
    v = block_sum(v);
    if (threadIdx.x == 0) partial[np] = v * inv_n;  // the mean, stored after the partials
}

// ---- K1: mirror + de-mean + pad, tiled transpose from the C-order input -----------
// cpwt_sub.f90:435-449 (mirror_data) and cpwt.f90:110-116.  Each 32x32 source tile is
// read once (coalesced along y, the input's fast axis) and written to its four mirror
// images (coalesced along x, the device layout's fast axis).
__global__ void __launch_bounds__(256)
    mirror_pad_kernel(const double *__restrict__ grid, int nx, int ny, int NX, const double *__restrict__ mean_ptr,
                      double2 *__restrict__ xpad)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int ii = i0 + r, jj = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (ii < nx && jj < ny) ? grid[(size_t)ii * ny + jj] : 0.0;
    }
    __syncthreads();
    const double mean = *mean_ptr;
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int ii = i0 + threadIdx.x, jj = j0 + c;
        if (ii < nx && jj < ny) {
            const double2 v = make_double2(tile[threadIdx.x][c] - mean, 0.0);
            const int im = 2 * nx - 1 - ii, jm = 2 * ny - 1 - jj;
            xpad[(size_t)jj * NX + ii] = v;
            xpad[(size_t)jj * NX + im] = v;
            xpad[(size_t)jm * NX + ii] = v;
            xpad[(size_t)jm * NX + im] = v;
        }
    }
}

// Same for the y-fastest device layout of the pruned engine (element (i, j) at [i*NY + j], spectrum
// (a, b) at [a*NY + b]): input and output share the fast axis, no transposition.
__global__ void __launch_bounds__(256)
    mirror_pad_t_kernel(const double *__restrict__ grid, int nx, int ny, int NY, const double *__restrict__ mean_ptr,
                        double2 *__restrict__ xpad)
{
    const int jj = blockIdx.x * blockDim.x + threadIdx.x, ii = blockIdx.y;
    if (jj >= ny) return;
    const double2 v = make_double2(grid[(size_t)ii * ny + jj] - *mean_ptr, 0.0);
    const int im = 2 * nx - 1 - ii, jm = 2 * ny - 1 - jj;
    xpad[(size_t)ii * NY + jj] = v;
    xpad[(size_t)ii * NY + jm] = v;
    xpad[(size_t)im * NY + jj] = v;
    xpad[(size_t)im * NY + jm] = v;
}

// ---- wavenumber axes and the scale-independent admissibility factors --------------
// defk (cpwt_sub.f90:41-62): index a <= n/2 -> +a*dk (Nyquist positive), else (a-n)*dk.
__global__ void axes_kernel(int n, double d_km, double *__restrict__ kk, double *__restrict__ g2)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double pi = 3.141592653589793;
    const double dk = 2 * pi / ((double)n * d_km);
    const double k = (a <= n / 2) ? (double)a * dk : -((double)(n - a) * dk);
    kk[a] = k;
    g2[a] = exp(-0.5 * (k * k));
}

// ---- dense engine ------------------------------------------------------------------
// Separable factors of the main Gaussian of the daughter (cpwt_sub.f90:90-92):
//   u1[ia][a] = exp(-0.5*(s*kx[a] - kx0[ia])^2),  v1[ia][b] = exp(-0.5*(s*ky[b] - ky0[ia])^2)
__global__ void dense_tables_kernel(int NX, int NY, const double *__restrict__ kx, const double *__restrict__ ky,
                                    PfxScaleConsts c, double *__restrict__ tab)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = blockIdx.y;
    if (t < NX) {
        const double d = c.s * kx[t] - c.kx0[ia];
        tab[(size_t)ia * NX + t] = exp(-0.5 * (d * d));
    } else if (t < NX + NY) {
        const int b = t - NX;
        const double d = c.s * ky[b] - c.ky0[ia];
        tab[(size_t)PFX_NA * NX + (size_t)ia * NY + b] = exp(-0.5 * (d * d));
    }
}

// K2: planes[ia + 11*g][b][a] = psi_ia(a, b) * X_g[b][a]   (cpwt_sub.f90:92-94, cpwt.f90:169;
// the reference forms daughter*CONJG(ft) with ft = sum x e^{+i..}; for real x that is psi * FFT2(x)).
template <int NG>
__global__ void __launch_bounds__(256)
    dense_multiply_kernel(int NX, int NY, const double2 *__restrict__ spec0, const double2 *__restrict__ spec1,
                          const double *__restrict__ tab, const double *__restrict__ u2, const double *__restrict__ v2,
                          PfxScaleConsts c, double2 *__restrict__ planes)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= NX) return;
    const size_t P = (size_t)NX * NY, o = (size_t)b * NX + a;
    const double2 x0 = spec0[o];
    double2 x1 = make_double2(0.0, 0.0);
    if (NG == 2) x1 = spec1[o];
    const double g = u2[a] * v2[b];
    const double *u1 = tab, *v1 = tab + (size_t)PFX_NA * NX;
#pragma unroll
    for (int ia = 0; ia < PFX_NA; ia++) {
        const double psi = c.norm * u1[(size_t)ia * NX + a] * v1[(size_t)ia * NY + b] - c.norm * (g * c.e0[ia]);
        planes[(size_t)ia * P + o] = make_double2(psi * x0.x, psi * x0.y);
        if (NG == 2) planes[(size_t)(ia + PFX_NA) * P + o] = make_double2(psi * x1.x, psi * x1.y);
    }
}

}  // namespace

// mean of the nx*ny grid (deterministic two-stage sum), left on the device
int grid_mean(pfx_plan *pl, const double *grid_dev, const double **mean_dev)
{
    const size_t n = (size_t)pl->nx * pl->ny;
    sum_partial_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, pl->stream>>>(grid_dev, n, pl->d_partial);
    sum_final_kernel<<<1, SUM_THREADS, 0, pl->stream>>>(pl->d_partial, SUM_BLOCKS, 1.0 / (double)n);
    PFX_CUDA(cudaGetLastError());
    *mean_dev = pl->d_partial + SUM_BLOCKS;
    return PFX_OK;
}

int forward_spectrum(pfx_plan *pl, const double *grid_dev, int slot)
{
    const size_t n = (size_t)pl->nx * pl->ny, P = (size_t)pl->NX * pl->NY;
    ProfScope prof(pl, PFX_KC_FORWARD, pl->stream);
    sum_partial_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, pl->stream>>>(grid_dev, n, pl->d_partial);
    sum_final_kernel<<<1, SUM_THREADS, 0, pl->stream>>>(pl->d_partial, SUM_BLOCKS, 1.0 / (double)n);
    double2 *xp = pl->d_spec[slot];
    if (pl->NX != 2 * pl->nx || pl->NY != 2 * pl->ny) {
        PFX_CUDA(cudaMemsetAsync(xp, 0, P * sizeof(double2), pl->stream));
        pl->launches++;
    }
    if (pl->spec_t) {
        mirror_pad_t_kernel<<<dim3((pl->ny + 255) / 256, pl->nx), 256, 0, pl->stream>>>(
            grid_dev, pl->nx, pl->ny, pl->NY, pl->d_partial + SUM_BLOCKS, xp);
    } else {
        dim3 grid((pl->nx + 31) / 32, (pl->ny + 31) / 32), block(32, 8);
        mirror_pad_kernel<<<grid, block, 0, pl->stream>>>(grid_dev, pl->nx, pl->ny, pl->NX, pl->d_partial + SUM_BLOCKS, xp);
    }
    PFX_CUDA(cudaGetLastError());
    PFX_CUFFT(cufftSetStream(pl->fft_fwd, pl->stream));
    PFX_CUFFT(cufftExecZ2Z(pl->fft_fwd, (cufftDoubleComplex *)xp, (cufftDoubleComplex *)xp, CUFFT_FORWARD));
    pl->launches += 4;
    return PFX_OK;
}

int plan_init_axes(pfx_plan *pl)
{
    axes_kernel<<<(pl->NX + 255) / 256, 256, 0, pl->stream>>>(pl->NX, pl->dx, pl->d_kx, pl->d_u2);
    axes_kernel<<<(pl->NY + 255) / 256, 256, 0, pl->stream>>>(pl->NY, pl->dy, pl->d_ky, pl->d_v2);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

int plan_init_dense(pfx_plan *pl)
{
    const size_t P = (size_t)pl->NX * pl->NY;
    PFX_CHECK(pl->alloc((void **)&pl->d_tab, (size_t)PFX_NA * (pl->NX + pl->NY) * sizeof(double)));
    PFX_CHECK(pl->alloc((void **)&pl->d_planes, 2 * PFX_NA * P * sizeof(double2)));
    int n[2] = {pl->NY, pl->NX};
    size_t need = 0;
    for (int g = 0; g < 2; g++) {
        size_t ws = 0;
        PFX_CUFFT(cufftCreate(&pl->fft_inv[g]));
        pl->fft_inv_ok[g] = true;
        PFX_CUFFT(cufftSetAutoAllocation(pl->fft_inv[g], 0));
        PFX_CUFFT(cufftMakePlanMany(pl->fft_inv[g], 2, n, nullptr, 1, (int)P, nullptr, 1, (int)P, CUFFT_Z2Z,
                                    PFX_NA * (g + 1), &ws));
        if (ws > need) need = ws;
    }
    if (need > pl->fft_work_bytes) {
        // the forward plan was given a smaller area; re-point every plan at one shared area
        void *w = nullptr;
        PFX_CHECK(pl->alloc(&w, need));
        pl->d_fft_work = w;
        pl->fft_work_bytes = need;
        PFX_CUFFT(cufftSetWorkArea(pl->fft_fwd, pl->d_fft_work));
    }
    for (int g = 0; g < 2; g++) PFX_CUFFT(cufftSetWorkArea(pl->fft_inv[g], pl->d_fft_work));
    return PFX_OK;
}

int dense_scale_planes(pfx_plan *pl, int is, int ngrids)
{
    const PfxScaleConsts &c = pl->consts[is];
    dim3 tg((pl->NX + pl->NY + 255) / 256, PFX_NA);
    {
        ProfScope prof(pl, PFX_KC_DENSE_MUL, pl->stream);
        dense_tables_kernel<<<tg, 256, 0, pl->stream>>>(pl->NX, pl->NY, pl->d_kx, pl->d_ky, c, pl->d_tab);
        dim3 mg((pl->NX + 255) / 256, pl->NY);
        if (ngrids == 2)
            dense_multiply_kernel<2><<<mg, 256, 0, pl->stream>>>(pl->NX, pl->NY, pl->d_spec[0], pl->d_spec[1],
                                                                 pl->d_tab, pl->d_u2, pl->d_v2, c, pl->d_planes);
        else
            dense_multiply_kernel<1><<<mg, 256, 0, pl->stream>>>(pl->NX, pl->NY, pl->d_spec[0], nullptr, pl->d_tab,
                                                                 pl->d_u2, pl->d_v2, c, pl->d_planes);
    }
    PFX_CUDA(cudaGetLastError());
    cufftHandle h = pl->fft_inv[ngrids - 1];
    PFX_CUFFT(cufftSetStream(h, pl->stream));
    ProfScope prof(pl, PFX_KC_DENSE_FFT, pl->stream);
    PFX_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex *)pl->d_planes, (cufftDoubleComplex *)pl->d_planes, CUFFT_INVERSE));
    pl->launches += 3;
    return PFX_OK;
}

PlaneView dense_view(const pfx_plan *pl, int g)
{
    // W[m,n] = IFFT2(psi*X)[m+1, n+1] / (nnx*nny): the +1 is the net effect of the reference's
    // fourn sign conventions and its full-reversal cfftswap (cpwt.f90:173-185, cpwt_sub.f90:457-469).
    const size_t P = (size_t)pl->NX * pl->NY;
    PlaneView v;
    v.base = pl->d_planes + (size_t)g * PFX_NA * P;
    v.plane_stride = P;
    v.ld = pl->NX;
    v.off_i = 1;
    v.off_j = 1;
    v.scale = 1.0 / (double)P;
    v.zc = nullptr;
    v.zsel = 0;
    for (int ia = 0; ia < PFX_NA; ia++) v.ce[ia] = 0.0;
    return v;
}

}  // namespace pfx
