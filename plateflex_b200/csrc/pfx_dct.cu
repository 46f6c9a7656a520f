// pfx_dct.cu -- forward spectra and the admissibility transform through cosine transforms of the UNMIRRORED grid.
//
// What it replaces: mirror_data + fourn forward (cpwt.f90:100-122, cpwt_sub.f90:435-449) and the inverse transform
// of the admissibility term (cpwt_sub.f90:93) -- on power-of-two grids, where the padded array IS the four-fold
// even reflection (NX = 2 nx, NY = 2 ny, no zero padding).  There the forward sum over the mirrored array is
//
//     X[a][b] = sum_{i<2nx, j<2ny} m[i][j] e^{-2 pi i (ia/NX + jb/NY)} = e^{i pi a/NX} e^{i pi b/NY} D[|a|][|b|],
//     D[a][b] = sum_{i<nx, j<ny} x[i][j] 2cos(pi a (2i+1)/NX) 2cos(pi b (2j+1)/NY)        (real, D[nx][.] = D[.][ny] = 0)
//
// i.e. a two-dimensional DCT-II of the nx x ny grid, which (Makhoul 1980) is ONE nx x ny real-to-complex FFT of the
// grid re-ordered even-samples-first / odd-samples-reversed, plus a twiddle pass:
//     D[a][b] = 2 Re( W_a ( U_b V[a][b] + conj(U_b) V[a][ny-b] ) ),   W_a = e^{-i pi a/(2nx)},  U_b = e^{-i pi b/(2ny)}.
// That is 1/16 of the 2nx x 2ny complex transform in both work and bytes, and the stored "spectrum" is nx*ny reals
// (134 MB at 4096^2 instead of 1.07 GB) which pass 1 of the pruned engine turns back into X[a][b] on the fly
// (pfx_pruned_v2_body.inc / pfx_pruned_v4.cuh, Pass1Group).  The admissibility plane Z = IFFT2(G X)/P is again an
// even-symmetric real field, the DCT-III of G D, computed as one complex-to-real nx x ny FFT per grid of
//     Q[a][b] = conj(W_a) ( B[a][b] - i B[nx-a][b] ),  B[a][b] = conj(U_b) ( E[a][b] - i E[a][ny-b] ),  E = G D,
// followed by the inverse re-ordering; the reference's one-sample shift reads sample nx as its mirror nx-1.
// Checked against the dense path by the oracle parity tests (tests/test_gpu_parity.py) -- same mathematics,
// different rounding (agreement ~1e-15 of the plane maximum).
#include "pfx_dct.hpp"

#include <cstdlib>
#include <cstring>

namespace pfx {

namespace {

// e^{i pi a / N} for the signed index a of slot idx (idx <= N/2 ? idx : idx - N)
__global__ void phase_kernel(int N, double2 *ph)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const int a = (idx <= N / 2) ? idx : idx - N;
    double s, c;
    sincospi((double)a / (double)N, &s, &c);
    ph[idx] = make_double2(c, s);
}

// e^{-i pi a / (2n)} for a = 0 .. n
__global__ void half_kernel(int n, double2 *h)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n) return;
    double s, c;
    sincospi(-(double)a / (double)(2 * n), &s, &c);
    h[a] = make_double2(c, s);
}

__device__ __forceinline__ int makhoul_src(int p, int n) { return (p < n / 2) ? 2 * p : 2 * (n - 1 - p) + 1; }
__device__ __forceinline__ int makhoul_dst(int i, int n) { return (i & 1) ? n - 1 - (i >> 1) : (i >> 1); }

// v[i'][j'] = x[src(i')][src(j')] - mean   (y fastest, as numpy hands the grid)
__global__ void __launch_bounds__(256) reorder_kernel(const double *__restrict__ grid, int nx, int ny,
                                                      const double *__restrict__ mean_ptr, double *__restrict__ v)
{
    const int jp = blockIdx.x * blockDim.x + threadIdx.x, ip = blockIdx.y;
    if (jp >= ny) return;
    v[(size_t)ip * ny + jp] = grid[(size_t)makhoul_src(ip, nx) * ny + makhoul_src(jp, ny)] - *mean_ptr;
}

__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// D[a][b] from the half spectrum V[a][b <= ny/2] of the re-ordered grid
__global__ void __launch_bounds__(256) dct_post_kernel(const double2 *__restrict__ V, int nx, int ny,
                                                       const double2 *__restrict__ hx, const double2 *__restrict__ hy,
                                                       double *__restrict__ D)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= ny) return;
    const int nyh = ny / 2 + 1, am = (nx - a) % nx, bm = (ny - b) % ny;
    auto vfull = [&](int bb) -> double2 {
        if (bb <= ny / 2) return V[(size_t)a * nyh + bb];
        const double2 c = V[(size_t)am * nyh + (ny - bb)];
        return make_double2(c.x, -c.y);
    };
    const double2 U = hy[b], W = hx[a];
    const double2 p = cmulc(U, vfull(b)), q = cmulc(make_double2(U.x, -U.y), vfull(bm));
    const double2 s = make_double2(p.x + q.x, p.y + q.y);
    D[(size_t)a * ny + b] = 2.0 * (W.x * s.x - W.y * s.y);
}

// Q[a][b <= ny/2] of one grid (see the header comment)
__global__ void __launch_bounds__(256) dct_q_kernel(const double *__restrict__ D, int nx, int ny,
                                                    const double *__restrict__ u2, const double *__restrict__ v2,
                                                    const double2 *__restrict__ hx, const double2 *__restrict__ hy,
                                                    double2 *__restrict__ Q)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    const int nyh = ny / 2 + 1;
    if (b >= nyh) return;
    auto E = [&](int aa, int bb) -> double {
        return (aa < nx && bb < ny) ? (u2[aa] * v2[bb]) * D[(size_t)aa * ny + bb] : 0.0;
    };
    const double2 Uc = make_double2(hy[b].x, -hy[b].y), Wc = make_double2(hx[a].x, -hx[a].y);
    // B[aa][b] = conj(U_b) (E[aa][b] - i E[aa][ny-b])
    const double2 B0 = cmulc(Uc, make_double2(E(a, b), -E(a, ny - b)));
    const double2 B1 = cmulc(Uc, make_double2(E(nx - a, b), -E(nx - a, ny - b)));
    // B0 - i B1
    Q[(size_t)a * nyh + b] = cmulc(Wc, make_double2(B0.x + B1.y, B0.y - B1.x));
}

// zc[j][i] = (Z_0, Z_1)(i, j) = scale * y_g[min(i+1, nx-1)][min(j+1, ny-1)],  y_g = inverse re-ordering of the
// complex-to-real output r_g; transposed to the x-fastest layout of the coefficient planes through a 32 x 32 tile
__global__ void __launch_bounds__(256) dct_zcrop_kernel(int nx, int ny, double scale, const double *__restrict__ r0,
                                                        const double *__restrict__ r1, double2 *__restrict__ zc)
{
    __shared__ double2 tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        if (i < nx && j < ny) {
            const int i1 = min(i + 1, nx - 1), j1 = min(j + 1, ny - 1);
            const size_t o = (size_t)makhoul_dst(i1, nx) * ny + makhoul_dst(j1, ny);
            tile[r][threadIdx.x] = make_double2(r0[o], r1 ? r1[o] : 0.0);
        }
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int i = i0 + threadIdx.x, j = j0 + c;
        if (i < nx && j < ny) {
            const double2 v = tile[threadIdx.x][c];
            zc[(size_t)j * nx + i] = make_double2(v.x * scale, v.y * scale);
        }
    }
}

}  // namespace

bool dct_applicable(const pfx_plan *pl)
{
    const char *e = getenv("PLATEFLEX_B200_DCT");
    if (e && !strcmp(e, "off")) return false;
    return pl->NX == 2 * pl->nx && pl->NY == 2 * pl->ny && pl->nx >= 4 && pl->ny >= 4;
}

int dct_init(pfx_plan *pl, DctState **out, std::vector<std::pair<void *, size_t>> &owned)
{
    *out = nullptr;
    DctState *st = new DctState();
    const int nx = pl->nx, ny = pl->ny, nyh = ny / 2 + 1;
    const size_t nxy = (size_t)nx * ny;
    auto dalloc = [&](void **p, size_t bytes) -> int {
        *p = nullptr;
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cosine-transform path: device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return PFX_ERR_OOM;
        }
        owned.emplace_back(*p, bytes);
        pl->device_bytes += bytes;
        return PFX_OK;
    };
    int rc = PFX_OK;
    auto fail = [&](int r) {
        dct_free(st);
        return r;
    };
    if ((rc = dalloc((void **)&st->v, 2 * nxy * sizeof(double)))) return fail(rc);
    if ((rc = dalloc((void **)&st->V, 2 * (size_t)nx * nyh * sizeof(double2)))) return fail(rc);
    if ((rc = dalloc((void **)&st->D, 2 * nxy * sizeof(double)))) return fail(rc);
    if ((rc = dalloc((void **)&st->phx, (size_t)pl->NX * sizeof(double2)))) return fail(rc);
    if ((rc = dalloc((void **)&st->phy, (size_t)pl->NY * sizeof(double2)))) return fail(rc);
    if ((rc = dalloc((void **)&st->hx, (size_t)(nx + 1) * sizeof(double2)))) return fail(rc);
    if ((rc = dalloc((void **)&st->hy, (size_t)(ny + 1) * sizeof(double2)))) return fail(rc);
    phase_kernel<<<(pl->NX + 255) / 256, 256, 0, pl->stream>>>(pl->NX, st->phx);
    phase_kernel<<<(pl->NY + 255) / 256, 256, 0, pl->stream>>>(pl->NY, st->phy);
    half_kernel<<<(nx + 256) / 256, 256, 0, pl->stream>>>(nx, st->hx);
    half_kernel<<<(ny + 256) / 256, 256, 0, pl->stream>>>(ny, st->hy);
    if (cudaGetLastError() != cudaSuccess) return fail(PFX_ERR_CUDA);
    int n[2] = {nx, ny};
    cufftResult r = cufftCreate(&st->fwd);
    if (r == CUFFT_SUCCESS) {
        st->fwd_ok = true;
        r = cufftPlanMany(&st->fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 1);
    }
    if (r == CUFFT_SUCCESS) r = cufftCreate(&st->inv);
    if (r == CUFFT_SUCCESS) {
        st->inv_ok = true;
        r = cufftPlanMany(&st->inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 1);
    }
    if (r != CUFFT_SUCCESS) {
        set_error("cosine-transform path: cufft plan %dx%d failed: %d", nx, ny, (int)r);
        return fail(r == CUFFT_ALLOC_FAILED ? PFX_ERR_OOM : PFX_ERR_CUFFT);
    }
    *out = st;
    return PFX_OK;
}

void dct_free(DctState *st)
{
    if (!st) return;
    if (st->fwd_ok) cufftDestroy(st->fwd);
    if (st->inv_ok) cufftDestroy(st->inv);
    delete st;  // device buffers belong to the pruned engine's `owned` list
}

int dct_forward(pfx_plan *pl, DctState *st, const double *grid_dev, int slot)
{
    const int nx = pl->nx, ny = pl->ny, nyh = ny / 2 + 1;
    const size_t nxy = (size_t)nx * ny;
    ProfScope prof(pl, PFX_KC_FORWARD, pl->stream);
    const double *mean = nullptr;
    PFX_CHECK(grid_mean(pl, grid_dev, &mean));
    double *v = st->v + (size_t)slot * nxy;
    double2 *V = st->V + (size_t)slot * nx * nyh;
    reorder_kernel<<<dim3((ny + 255) / 256, nx), 256, 0, pl->stream>>>(grid_dev, nx, ny, mean, v);
    PFX_CUDA(cudaGetLastError());
    PFX_CUFFT(cufftSetStream(st->fwd, pl->stream));
    PFX_CUFFT(cufftExecD2Z(st->fwd, v, (cufftDoubleComplex *)V));
    dct_post_kernel<<<dim3((ny + 255) / 256, nx), 256, 0, pl->stream>>>(V, nx, ny, st->hx, st->hy, st->D + (size_t)slot * nxy);
    PFX_CUDA(cudaGetLastError());
    pl->launches += 5;
    return PFX_OK;
}

int dct_prepare(pfx_plan *pl, DctState *st, int ngrids, double2 *zc)
{
    const int nx = pl->nx, ny = pl->ny, nyh = ny / 2 + 1;
    const size_t nxy = (size_t)nx * ny;
    ProfScope prof(pl, PFX_KC_PREPARE, pl->stream);
    PFX_CUFFT(cufftSetStream(st->inv, pl->stream));
    for (int g = 0; g < ngrids; g++) {
        double2 *Q = st->V + (size_t)g * nx * nyh;
        dct_q_kernel<<<dim3((nyh + 255) / 256, nx), 256, 0, pl->stream>>>(st->D + (size_t)g * nxy, nx, ny, pl->d_u2, pl->d_v2,
                                                                          st->hx, st->hy, Q);
        PFX_CUDA(cudaGetLastError());
        PFX_CUFFT(cufftExecZ2D(st->inv, (cufftDoubleComplex *)Q, st->v + (size_t)g * nxy));
    }
    const double P = (double)pl->NX * (double)pl->NY;
    dct_zcrop_kernel<<<dim3((nx + 31) / 32, (ny + 31) / 32), dim3(32, 8), 0, pl->stream>>>(
        nx, ny, 1.0 / P, st->v, ngrids == 2 ? st->v + nxy : nullptr, zc);
    PFX_CUDA(cudaGetLastError());
    pl->launches += 2 * ngrids + 1;
    return PFX_OK;
}

}  // namespace pfx
