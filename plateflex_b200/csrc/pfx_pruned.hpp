// pfx_pruned.hpp -- descriptors shared by the host planner (pfx_cwt_pruned.cu) and the
// specialised kernels (pfx_pruned_kernels.cuh) of the band-pruned transform engine.
#pragma once
#include "pfx_common.cuh"

namespace pfx {

constexpr int PRN_MAX_STAGES = 6;
constexpr int PRN_NT = 256;       // threads per CTA
constexpr int PRN_MAX_LOGL = 14;  // FFT lengths up to 16384

// Radix plan of a length-2^logl FFT (log2 of each stage's radix, first executed first): radix 8
// wherever possible; a remainder of one bit turns the last 8 into 4*4, a remainder of two bits
// appends a radix-4 stage.  Radix 8 keeps the kernels near 80 registers (3 CTAs of 256 threads per SM).
__host__ __device__ constexpr int prn_nstages(int logl)
{
    return (logl <= 3) ? 1 : ((logl % 3 == 0) ? logl / 3 : logl / 3 + 1);
}
__host__ __device__ constexpr int prn_lrad(int logl, int s)
{
    if (logl <= 3) return s == 0 ? logl : 0;
    const int q = logl / 3, rem = logl % 3, n = prn_nstages(logl);
    if (s >= n) return 0;
    if (rem == 0) return 3;
    if (rem == 2) return s < q ? 3 : 2;
    return s < q - 1 ? 3 : 2;  // rem == 1: (q-1) eights, then 4, 4
}
__host__ __device__ constexpr int prn_logm(int logl, int s)  // log2 of the sub-FFT size entering stage s
{
    int m = 0;
    for (int i = 0; i < s; i++) m += prn_lrad(logl, i);
    return m;
}
__host__ __device__ constexpr int prn_stw_off(int logl, int s)  // offset of stage s's twiddles in shared memory
{
    int off = 0;
    for (int i = 1; i < s; i++) off += 1 << prn_logm(logl, i);
    return off;
}
__host__ __device__ constexpr int prn_stw_total(int logl) { return prn_stw_off(logl, prn_nstages(logl)); }

struct AxisDesc {
    int N, logN;    // padded length of the axis
    int L, logL;    // FFT length (power of two >= support width)
    int r, logr;    // residues per line, N / L
    int rc, logrc;  // residues handled per pass over shared memory (r / nsplit)
    int nout;       // outputs needed along this axis (nx or ny)
    int pad;        // 1: element padding e + (e >> 3) (used when rc < 8 and L >= 256)
};

struct ScaleDesc {
    AxisDesc ax, ay;
    int alo[PFX_NA], ka[PFX_NA];  // support [alo, alo+ka) along x per azimuth (signed wavenumber index)
    int blo[PFX_NA], kb[PFX_NA];
    double ce[PFX_NA];            // norm * exp(-0.5*(kx0^2+ky0^2)): factor of the Z correction
    int kamax;                    // rows of T per transform
    size_t tabu, tabv;            // offsets (doubles) of the weight tables of this scale
    size_t twa, twb;              // offsets (double2) of the stage-1 twiddle tables
};

struct PrunedArgs {
    ScaleDesc sd;
    int nx, ny, NX, NY, ngrids;
    const double2 *spec0, *spec1;  // padded spectra [a][b] (y-wavenumber fastest)
    const double *tabu, *tabv;     // weight tables (all scales)
    const double2 *twa, *twb;      // stage-1 twiddle tables (all scales)
    const double2 *rootx, *rooty;  // exp(2 pi i m / N)
    double2 *T;                    // intermediate, t = ia + 11*g: [t][ka row][j], or with t_tr [t][j][slot]
    // t_tr: pass 1 writes T transposed, one contiguous line of t_lp slots per output row j, the sample of
    // wavenumber index a at slot n + (t_pad ? n >> 3 : 0), n = a mod Lx -- the order in which the first stage
    // of the third-generation pass 2 (pfx_pruned_v3.cuh) reads it after one bulk copy of the line
    int t_tr = 0, t_lp = 0, t_pad = 0;
    double2 *W;                    // [t][j][i] output planes (main Gaussian term only)
    // Cosine-transform form of the spectra (pfx_dct.cu; power-of-two grids, NX = 2 nx, NY = 2 ny): when dct0 is set
    // pass 1 reads  X_g[a][b] = phx[a] phy[b] D_g[|a|][|b|]  (D real, [nx][ny], zero at |a| = nx or |b| = ny)
    // instead of spec0 / spec1; the x phase is applied with the u1 weight when the column is stored.
    const double *dct0 = nullptr, *dct1 = nullptr;
    const double2 *phx = nullptr, *phy = nullptr;
};

inline size_t prn_plane_elems(const AxisDesc &A) { return (size_t)((A.L - 1) + (A.pad ? (A.L - 1) >> 3 : 0) + 1) * A.rc; }
inline size_t prn_smem_bytes(const AxisDesc &A, int cpc)
{
    return (prn_plane_elems(A) * cpc + (A.r > 1 ? (size_t)A.L * cpc : 0) + prn_stw_total(A.logL) + 2 * (size_t)A.rc) *
           sizeof(double2);
}

// element offset in T of output j of support column m (wavenumber index alo + m) of transform t, split into
// the part that is fixed for a column and the stride between consecutive j
__host__ __device__ inline size_t prn_t_column(const PrunedArgs &a, int t, int m, int alo, size_t &jstride)
{
    if (a.t_tr) {
        const int n = (alo + m) & (a.sd.ax.L - 1);
        jstride = (size_t)a.t_lp;
        return (size_t)t * a.ny * a.t_lp + (size_t)(n + (a.t_pad ? (n >> 3) : 0));
    }
    jstride = 1;
    return ((size_t)t * a.sd.kamax + m) * a.ny;
}

// launchers, one translation unit each (pfx_pruned_p1.cu, pfx_pruned_p2.cu)
int prn_launch_pass1(const PrunedArgs &a, dim3 grid, size_t smem, cudaStream_t st);
int prn_launch_pass2(const PrunedArgs &a, int cpc, dim3 grid, size_t smem, cudaStream_t st);
int prn_configure_pass1(int logL, int pad, size_t smem);
int prn_configure_pass2(int logL, int pad, int cpc, size_t smem);

}  // namespace pfx
