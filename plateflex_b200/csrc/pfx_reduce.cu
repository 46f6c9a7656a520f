// pfx_reduce.cu -- K3: azimuth reduction, admittance / coherence and the
// running-prefix jackknife, one thread per grid cell, all 11 azimuths of the
// cell held in registers.  Replaces the full-cube passes of
//   cpwt.f90:200-241 (wlet_scalogram), 250-293 (wlet_xscalogram), 303-364
//   (wlet_admit_coh) and cpwt_sub.f90:107-160, 298-366 (the jackknives)
// without ever materialising the (nx,ny,na,ns) temporaries sg/sg1/sg2/xsg.
//
// HBM traffic per cell and scale: 11 (or 22) coalesced 16-byte loads and 2-4
// 8-byte stores; consecutive threads own consecutive x so every warp reads and
// writes whole 128-byte lines.
#include <cstdlib>

#include "pfx_common.cuh"

namespace pfx {

namespace {

constexpr int NA = PFX_NA;

// 1/x to ~1 ulp without the special-case branches of the IEEE division sequence: the hardware
// seed (rcp.approx.ftz.f64, >= 20 bits) and one third-order step r(1 + e + e^2), e = 1 - x r, whose
// error is e^3 < 2^-60.  x == 0 or non-finite yields inf / NaN, which the callers replace by the
// reference's guarded values.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

__device__ __forceinline__ double jk_spread(const double (&v2)[NA])
{
    // cpwt_sub.f90:144-152 / 344-357: mean of the NA delete-one values, then
    // sqrt((na-1)/na * sum (v - mean)^2), NaN -> 0.
    double m = 0.0;
#pragma unroll
    for (int a = 0; a < NA; a++) m += v2[a];
    m = m * (1.0 / (double)NA);
    double d = 0.0;
#pragma unroll
    for (int a = 0; a < NA; a++) d = d + (v2[a] - m) * (v2[a] - m);
    d = sqrt(d * ((double)(NA - 1) / (double)NA));
    return (d != d) ? 0.0 : d;
}

// cpwt_sub.f90:107-160.  For the left-out azimuth a the reference walks the
// other ten in order and averages the ten running sums.  The partial sums over
// azimuths < a are the same for every a, so the walk for a resumes from the
// shared prefix; additions happen in the reference's order.
__device__ __forceinline__ double jk_sgram(const double (&s)[NA])
{
    double v2[NA];
    double p = 0.0, cum = 0.0;
#pragma unroll
    for (int a = 0; a < NA; a++) {
        double q = p, acc = cum;
#pragma unroll
        for (int a2 = a + 1; a2 < NA; a2++) {
            q += s[a2];
            acc += q;
        }
        v2[a] = acc * (1.0 / (double)(NA - 1));
        if (a < NA - 1) {
            p += s[a];
            cum += p;
        }
    }
    return jk_spread(v2);
}

// cpwt_sub.f90:329-335: co = Re(xp)^2/(p1*p2), ad = Re(xp)/p1, both 0 when a sum is 0 -- added straight into the
// running sums of the walk (sa += ad, sc += co).
// FAST (every per-azimuth power of the cell is a normal number in [1e-150, 1e150], checked once per cell by the
// caller on the high words, so no prefix product can underflow or overflow): one branch-free reciprocal serves both quotients
// (ad = xp*p2/(p1*p2)) with one fused multiply-add each; the sums are non-negative, so "a sum is 0" is "the
// product is 0", read from its exponent and mantissa bits with integer instructions; results differ from the
// reference's IEEE divisions by a few ulp, far inside the 1e-9 parity budget, at well under half of the FP64 work.
// Otherwise (grids of extreme amplitude, exact zeros): jk_admit_coh_exact below.
__device__ __forceinline__ void ratio_acc(double q1, double q2, double qx, double &sa, double &sc)
{
    const double prod = q1 * q2;
    const double xr = qx * fast_rcp(prod);
    const bool ok = ((__double2hiint(prod) & 0x7fffffff) | __double2loint(prod)) != 0;
    const double f = ok ? xr : 0.0;
    sc = fma(qx, f, sc);
    sa = fma(q2, f, sa);
}

// cpwt_sub.f90:298-366, same shared-prefix walk: 10 + 55 ratio evaluations
// instead of the reference's 110.
#ifndef PFX_RED_PAIR
#define PFX_RED_PAIR 1
#endif
__device__ __forceinline__ void jk_admit_coh(const double (&s1)[NA], const double (&s2)[NA], const double (&xr)[NA],
                                             double &ead, double &eco)
{
    double ad2[NA], co2[NA];
    double p1 = 0.0, p2 = 0.0, px = 0.0, cad = 0.0, cco = 0.0;
#if PFX_RED_PAIR
    // Walks a and a+1 side by side: from azimuth a+2 on both add the same powers, so their two dependent chains
    // (sum, product, reciprocal, two fused multiply-adds) sit next to each other in program order and overlap in
    // the FP64 pipe; every walk still adds in the reference's order, so the sums are those of the one-walk loop.
#pragma unroll
    for (int a = 0; a < NA; a += 2) {
        double qa1 = p1, qa2 = p2, qax = px, saa = cad, sca = cco;  // walk a resumes from the prefix before a
        if (a < NA - 1) {                                            // the shared prefix takes azimuth a
            p1 += s1[a];
            p2 += s2[a];
            px += xr[a];
            ratio_acc(p1, p2, px, cad, cco);
        }
        double qb1 = p1, qb2 = p2, qbx = px, sab = cad, scb = cco;  // walk a+1 resumes from the prefix before a+1
        if (a + 1 < NA - 1) {                                        // ... and azimuth a+1
            p1 += s1[a + 1];
            p2 += s2[a + 1];
            px += xr[a + 1];
            ratio_acc(p1, p2, px, cad, cco);
        }
        if (a + 1 < NA) {  // walk a alone over azimuth a+1
            qa1 += s1[a + 1];
            qa2 += s2[a + 1];
            qax += xr[a + 1];
            ratio_acc(qa1, qa2, qax, saa, sca);
        }
#pragma unroll
        for (int a2 = a + 2; a2 < NA; a2++) {
            qa1 += s1[a2];
            qb1 += s1[a2];
            qa2 += s2[a2];
            qb2 += s2[a2];
            qax += xr[a2];
            qbx += xr[a2];
            ratio_acc(qa1, qa2, qax, saa, sca);
            ratio_acc(qb1, qb2, qbx, sab, scb);
        }
        ad2[a] = saa * (1.0 / (double)(NA - 1));
        co2[a] = sca * (1.0 / (double)(NA - 1));
        if (a + 1 < NA) {
            ad2[a + 1] = sab * (1.0 / (double)(NA - 1));
            co2[a + 1] = scb * (1.0 / (double)(NA - 1));
        }
    }
#else
#pragma unroll
    for (int a = 0; a < NA; a++) {
        double q1 = p1, q2 = p2, qx = px, sa = cad, sc = cco;
#pragma unroll
        for (int a2 = a + 1; a2 < NA; a2++) {
            q1 += s1[a2];
            q2 += s2[a2];
            qx += xr[a2];
            ratio_acc(q1, q2, qx, sa, sc);
        }
        ad2[a] = sa * (1.0 / (double)(NA - 1));
        co2[a] = sc * (1.0 / (double)(NA - 1));
        if (a < NA - 1) {
            p1 += s1[a];
            p2 += s2[a];
            px += xr[a];
            ratio_acc(p1, p2, px, cad, cco);
        }
    }
#endif
    ead = jk_spread(ad2);
    eco = jk_spread(co2);
}

// High words of the doubles 1e-150 and 1e150: a power whose high word lies in [LO, HI] is a normal number in
// [1e-150, 1e150] (zeros, subnormals, huge values, infinities and NaNs all fall outside)
constexpr int PW_HI_LO = 0x20ca2fe8, PW_HI_HI = 0x5f138d34;

// the (real) admissibility terms of both grids at cell (i, j)
__device__ __forceinline__ double2 load_z(const PlaneView &v, int i, int j)
{
    return v.zc ? __ldg(v.zc + (size_t)j * v.ld + i) : make_double2(0.0, 0.0);
}

__device__ __forceinline__ double2 load_w(const PlaneView &v, int p, int i, int j, double2 zz)
{
    const double2 w = __ldg(v.base + (size_t)p * v.plane_stride + (size_t)(j + v.off_j) * v.ld + (i + v.off_i));
    if (v.zc) return make_double2(w.x - v.ce[p] * (v.zsel ? zz.y : zz.x), w.y);
    return make_double2(w.x * v.scale, w.y * v.scale);
}

// The rare path of the admittance / coherence jackknife (cells with a power outside [1e-150, 1e150], see
// reduce_kernel and fixup_kernel): cpwt_sub.f90:310-357 literally -- eleven walks of ten running sums, the
// reference's guard, IEEE divisions, the coherence quotient formed as (xp/p1)*(xp/p2) so that p1*p2 is never taken
// -- as rolled loops that read the coefficients again.
__device__ __forceinline__ void jk_admit_coh_exact(const PlaneView &w1, const PlaneView &w2, int i, int j, double2 zz, double &ead,
                                   double &eco)
{
    double ad2[NA], co2[NA];
#pragma unroll 1
    for (int a = 0; a < NA; a++) {
        double p1 = 0.0, p2 = 0.0, px = 0.0, sa = 0.0, sc = 0.0;
#pragma unroll 1
        for (int a2 = 0; a2 < NA; a2++) {
            if (a2 == a) continue;
            const double2 u = load_w(w1, a2, i, j, zz), v = load_w(w2, a2, i, j, zz);
            p1 += fabs(u.x * u.x + u.y * u.y);
            p2 += fabs(v.x * v.x + v.y * v.y);
            px += u.x * v.x + u.y * v.y;
            if (p1 != 0.0 && p2 != 0.0) {
                const double ad = px / p1;
                sa += ad;
                sc += ad * (px / p2);
            }
        }
        ad2[a] = sa * (1.0 / (double)(NA - 1));
        co2[a] = sc * (1.0 / (double)(NA - 1));
    }
    ead = jk_spread(ad2);
    eco = jk_spread(co2);
}

// F32: the outputs are float arrays (same element offsets): each float64 result is rounded once on store.
template <bool F32>
__device__ __forceinline__ void put(double *base, size_t o, double v)
{
    if (F32)
        reinterpret_cast<float *>(base)[o] = (float)v;
    else
        base[o] = v;
}

// the four maps of the shard that owns row j: [map][scale][row in shard][x]
__device__ __forceinline__ void shard_dest(const ShardOut &sh, int nx, int ny, int i, int j, double *&out0, double *&out1,
                                           double *&out2, double *&out3, size_t &o)
{
    const int q = j / sh.rows_per_shard, jl = j - q * sh.rows_per_shard;
    const int rq = min(sh.rows_per_shard, ny - q * sh.rows_per_shard);
    const size_t map_stride = (size_t)sh.ns * rq * nx;
    out0 = sh.base[q];
    out1 = out0 + map_stride;
    out2 = out1 + map_stride;
    out3 = out2 + map_stride;
    o = ((size_t)sh.scale * rq + jl) * nx + i;
}

// smallest high word of the 22 powers and the check itself (see reduce_kernel)
__device__ __forceinline__ bool powers_out_of_range(int hmin, double t1sum, double t2sum)
{
    return hmin < PW_HI_LO || max(__double2hiint(t1sum), __double2hiint(t2sum)) > PW_HI_HI;
}

template <int MODE, bool SHARDED, bool F32, int OCC = 4>
__global__ void __launch_bounds__(128, OCC)
    reduce_kernel(int nx, int ny, PlaneView w1, PlaneView w2, double *__restrict__ out0, double *__restrict__ out1,
                  double *__restrict__ out2, double *__restrict__ out3, ShardOut sh, int *odd_flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx) return;
    size_t o = (size_t)j * nx + i;
    if (SHARDED) {
        double *p0, *p1, *p2, *p3;
        shard_dest(sh, nx, ny, i, j, p0, p1, p2, p3, o);
        out0 = p0;
        out1 = p1;
        out2 = p2;
        out3 = p3;
    }
    const double2 z1 = load_z(w1, i, j);  // both grids' terms live in one array
    const double2 z2 = z1;

    if (MODE == RED_SCALOGRAM) {
        double s[NA];
        double tot = 0.0;
#pragma unroll
        for (int a = 0; a < NA; a++) {
            const double2 w = load_w(w1, a, i, j, z1);
            s[a] = fabs(w.x * w.x + w.y * w.y);  // CABS(w*CONJG(w)), cpwt.f90:224
            tot += s[a];
        }
        put<F32>(out0, o, tot * (1.0 / (double)NA));  // cpwt.f90:234
        put<F32>(out1, o, jk_sgram(s));
    } else if (MODE == RED_XSCALOGRAM) {
        double xr[NA], xi[NA];
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int a = 0; a < NA; a++) {
            const double2 u = load_w(w1, a, i, j, z1), v = load_w(w2, a, i, j, z2);
            xr[a] = u.x * v.x + u.y * v.y;  // w1*CONJG(w2), cpwt.f90:274
            xi[a] = u.y * v.x - u.x * v.y;
            tr += xr[a];
            ti += xi[a];
        }
        reinterpret_cast<double2 *>(out0)[o] = make_double2(tr * (1.0 / (double)NA), ti * (1.0 / (double)NA));
        const double er = jk_sgram(xr), ei = jk_sgram(xi);
        put<F32>(out1, o, sqrt(er * er + ei * ei));  // cpwt.f90:290
    } else {
        double s1[NA], s2[NA], xr[NA];
        double t1 = 0.0, t2 = 0.0, tx = 0.0;
        int hmin = 0x7fffffff;  // smallest high word of the 22 powers (non-negative doubles)
#pragma unroll
        for (int a = 0; a < NA; a++) {
            const double2 u = load_w(w1, a, i, j, z1), v = load_w(w2, a, i, j, z2);
            s1[a] = fabs(u.x * u.x + u.y * u.y);  // cpwt.f90:333
            s2[a] = fabs(v.x * v.x + v.y * v.y);  // cpwt.f90:334
            xr[a] = u.x * v.x + u.y * v.y;        // Re(w1*CONJG(w2)), cpwt.f90:335
#ifndef PFX_RED_NOCHECK
            hmin = min(hmin, min(__double2hiint(s1[a]), __double2hiint(s2[a])));
#endif
            t1 += s1[a];
            t2 += s2[a];
            tx += xr[a];
        }
        // a power that is zero, subnormal or below 1e-150, or a sum over azimuths beyond 1e150 / not finite (every
        // prefix sum lies between the smallest power and that sum): the cell takes the exact path
        // -- flagged for fixup_kernel, which recomputes its two error values after this kernel
#ifndef PFX_RED_NOCHECK
        if (powers_out_of_range(hmin, t1, t2)) *odd_flag = 1;
#endif
        t1 = t1 * (1.0 / (double)NA);  // cpwt.f90:348-350
        t2 = t2 * (1.0 / (double)NA);
        tx = tx * (1.0 / (double)NA);
        put<F32>(out0, o, tx / t1);                                            // cpwt.f90:354
        put<F32>(out2, o, (tx * tx) / (t1 * t2));                              // cpwt.f90:355
        double ead, eco;
        jk_admit_coh(s1, s2, xr, ead, eco);
        put<F32>(out1, o, ead);
        put<F32>(out3, o, eco);
    }
}

// Second half of the admittance / coherence epilogue: when reduce_kernel met a cell whose powers leave
// [1e-150, 1e150] (odd_flag), every such cell gets its two jackknife errors from the exact path.  A small persistent
// grid that returns at once in the common case.
template <bool SHARDED, bool F32>
__global__ void __launch_bounds__(128)
    fixup_kernel(int nx, int ny, PlaneView w1, PlaneView w2, double *__restrict__ out1g, double *__restrict__ out3g,
                 ShardOut sh, const int *odd_flag)
{
    if (*odd_flag == 0) return;
    const size_t ncell = (size_t)nx * ny;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(c % nx), j = (int)(c / nx);
        const double2 zz = load_z(w1, i, j);
        int hmin = 0x7fffffff;
        double t1 = 0.0, t2 = 0.0;
#pragma unroll 1
        for (int a = 0; a < NA; a++) {
            const double2 u = load_w(w1, a, i, j, zz), v = load_w(w2, a, i, j, zz);
            const double p1 = fabs(u.x * u.x + u.y * u.y), p2 = fabs(v.x * v.x + v.y * v.y);
            hmin = min(hmin, min(__double2hiint(p1), __double2hiint(p2)));
            t1 += p1;
            t2 += p2;
        }
        if (!powers_out_of_range(hmin, t1, t2)) continue;
        double *out0 = nullptr, *out1 = out1g, *out2 = nullptr, *out3 = out3g;
        size_t o = c;
        if (SHARDED) shard_dest(sh, nx, ny, i, j, out0, out1, out2, out3, o);
        double ead, eco;
        jk_admit_coh_exact(w1, w2, i, j, zz, ead, eco);
        put<F32>(out1, o, ead);
        put<F32>(out3, o, eco);
    }
}

template <bool C64>
__global__ void crop_kernel(int nx, int ny, PlaneView w, void *__restrict__ slab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int p = blockIdx.z;
    if (i >= nx) return;
    const double2 v = load_w(w, p, i, j, load_z(w, i, j));
    const size_t o = ((size_t)p * ny + j) * nx + i;
    if (C64)
        static_cast<float2 *>(slab)[o] = make_float2((float)v.x, (float)v.y);
    else
        static_cast<double2 *>(slab)[o] = v;
}

}  // namespace

int launch_reduce(int mode, int nx, int ny, PlaneView w1, PlaneView w2, double *out0, double *out1, double *out2,
                  double *out3, cudaStream_t stream, const ShardOut *shards, bool f32, int *flag_word)
{
    dim3 block(128), grid((nx + 127) / 128, ny);
    const ShardOut none;
    static const int occ = [] {  // CTAs of 128 threads per SM the admittance / coherence epilogue is compiled for
        const char *e = getenv("PLATEFLEX_B200_REDUCE_OCC");
        return (e && e[0] == '3') ? 3 : 4;
    }();
    const bool sharded = shards && shards->nshards > 0;
    if (sharded) PFX_ARG(mode == RED_ADMIT_COH && !f32, "row-sharded output exists for the float64 admittance / coherence epilogue only");
    PFX_ARG(!(f32 && mode == RED_XSCALOGRAM), "float32 outputs exist for the scalogram and admittance / coherence epilogues");
    switch (mode) {
        case RED_SCALOGRAM:
            if (f32)
                reduce_kernel<RED_SCALOGRAM, false, true><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, nullptr);
            else
                reduce_kernel<RED_SCALOGRAM, false, false><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, nullptr);
            break;
        case RED_XSCALOGRAM:
            reduce_kernel<RED_XSCALOGRAM, false, false><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, nullptr);
            break;
        case RED_ADMIT_COH: {
            // per-launch flag word, allocated and freed in stream order: "some cell needs the exact jackknife"
            int *flag = flag_word;
            if (!flag_word) PFX_CUDA(cudaMallocAsync((void **)&flag, sizeof(int), stream));
            PFX_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
            const ShardOut &sh = sharded ? *shards : none;
            const dim3 fgrid(148 * 4);  // a small persistent grid (any size is correct)
            if (sharded) {
                if (occ == 3)
                    reduce_kernel<RED_ADMIT_COH, true, false, 3><<<grid, block, 0, stream>>>(nx, ny, w1, w2, nullptr, nullptr, nullptr, nullptr, sh, flag);
                else
                    reduce_kernel<RED_ADMIT_COH, true, false><<<grid, block, 0, stream>>>(nx, ny, w1, w2, nullptr, nullptr, nullptr, nullptr, sh, flag);
                fixup_kernel<true, false><<<fgrid, block, 0, stream>>>(nx, ny, w1, w2, nullptr, nullptr, sh, flag);
            } else if (f32) {
                reduce_kernel<RED_ADMIT_COH, false, true><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, flag);
                fixup_kernel<false, true><<<fgrid, block, 0, stream>>>(nx, ny, w1, w2, out1, out3, none, flag);
            } else {
                if (occ == 3)
                    reduce_kernel<RED_ADMIT_COH, false, false, 3><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, flag);
                else
                    reduce_kernel<RED_ADMIT_COH, false, false><<<grid, block, 0, stream>>>(nx, ny, w1, w2, out0, out1, out2, out3, none, flag);
                fixup_kernel<false, false><<<fgrid, block, 0, stream>>>(nx, ny, w1, w2, out1, out3, none, flag);
            }
            PFX_CUDA(cudaGetLastError());
            if (!flag_word) PFX_CUDA(cudaFreeAsync(flag, stream));
            break;
        }
        default:
            set_error("launch_reduce: bad mode %d", mode);
            return PFX_ERR_ARG;
    }
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

int launch_crop(int nx, int ny, PlaneView w, void *slab, bool c64, cudaStream_t stream)
{
    dim3 block(128), grid((nx + 127) / 128, ny, PFX_NA);
    if (c64)
        crop_kernel<true><<<grid, block, 0, stream>>>(nx, ny, w, slab);
    else
        crop_kernel<false><<<grid, block, 0, stream>>>(nx, ny, w, slab);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
