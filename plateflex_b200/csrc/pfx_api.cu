// pfx_api.cu -- C ABI of libplateflex_b200.so: plan management and the CWT entry
// points declared in include/plateflex_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

#include "pfx_plan.hpp"

namespace pfx {

static thread_local std::string g_err;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace pfx

using namespace pfx;

namespace {
// independent FMA chains: every thread keeps 8 accumulators busy, 32 fused multiply-adds per loop trip
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double x, double y)
{
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = (double)(threadIdx.x + j) * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma(a[j], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

int pfx_plan::alloc(void **p, size_t bytes)
{
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return PFX_ERR_OOM;
    }
    owned.push_back(*p);
    device_bytes += bytes;
    return PFX_OK;
}

int pfx_plan::scratch(int slot, size_t bytes, void **p)
{
    if (bytes > scratch_bytes[slot]) {
        auto release = [&](int s) {
            if (d_scratch[s]) {
                cudaFree(d_scratch[s]);
                device_bytes -= scratch_bytes[s];
                d_scratch[s] = nullptr;
                scratch_bytes[s] = 0;
            }
        };
        PFX_CUDA(cudaStreamSynchronize(stream));
        release(slot);
        cudaError_t e = cudaMalloc(&d_scratch[slot], bytes);
        if (e != cudaSuccess) {
            // Staging buffers of earlier calls (other output dtypes, other entry points) are only a cache:
            // give back those not handed out in the current call and try once more.
            cudaGetLastError();
            PFX_CUDA(cudaDeviceSynchronize());
            for (int s = 0; s < (int)(sizeof d_scratch / sizeof d_scratch[0]); s++)
                if (scratch_call[s] != call_id) release(s);
            e = cudaMalloc(&d_scratch[slot], bytes);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            d_scratch[slot] = nullptr;
            set_error("device scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return PFX_ERR_OOM;
        }
        scratch_bytes[slot] = bytes;
        device_bytes += bytes;
    }
    scratch_call[slot] = call_id;
    *p = d_scratch[slot];
    return PFX_OK;
}

cudaEvent_t pfx_plan::prof_event()
{
    if (!spare_events.empty()) {
        cudaEvent_t e = spare_events.back();
        spare_events.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return e;
}

namespace {
// dst[j * rows + i] = src[i * cols + j]: the Fortran-order image of a C-order (rows, cols) array
template <class T>
__global__ void __launch_bounds__(256) transpose_kernel(int rows, int cols, const T *__restrict__ src, T *__restrict__ dst)
{
    __shared__ T tile[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        if (i < rows && j < cols) tile[r][threadIdx.x] = src[(size_t)i * cols + j];
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int i = i0 + threadIdx.x, j = j0 + c;
        if (i < rows && j < cols) dst[(size_t)j * rows + i] = tile[threadIdx.x][c];
    }
}
}  // namespace


extern "C" {

int pfx_version(void) { return 101; }

const char *pfx_last_error(void) { return g_err.c_str(); }

int pfx_device_count(int *count)
{
    PFX_ARG(count, "count is null");
    PFX_CUDA(cudaGetDeviceCount(count));
    return PFX_OK;
}

int pfx_npow2(int n)
{
    int n2 = 1;
    do n2 *= 2;
    while (n2 < n);
    return n2;
}

int pfx_fp64_peak(int device, double *tflops)
{
    PFX_ARG(tflops, "tflops is null");
    DeviceGuard guard(device);
    PFX_ARG(guard.ok, "device ordinal out of range");
    int nsm = 0;
    PFX_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int blocks = nsm * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    PFX_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {  // first trip warms up; best of the rest
        cudaEventRecord(e0, 0);
        dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) {
        set_error("fp64 peak kernel failed: %s", cudaGetErrorString(e));
        return PFX_ERR_CUDA;
    }
    *tflops = 2.0 * 32.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return PFX_OK;
}

int pfx_dev_alloc(int device, size_t bytes, void **dev_ptr)
{
    PFX_ARG(dev_ptr, "dev_ptr is null");
    *dev_ptr = nullptr;
    DeviceGuard guard(device);
    PFX_ARG(guard.ok, "device ordinal out of range");
    cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;
    }
    return PFX_OK;
}

int pfx_dev_free(int device, void *dev_ptr)
{
    DeviceGuard guard(device);
    if (dev_ptr) PFX_CUDA(cudaFree(dev_ptr));
    return PFX_OK;
}

int pfx_dev_h2d(int device, void *dst_dev, const void *src_host, size_t bytes)
{
    PFX_ARG(dst_dev && src_host, "null pointer");
    DeviceGuard guard(device);
    PFX_CHECK(staged_h2d(dst_dev, src_host, bytes, 0));
    PFX_CUDA(cudaStreamSynchronize(0));
    return PFX_OK;
}

int pfx_dev_h2d_transposed(int device, void *dst_dev, const void *src_host, int rows, int cols, int elem_bytes)
{
    PFX_ARG(dst_dev && src_host, "null pointer");
    PFX_ARG(rows >= 1 && cols >= 1 && (elem_bytes == 8 || elem_bytes == 1), "rows, cols >= 1 and 8- or 1-byte elements");
    DeviceGuard guard(device);
    const size_t bytes = (size_t)rows * cols * elem_bytes;
    void *tmp = nullptr;
    PFX_CUDA(cudaMalloc(&tmp, bytes));
    int rc = staged_h2d(tmp, src_host, bytes, 0);
    if (rc == PFX_OK) {
        const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
        if (elem_bytes == 8)
            transpose_kernel<double><<<grid, block>>>(rows, cols, static_cast<const double *>(tmp), static_cast<double *>(dst_dev));
        else
            transpose_kernel<unsigned char><<<grid, block>>>(rows, cols, static_cast<const unsigned char *>(tmp),
                                                             static_cast<unsigned char *>(dst_dev));
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(0) != cudaSuccess) {
            set_error("transposing upload failed");
            rc = PFX_ERR_CUDA;
        }
    }
    cudaFree(tmp);
    return rc;
}

int pfx_dev_d2h(int device, void *dst_host, const void *src_dev, size_t bytes)
{
    PFX_ARG(dst_host && src_dev, "null pointer");
    DeviceGuard guard(device);
    PFX_CHECK(staged_d2h(dst_host, src_dev, bytes, 0));  // synchronous
    return PFX_OK;
}

int pfx_plan_create(pfx_plan **out, int nx, int ny, double dx_km, double dy_km, const double *k_rad_m, int ns, double k0,
                    int engine, int device)
{
    PFX_ARG(out, "plan is null");
    *out = nullptr;
    PFX_ARG(nx >= 2 && ny >= 2 && nx <= (1 << 14) && ny <= (1 << 14), "nx, ny must be in [2, 16384]");
    PFX_ARG(dx_km > 0 && dy_km > 0, "dx, dy must be positive");
    PFX_ARG(k_rad_m && ns >= 1, "k is null or ns < 1");
    PFX_ARG(k0 > 0, "k0 must be positive");
    PFX_ARG(engine == PFX_ENGINE_PRUNED || engine == PFX_ENGINE_DENSE, "unknown engine");
    for (int i = 0; i < ns; i++) PFX_ARG(k_rad_m[i] > 0 && std::isfinite(k_rad_m[i]), "k must be positive and finite");
    int ndev = 0;
    PFX_CUDA(cudaGetDeviceCount(&ndev));
    PFX_ARG(device >= 0 && device < ndev, "device ordinal out of range");
    DeviceGuard guard(device);

    pfx_plan *pl = new pfx_plan();
    pl->nx = nx;
    pl->ny = ny;
    pl->ns = ns;
    pl->NX = pfx_npow2(2 * nx);  // cpwt.f90:94-95
    pl->NY = pfx_npow2(2 * ny);
    pl->dx = dx_km;
    pl->dy = dy_km;
    pl->k0 = k0;
    pl->engine = engine;
    pl->spec_t = engine == PFX_ENGINE_PRUNED;
    pl->device = device;
    pl->k.assign(k_rad_m, k_rad_m + ns);

    // Scale / angle constants (cpwt.f90:131,148-161; cpwt_sub.f90:85-87).
    const double pi = 3.141592653589793;
    const double da = 2. * std::sqrt(-2. * std::log(0.75)) / k0;
    pl->consts.resize(ns);
    for (int is = 0; is < ns; is++) {
        PfxScaleConsts &c = pl->consts[is];
        const double kk = pl->k[is] * 1.e3;
        c.s = k0 / kk;
        c.norm = c.s / std::sqrt(pi);
        for (int ia = 0; ia < PFX_NA; ia++) {
            const double angle = (double)ia * da - pi / 2.e0;
            c.kx0[ia] = k0 * std::cos(angle);
            c.ky0[ia] = k0 * std::sin(angle);
            c.e0[ia] = std::exp(-0.5 * (c.kx0[ia] * c.kx0[ia] + c.ky0[ia] * c.ky0[ia]));
        }
    }

    auto fail = [&](int rc) {
        pfx_plan_destroy(pl);
        return rc;
    };
    const size_t P = (size_t)pl->NX * pl->NY;
    int rc;
    if ((rc = pl->alloc((void **)&pl->d_kx, pl->NX * sizeof(double)))) return fail(rc);
    if ((rc = pl->alloc((void **)&pl->d_ky, pl->NY * sizeof(double)))) return fail(rc);
    if ((rc = pl->alloc((void **)&pl->d_u2, pl->NX * sizeof(double)))) return fail(rc);
    if ((rc = pl->alloc((void **)&pl->d_v2, pl->NY * sizeof(double)))) return fail(rc);
    if ((rc = pl->alloc((void **)&pl->d_partial, 512 * sizeof(double)))) return fail(rc);
    if ((rc = pl->alloc((void **)&pl->d_flag_words, 2 * sizeof(int)))) return fail(rc);
    // the pruned engine allocates the padded spectra itself unless it can use the cosine-transform form (pfx_dct.cu)
    if (engine != PFX_ENGINE_PRUNED)
        for (int g = 0; g < 2; g++)
            if ((rc = pl->alloc((void **)&pl->d_spec[g], P * sizeof(double2)))) return fail(rc);
    if ((rc = plan_init_axes(pl))) return fail(rc);
    {
        cudaError_t e = cudaStreamCreateWithFlags(&pl->s_aux, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; b++) {
            e = cudaEventCreateWithFlags(&pl->ev_planes[b], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_reduced[b], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) {
            int lo = 0, hi = 0;  // pass 1 of the next scale takes SM slots as soon as pass-2 CTAs retire
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            e = cudaStreamCreateWithPriority(&pl->s_p1, cudaStreamNonBlocking, hi);
        }
        for (int b = 0; b < 2 && e == cudaSuccess; b++) {
            e = cudaEventCreateWithFlags(&pl->ev_t_ready[b], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_t_free[b], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_spec, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            set_error("stream / event creation failed: %s", cudaGetErrorString(e));
            return fail(PFX_ERR_CUDA);
        }
    }

    // forward FFT plan (cpwt.f90:122): one 2-D Z2Z per grid, work area owned by the plan
    {
        size_t ws = 0;
        cufftResult r = cufftCreate(&pl->fft_fwd);
        if (r == CUFFT_SUCCESS) {
            pl->fft_fwd_ok = true;
            r = cufftSetAutoAllocation(pl->fft_fwd, 0);
        }
        int n[2] = {pl->NY, pl->NX};  // slowest first
        if (pl->spec_t) std::swap(n[0], n[1]);
        if (r == CUFFT_SUCCESS) r = cufftMakePlanMany(pl->fft_fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2Z, 1, &ws);
        if (r != CUFFT_SUCCESS) {
            set_error("cufft forward plan %dx%d failed: %d", pl->NY, pl->NX, (int)r);
            return fail(PFX_ERR_CUFFT);
        }
        if ((rc = pl->alloc(&pl->d_fft_work, ws))) return fail(rc);
        pl->fft_work_bytes = ws;
        if (cufftSetWorkArea(pl->fft_fwd, pl->d_fft_work) != CUFFT_SUCCESS) return fail(PFX_ERR_CUFFT);
    }
    rc = (engine == PFX_ENGINE_DENSE) ? plan_init_dense(pl) : plan_init_pruned(pl);
    if (rc) return fail(rc);
    cudaError_t e = cudaStreamSynchronize(pl->stream);
    if (e != cudaSuccess) {
        set_error("plan initialisation failed: %s", cudaGetErrorString(e));
        return fail(PFX_ERR_CUDA);
    }
    *out = pl;
    return PFX_OK;
}

int pfx_plan_destroy(pfx_plan *pl)
{
    if (!pl) return PFX_OK;
    DeviceGuard guard(pl->device);
    cudaDeviceSynchronize();
    plan_free_pruned(pl);
    if (pl->s_aux) cudaStreamDestroy(pl->s_aux);
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_p1) cudaStreamDestroy(pl->s_p1);
    for (int b = 0; b < 2; b++) {
        if (pl->ev_t_ready[b]) cudaEventDestroy(pl->ev_t_ready[b]);
        if (pl->ev_t_free[b]) cudaEventDestroy(pl->ev_t_free[b]);
    }
    if (pl->ev_spec) cudaEventDestroy(pl->ev_spec);
    for (int b = 0; b < 2; b++) {
        if (pl->ev_planes[b]) cudaEventDestroy(pl->ev_planes[b]);
        if (pl->ev_reduced[b]) cudaEventDestroy(pl->ev_reduced[b]);
    }
    if (pl->ev_join) cudaEventDestroy(pl->ev_join);
    for (const PfxProfSpan &sp : pl->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (cudaEvent_t e : pl->spare_events) cudaEventDestroy(e);
    if (pl->fft_fwd_ok) cufftDestroy(pl->fft_fwd);
    for (int g = 0; g < 2; g++)
        if (pl->fft_inv_ok[g]) cufftDestroy(pl->fft_inv[g]);
    for (void *p : pl->owned) cudaFree(p);
    for (void *p : pl->d_scratch)
        if (p) cudaFree(p);
    delete pl;
    return PFX_OK;
}

int pfx_plan_set_stream(pfx_plan *pl, void *stream)
{
    PFX_ARG(pl, "plan is null");
    pl->stream = (cudaStream_t)stream;
    return PFX_OK;
}

int pfx_plan_info(const pfx_plan *pl, int *nnx, int *nny, size_t *device_bytes)
{
    PFX_ARG(pl, "plan is null");
    if (nnx) *nnx = pl->NX;
    if (nny) *nny = pl->NY;
    if (device_bytes) *device_bytes = pl->device_bytes;
    return PFX_OK;
}

int pfx_plan_launch_count(const pfx_plan *pl, long long *kernels)
{
    PFX_ARG(pl && kernels, "null argument");
    *kernels = pl->launches;
    return PFX_OK;
}

int pfx_plan_scale_cost(const pfx_plan *pl, double *cost)
{
    PFX_ARG(pl && cost, "null argument");
    for (int is = 0; is < pl->ns; is++)
        cost[is] = (pl->engine == PFX_ENGINE_PRUNED) ? pruned_scale_cost(pl, is) : 1.0;  // dense: every scale costs the same
    return PFX_OK;
}

int pfx_plan_profile_enable(pfx_plan *pl, int on)
{
    PFX_ARG(pl, "plan is null");
    pl->profiling = on != 0;
    return PFX_OK;
}

int pfx_plan_profile_read(pfx_plan *pl, double *ms, long long *count, int n)
{
    PFX_ARG(pl && ms && count && n >= 1, "null argument");
    DeviceGuard guard(pl->device);
    for (int i = 0; i < n; i++) {
        ms[i] = 0.0;
        count[i] = 0;
    }
    for (const PfxProfSpan &sp : pl->spans) {
        PFX_CUDA(cudaEventSynchronize(sp.b));
        float t = 0.f;
        PFX_CUDA(cudaEventElapsedTime(&t, sp.a, sp.b));
        if (sp.cls < n) {
            ms[sp.cls] += (double)t;
            count[sp.cls]++;
        }
        pl->spare_events.push_back(sp.a);
        pl->spare_events.push_back(sp.b);
    }
    pl->spans.clear();
    return PFX_OK;
}

int pfx_plan_set_cutoff(pfx_plan *pl, double nsigma)
{
    PFX_ARG(pl, "plan is null");
    PFX_ARG(nsigma >= 4.0, "cutoff must be >= 4 sigma");
    pl->cutoff = nsigma;
    if (pl->engine == PFX_ENGINE_PRUNED) {
        DeviceGuard guard(pl->device);
        return plan_init_pruned(pl);
    }
    return PFX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// Shared driver: forward spectra once (cpwt.f90:122), then per scale the 11 azimuth
// planes of each grid followed by the requested epilogue.
// ------------------------------------------------------------------------------------
namespace {

enum { OUT_CUBE = 100 };

int check_range(const pfx_plan *pl, int is0, int is1)
{
    PFX_ARG(pl, "plan is null");
    PFX_ARG(is0 >= 0 && is1 <= pl->ns && is0 < is1, "scale range [is0, is1) out of bounds");
    return PFX_OK;
}

// h_outs (optional): host destinations of the outputs; each finished scale is copied back on the copy
// stream while later scales are still being computed.  elems[i] = doubles per cell of output i.
// h_outs (optional): host destinations of the outputs; each finished scale is copied back on the copy
// stream while later scales are still being computed.  With d32 / h32 set the epilogue stores float32
// (each float64 result rounded once) into the device staging arrays d32 and the host receives those.
int run_cwt(pfx_plan *pl, int mode, const double *g1, const double *g2, int is0, int is1, double *const *d_outs,
            double *const *h_outs, const int *elems, int nouts, const ShardOut *shards = nullptr,
            float *const *d32 = nullptr, float *const *h32 = nullptr, const int *scale_list = nullptr)
{
    // scale_list (optional, is1 - is0 entries): the scales to transform, in this order, instead of the
    // contiguous range [is0, is1) -- a multi-GPU caller may own any subset of the scale axis
    const int is_base = is0;
    const int ngrids = g2 ? 2 : 1;
    const size_t nxy = (size_t)pl->nx * pl->ny;
    {
        NvtxRange r("pfx:forward");
        const bool pr = pl->engine == PFX_ENGINE_PRUNED;
        PFX_CHECK(pr ? pruned_forward(pl, g1, 0) : forward_spectrum(pl, g1, 0));
        if (g2) PFX_CHECK(pr ? pruned_forward(pl, g2, 1) : forward_spectrum(pl, g2, 1));
    }
    const bool pruned = pl->engine == PFX_ENGINE_PRUNED;
    if (pruned) {
        NvtxRange r("pfx:prepare");
        PFX_CHECK(pruned_prepare(pl, ngrids));
    }
    // With two plane buffers (pruned engine) the reduction of scale s overlaps the passes of scale s+1.
    const int nbuf = pruned ? pruned_num_buffers(pl) : 1;
    // Pageable host destinations (plain numpy arrays): every scale is enqueued first, then the finished
    // scales are drained through the pinned staging chunks of pfx_hostio.cu while the later ones compute.
    const bool staged = (h32 && host_is_pageable(h32[0])) || (!h32 && h_outs && host_is_pageable(h_outs[0]));
    struct Events {
        std::vector<cudaEvent_t> v;
        ~Events()
        {
            for (cudaEvent_t e : v) cudaEventDestroy(e);
        }
    } scale_done;
    // (per-kernel timing serialises everything on the caller's stream so that event pairs time one kernel each)
    const bool overlap = nbuf == 2 && mode != OUT_CUBE && !pl->profiling;
    cudaStream_t red_stream = overlap ? pl->s_aux : pl->stream;
    // pass 1 of scale n+1 on s_p1 while pass 2 of scale n runs (PLATEFLEX_B200_P1_AHEAD=off: one stream)
    static const bool p1_ahead_on = [] {
        const char *e = getenv("PLATEFLEX_B200_P1_AHEAD");
        return !(e && !strcmp(e, "off"));
    }();
    const bool p1_ahead = overlap && pruned && p1_ahead_on;
    const int nsc_call = is1 - is_base;
    auto scale_at = [&](int n) { return scale_list ? scale_list[n] : is_base + n; };
    if (p1_ahead) {
        PFX_CUDA(cudaEventRecord(pl->ev_spec, pl->stream));  // spectra and Z are ready
        PFX_CUDA(cudaStreamWaitEvent(pl->s_p1, pl->ev_spec, 0));
        PFX_CHECK(pruned_pass1(pl, scale_at(0), ngrids, 0, pl->s_p1));
        PFX_CUDA(cudaEventRecord(pl->ev_t_ready[0], pl->s_p1));
    }
    for (int n = 0; n < is1 - is_base; n++) {
        const int is = scale_list ? scale_list[n] : is_base + n;  // n: position in this call's sequence
        const int b = n & 1;
        const size_t so = (size_t)n * nxy;
        PlaneView v1, v2;
        NvtxRange scale_range("pfx:scale", is, "");
        if (!pruned) {
            PFX_CHECK(dense_scale_planes(pl, is, ngrids));
            v1 = dense_view(pl, 0);
            v2 = dense_view(pl, ngrids - 1);
        } else {
            if (overlap) {
                pruned_select_buffer(pl, b);
                if (n >= 2) PFX_CUDA(cudaStreamWaitEvent(pl->stream, pl->ev_reduced[b], 0));
            }
            if (p1_ahead) {
                PFX_CUDA(cudaStreamWaitEvent(pl->stream, pl->ev_t_ready[b], 0));
                PFX_CHECK(pruned_pass2(pl, is, ngrids, b, pl->stream));
                PFX_CUDA(cudaEventRecord(pl->ev_t_free[b], pl->stream));
                if (n + 1 < nsc_call) {
                    // T[b^1] was last read by pass 2 of scale n-1
                    if (n >= 1) PFX_CUDA(cudaStreamWaitEvent(pl->s_p1, pl->ev_t_free[b ^ 1], 0));
                    PFX_CHECK(pruned_pass1(pl, scale_at(n + 1), ngrids, b ^ 1, pl->s_p1));
                    PFX_CUDA(cudaEventRecord(pl->ev_t_ready[b ^ 1], pl->s_p1));
                }
            } else {
                PFX_CHECK(pruned_scale_planes(pl, is, ngrids));
            }
            v1 = pruned_view(pl, 0);
            v2 = pruned_view(pl, ngrids - 1);
        }
        if (overlap) {
            PFX_CUDA(cudaEventRecord(pl->ev_planes[b], pl->stream));
            PFX_CUDA(cudaStreamWaitEvent(red_stream, pl->ev_planes[b], 0));
        }
        if (mode == OUT_CUBE) {
            ProfScope prof(pl, PFX_KC_CROP, red_stream);
            PFX_CHECK(launch_crop(pl->nx, pl->ny, v1, reinterpret_cast<double2 *>(d_outs[0]) + so * PFX_NA, false, red_stream));
        } else {
            double *o[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int i = 0; i < nouts && d_outs && !d32; i++) o[i] = d_outs[i] + so * elems[i];
            for (int i = 0; i < nouts && d32; i++) o[i] = reinterpret_cast<double *>(d32[i] + so * elems[i]);
            NvtxRange r("pfx:reduce");
            ProfScope prof(pl, PFX_KC_REDUCE, red_stream);
            ShardOut sh;
            if (shards) {
                sh = *shards;
                sh.scale = is;
            }
            PFX_CHECK(launch_reduce(mode, pl->nx, pl->ny, v1, v2, o[0], o[1], o[2], o[3], red_stream, shards ? &sh : nullptr,
                                    d32 != nullptr, pl->d_flag_words + b));
        }
        pl->launches++;
        if (overlap || h_outs || h32) PFX_CUDA(cudaEventRecord(pl->ev_reduced[b], red_stream));
        if (staged) {
            cudaEvent_t e = nullptr;
            PFX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            scale_done.v.push_back(e);
            PFX_CUDA(cudaEventRecord(e, red_stream));
        } else if (h32) {
            PFX_CUDA(cudaStreamWaitEvent(pl->s_copy, pl->ev_reduced[b], 0));
            for (int i = 0; i < nouts; i++)
                PFX_CUDA(cudaMemcpyAsync(h32[i] + so * elems[i], d32[i] + so * elems[i],
                                         nxy * elems[i] * sizeof(float), cudaMemcpyDeviceToHost, pl->s_copy));
        } else if (h_outs) {
            PFX_CUDA(cudaStreamWaitEvent(pl->s_copy, pl->ev_reduced[b], 0));
            for (int i = 0; i < nouts; i++)
                PFX_CUDA(cudaMemcpyAsync(h_outs[i] + so * elems[i], d_outs[i] + so * elems[i],
                                         nxy * elems[i] * sizeof(double), cudaMemcpyDeviceToHost, pl->s_copy));
        }
    }
    // join the internal streams back into the caller's stream
    if (overlap) {
        PFX_CUDA(cudaEventRecord(pl->ev_join, pl->s_aux));
        PFX_CUDA(cudaStreamWaitEvent(pl->stream, pl->ev_join, 0));
    }
    if (staged) {
        for (int n = 0; n < is1 - is_base; n++) {
            const size_t so = (size_t)n * nxy;
            PFX_CUDA(cudaEventSynchronize(scale_done.v[n]));
            for (int i = 0; i < nouts; i++) {
                if (h32)
                    PFX_CHECK(staged_d2h(h32[i] + so * elems[i], d32[i] + so * elems[i], nxy * elems[i] * sizeof(float),
                                         pl->s_copy));
                else
                    PFX_CHECK(staged_d2h(h_outs[i] + so * elems[i], d_outs[i] + so * elems[i],
                                         nxy * elems[i] * sizeof(double), pl->s_copy));
            }
        }
    }
    if (h_outs || h32) {
        PFX_CUDA(cudaEventRecord(pl->ev_join, pl->s_copy));
        PFX_CUDA(cudaStreamWaitEvent(pl->stream, pl->ev_join, 0));
    }
    return PFX_OK;
}

// Host wrapper: stage the inputs, run, stream the outputs back scale by scale, synchronise.
int run_cwt_host(pfx_plan *pl, int mode, const double *g1, const double *g2, int is0, int is1, double *const *outs,
                 const int *out_elems_per_cell, int nouts, float *const *outs32 = nullptr)
{
    DeviceGuard guard(pl->device);
    pl->call_id++;
    const size_t nxy = (size_t)pl->nx * pl->ny;
    const int nsc = is1 - is0;
    double *d_g[2] = {nullptr, nullptr};
    PFX_CHECK(pl->scratch(0, nxy * sizeof(double), (void **)&d_g[0]));
    PFX_CUDA(cudaMemcpyAsync(d_g[0], g1, nxy * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    if (g2) {
        PFX_CHECK(pl->scratch(1, nxy * sizeof(double), (void **)&d_g[1]));
        PFX_CUDA(cudaMemcpyAsync(d_g[1], g2, nxy * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    }
    double *d_o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < nouts && !outs32; i++)
        PFX_CHECK(pl->scratch(2 + i, nxy * nsc * out_elems_per_cell[i] * sizeof(double), (void **)&d_o[i]));
    float *d_32[4] = {nullptr, nullptr, nullptr, nullptr};
    if (outs32)
        for (int i = 0; i < nouts; i++)
            PFX_CHECK(pl->scratch(6 + i, nxy * nsc * out_elems_per_cell[i] * sizeof(float), (void **)&d_32[i]));
    PFX_CHECK(run_cwt(pl, mode, d_g[0], d_g[1], is0, is1, d_o, outs32 ? nullptr : outs, out_elems_per_cell, nouts, nullptr,
                      outs32 ? d_32 : nullptr, outs32));
    PFX_CUDA(cudaStreamSynchronize(pl->stream));
    return PFX_OK;
}

}  // namespace

extern "C" {

int pfx_wlet_transform_dev(pfx_plan *pl, const double *grid, int is0, int is1, double *cube)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && cube, "null pointer");
    DeviceGuard guard(pl->device);
    double *outs[1] = {cube};
    const int el[1] = {2 * PFX_NA};
    return run_cwt(pl, OUT_CUBE, grid, nullptr, is0, is1, outs, nullptr, el, 1);
}

// Cube to the host, one scale slab (nx*ny*11 complex) at a time through two device slabs: the whole cube need
// not fit on the device, and the transform of scale s+1 runs while slab s crosses PCIe (classes.py:225 keeps
// the full cube in host memory).  c64: slabs and host cube are complex64, the dtype
// numpy.f2py gives the reference's COMPLEX array.
static int transform_to_host(pfx_plan *pl, const double *grid, int is0, int is1, void *cube, bool c64)
{
    DeviceGuard guard(pl->device);
    pl->call_id++;
    const size_t nxy = (size_t)pl->nx * pl->ny, slab_el = nxy * PFX_NA;
    const size_t slab_bytes = slab_el * (c64 ? sizeof(float2) : sizeof(double2));
    double *d_g = nullptr;
    void *d_slab[2] = {nullptr, nullptr};
    PFX_CHECK(pl->scratch(0, nxy * sizeof(double), (void **)&d_g));
    PFX_CHECK(pl->scratch(2, slab_bytes, &d_slab[0]));
    PFX_CHECK(pl->scratch(3, slab_bytes, &d_slab[1]));
    PFX_CUDA(cudaMemcpyAsync(d_g, grid, nxy * sizeof(double), cudaMemcpyHostToDevice, pl->stream));
    PFX_CHECK(pl->engine == PFX_ENGINE_PRUNED ? pruned_forward(pl, d_g, 0) : forward_spectrum(pl, d_g, 0));
    if (pl->engine == PFX_ENGINE_PRUNED) PFX_CHECK(pruned_prepare(pl, 1));
    auto enqueue = [&](int is, int b) -> int {  // planes of scale `is`, cropped into slab b
        PlaneView v;
        if (pl->engine == PFX_ENGINE_DENSE) {
            PFX_CHECK(dense_scale_planes(pl, is, 1));
            v = dense_view(pl, 0);
        } else {
            PFX_CHECK(pruned_scale_planes(pl, is, 1));
            v = pruned_view(pl, 0);
        }
        ProfScope prof(pl, PFX_KC_CROP, pl->stream);
        PFX_CHECK(launch_crop(pl->nx, pl->ny, v, d_slab[b], c64, pl->stream));
        pl->launches++;
        PFX_CUDA(cudaEventRecord(pl->ev_planes[b], pl->stream));
        return PFX_OK;
    };
    PFX_CHECK(enqueue(is0, 0));
    for (int is = is0; is < is1; is++) {
        const int b = (is - is0) & 1;
        if (is + 1 < is1) {
            // slab b^1 was last read by the copy of scale is-1
            if (is - is0 >= 1) PFX_CUDA(cudaStreamWaitEvent(pl->stream, pl->ev_reduced[b ^ 1], 0));
            PFX_CHECK(enqueue(is + 1, b ^ 1));
        }
        PFX_CUDA(cudaStreamWaitEvent(pl->s_copy, pl->ev_planes[b], 0));
        // (pageable cube: staged through pinned chunks, a thread team takes the first-touch faults)
        PFX_CHECK(staged_d2h(static_cast<char *>(cube) + (size_t)(is - is0) * slab_bytes, d_slab[b], slab_bytes, pl->s_copy));
        PFX_CUDA(cudaEventRecord(pl->ev_reduced[b], pl->s_copy));
    }
    PFX_CUDA(cudaStreamSynchronize(pl->s_copy));
    PFX_CUDA(cudaStreamSynchronize(pl->stream));
    return PFX_OK;
}

int pfx_wlet_transform_host(pfx_plan *pl, const double *grid, int is0, int is1, double *cube)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && cube, "null pointer");
    return transform_to_host(pl, grid, is0, is1, cube, false);
}

int pfx_wlet_transform_host_c64(pfx_plan *pl, const double *grid, int is0, int is1, float *cube)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && cube, "null pointer");
    return transform_to_host(pl, grid, is0, is1, cube, true);
}

int pfx_wlet_scalogram_dev(pfx_plan *pl, const double *grid, int is0, int is1, double *wl_sg, double *ewl_sg)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && wl_sg && ewl_sg, "null pointer");
    DeviceGuard guard(pl->device);
    double *outs[2] = {wl_sg, ewl_sg};
    const int el[2] = {1, 1};
    return run_cwt(pl, RED_SCALOGRAM, grid, nullptr, is0, is1, outs, nullptr, el, 2);
}

int pfx_wlet_scalogram_host(pfx_plan *pl, const double *grid, int is0, int is1, double *wl_sg, double *ewl_sg)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && wl_sg && ewl_sg, "null pointer");
    double *outs[2] = {wl_sg, ewl_sg};
    const int el[2] = {1, 1};
    return run_cwt_host(pl, RED_SCALOGRAM, grid, nullptr, is0, is1, outs, el, 2);
}

int pfx_wlet_xscalogram_dev(pfx_plan *pl, const double *g1, const double *g2, int is0, int is1, double *wl_xsg,
                            double *ewl_xsg)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(g1 && g2 && wl_xsg && ewl_xsg, "null pointer");
    DeviceGuard guard(pl->device);
    double *outs[2] = {wl_xsg, ewl_xsg};
    const int el[2] = {2, 1};
    return run_cwt(pl, RED_XSCALOGRAM, g1, g2, is0, is1, outs, nullptr, el, 2);
}

int pfx_wlet_xscalogram_host(pfx_plan *pl, const double *g1, const double *g2, int is0, int is1, double *wl_xsg,
                             double *ewl_xsg)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(g1 && g2 && wl_xsg && ewl_xsg, "null pointer");
    double *outs[2] = {wl_xsg, ewl_xsg};
    const int el[2] = {2, 1};
    return run_cwt_host(pl, RED_XSCALOGRAM, g1, g2, is0, is1, outs, el, 2);
}

int pfx_wlet_admit_coh_dev(pfx_plan *pl, const double *topo, const double *grav, int is0, int is1, double *wl_admit,
                           double *ewl_admit, double *wl_coh, double *ewl_coh)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(topo && grav && wl_admit && ewl_admit && wl_coh && ewl_coh, "null pointer");
    DeviceGuard guard(pl->device);
    double *outs[4] = {wl_admit, ewl_admit, wl_coh, ewl_coh};
    const int el[4] = {1, 1, 1, 1};
    return run_cwt(pl, RED_ADMIT_COH, topo, grav, is0, is1, outs, nullptr, el, 4);
}

int pfx_wlet_admit_coh_host(pfx_plan *pl, const double *topo, const double *grav, int is0, int is1, double *wl_admit,
                            double *ewl_admit, double *wl_coh, double *ewl_coh)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(topo && grav && wl_admit && ewl_admit && wl_coh && ewl_coh, "null pointer");
    double *outs[4] = {wl_admit, ewl_admit, wl_coh, ewl_coh};
    const int el[4] = {1, 1, 1, 1};
    return run_cwt_host(pl, RED_ADMIT_COH, topo, grav, is0, is1, outs, el, 4);
}

// ---- float32 outputs: the dtype numpy.f2py gives the reference's REAL arrays ---------------
int pfx_wlet_scalogram_host_f32(pfx_plan *pl, const double *grid, int is0, int is1, float *wl_sg, float *ewl_sg)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(grid && wl_sg && ewl_sg, "null pointer");
    float *outs[2] = {wl_sg, ewl_sg};
    const int el[2] = {1, 1};
    return run_cwt_host(pl, RED_SCALOGRAM, grid, nullptr, is0, is1, nullptr, el, 2, outs);
}

int pfx_wlet_admit_coh_host_f32(pfx_plan *pl, const double *topo, const double *grav, int is0, int is1, float *wl_admit,
                                float *ewl_admit, float *wl_coh, float *ewl_coh)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(topo && grav && wl_admit && ewl_admit && wl_coh && ewl_coh, "null pointer");
    float *outs[4] = {wl_admit, ewl_admit, wl_coh, ewl_coh};
    const int el[4] = {1, 1, 1, 1};
    return run_cwt_host(pl, RED_ADMIT_COH, topo, grav, is0, is1, nullptr, el, 4, outs);
}

// ---- row-sharded epilogue and IPC-mappable buffers ---------------------------------------
int pfx_wlet_admit_coh_sharded_dev(pfx_plan *pl, const double *topo, const double *grav, int is0, int is1, int nshards,
                                   void *const *shard_bases)
{
    PFX_CHECK(check_range(pl, is0, is1));
    PFX_ARG(topo && grav && shard_bases, "null pointer");
    PFX_ARG(nshards >= 1 && nshards <= PFX_MAX_SHARDS, "nshards must be in [1, PFX_MAX_SHARDS]");
    ShardOut sh;
    sh.nshards = nshards;
    sh.rows_per_shard = (pl->ny + nshards - 1) / nshards;
    sh.ns = pl->ns;
    for (int q = 0; q < nshards; q++) {
        const bool has_rows = q * sh.rows_per_shard < pl->ny;
        PFX_ARG(shard_bases[q] || !has_rows, "null shard buffer");
        sh.base[q] = reinterpret_cast<double *>(shard_bases[q]);
    }
    DeviceGuard guard(pl->device);
    const int el[4] = {1, 1, 1, 1};
    return run_cwt(pl, RED_ADMIT_COH, topo, grav, is0, is1, nullptr, nullptr, el, 4, &sh);
}

int pfx_wlet_admit_coh_sharded_list_dev(pfx_plan *pl, const double *topo, const double *grav, const int *scales,
                                        int nscales, int nshards, void *const *shard_bases)
{
    PFX_ARG(pl && topo && grav && scales && shard_bases, "null pointer");
    PFX_ARG(nscales >= 1 && nscales <= pl->ns, "nscales out of range");
    PFX_ARG(nshards >= 1 && nshards <= PFX_MAX_SHARDS, "nshards must be in [1, PFX_MAX_SHARDS]");
    for (int n = 0; n < nscales; n++) PFX_ARG(scales[n] >= 0 && scales[n] < pl->ns, "scale index out of range");
    ShardOut sh;
    sh.nshards = nshards;
    sh.rows_per_shard = (pl->ny + nshards - 1) / nshards;
    sh.ns = pl->ns;
    for (int q = 0; q < nshards; q++) {
        const bool has_rows = q * sh.rows_per_shard < pl->ny;
        PFX_ARG(shard_bases[q] || !has_rows, "null shard buffer");
        sh.base[q] = reinterpret_cast<double *>(shard_bases[q]);
    }
    DeviceGuard guard(pl->device);
    const int el[4] = {1, 1, 1, 1};
    return run_cwt(pl, RED_ADMIT_COH, topo, grav, 0, nscales, nullptr, nullptr, el, 4, &sh, nullptr, nullptr, scales);
}

size_t pfx_shard_bytes(int nx, int ny, int ns, int nshards, int q)
{
    if (nx < 1 || ny < 1 || ns < 1 || nshards < 1 || q < 0 || q >= nshards) return 0;
    const int R = (ny + nshards - 1) / nshards;
    const int lo = q * R < ny ? q * R : ny, hi = (q + 1) * R < ny ? (q + 1) * R : ny;
    return (size_t)4 * ns * (hi - lo) * nx * sizeof(double);
}

int pfx_ipc_alloc(int device, size_t bytes, void **dev_ptr, unsigned char *handle)
{
    PFX_ARG(dev_ptr && handle, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) <= PFX_IPC_HANDLE_BYTES, "IPC handle does not fit");
    DeviceGuard guard(device);
    *dev_ptr = nullptr;
    PFX_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr);
        *dev_ptr = nullptr;
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return PFX_ERR_CUDA;
    }
    memset(handle, 0, PFX_IPC_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return PFX_OK;
}

int pfx_ipc_free(int device, void *dev_ptr)
{
    DeviceGuard guard(device);
    if (dev_ptr) PFX_CUDA(cudaFree(dev_ptr));
    return PFX_OK;
}

int pfx_ipc_open(int device, const unsigned char *handle, void **peer_ptr)
{
    PFX_ARG(handle && peer_ptr, "null pointer");
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    *peer_ptr = nullptr;
    PFX_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PFX_OK;
}

int pfx_ipc_close(int device, void *peer_ptr)
{
    DeviceGuard guard(device);
    if (peer_ptr) PFX_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return PFX_OK;
}

// ---- cube-input forms -----------------------------------------------------------------
static int cube_reduce_host(int device, int mode, int nx, int ny, int ns, const double *w1, const double *w2, double *o0,
                            double *o1, double *o2, double *o3)
{
    PFX_ARG(nx >= 1 && ny >= 1 && ns >= 1, "bad cube shape");
    int ndev = 0;
    PFX_CUDA(cudaGetDeviceCount(&ndev));
    PFX_ARG(device >= 0 && device < ndev, "device ordinal out of range");
    DeviceGuard guard(device);
    const size_t nxy = (size_t)nx * ny, slab = nxy * PFX_NA;
    double2 *d_w[2] = {nullptr, nullptr};
    double *d_o[4] = {nullptr, nullptr, nullptr, nullptr};
    const int e0 = (mode == RED_XSCALOGRAM) ? 2 : 1;
    int rc = PFX_OK;
    auto cleanup = [&]() {
        for (auto p : d_w) cudaFree(p);
        for (auto p : d_o) cudaFree(p);
    };
#define CUBE_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
            cleanup();                                                                       \
            return e_ == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;             \
        }                                                                                    \
    } while (0)
    CUBE_CUDA(cudaMalloc(&d_w[0], slab * sizeof(double2)));
    if (w2) CUBE_CUDA(cudaMalloc(&d_w[1], slab * sizeof(double2)));
    CUBE_CUDA(cudaMalloc(&d_o[0], nxy * e0 * sizeof(double)));
    CUBE_CUDA(cudaMalloc(&d_o[1], nxy * sizeof(double)));
    if (mode == RED_ADMIT_COH) {
        CUBE_CUDA(cudaMalloc(&d_o[2], nxy * sizeof(double)));
        CUBE_CUDA(cudaMalloc(&d_o[3], nxy * sizeof(double)));
    }
    for (int is = 0; is < ns && rc == PFX_OK; is++) {
        CUBE_CUDA(cudaMemcpyAsync(d_w[0], w1 + 2 * slab * is, slab * sizeof(double2), cudaMemcpyHostToDevice, 0));
        if (w2) CUBE_CUDA(cudaMemcpyAsync(d_w[1], w2 + 2 * slab * is, slab * sizeof(double2), cudaMemcpyHostToDevice, 0));
        PlaneView v1{d_w[0], nxy, nx, 0, 0, 1.0, nullptr, 0, {}}, v2{w2 ? d_w[1] : d_w[0], nxy, nx, 0, 0, 1.0, nullptr, 0, {}};
        rc = launch_reduce(mode, nx, ny, v1, v2, d_o[0], d_o[1], d_o[2], d_o[3], 0);
        if (rc != PFX_OK) break;
        CUBE_CUDA(cudaMemcpyAsync(o0 + nxy * e0 * is, d_o[0], nxy * e0 * sizeof(double), cudaMemcpyDeviceToHost, 0));
        CUBE_CUDA(cudaMemcpyAsync(o1 + nxy * is, d_o[1], nxy * sizeof(double), cudaMemcpyDeviceToHost, 0));
        if (mode == RED_ADMIT_COH) {
            CUBE_CUDA(cudaMemcpyAsync(o2 + nxy * is, d_o[2], nxy * sizeof(double), cudaMemcpyDeviceToHost, 0));
            CUBE_CUDA(cudaMemcpyAsync(o3 + nxy * is, d_o[3], nxy * sizeof(double), cudaMemcpyDeviceToHost, 0));
        }
        CUBE_CUDA(cudaStreamSynchronize(0));
    }
#undef CUBE_CUDA
    cleanup();
    return rc;
}

int pfx_cube_scalogram_host(int device, int nx, int ny, int ns, const double *wl_trans, double *wl_sg, double *ewl_sg)
{
    PFX_ARG(wl_trans && wl_sg && ewl_sg, "null pointer");
    return cube_reduce_host(device, RED_SCALOGRAM, nx, ny, ns, wl_trans, nullptr, wl_sg, ewl_sg, nullptr, nullptr);
}

int pfx_cube_xscalogram_host(int device, int nx, int ny, int ns, const double *w1, const double *w2, double *wl_xsg,
                             double *ewl_xsg)
{
    PFX_ARG(w1 && w2 && wl_xsg && ewl_xsg, "null pointer");
    return cube_reduce_host(device, RED_XSCALOGRAM, nx, ny, ns, w1, w2, wl_xsg, ewl_xsg, nullptr, nullptr);
}

int pfx_cube_admit_coh_host(int device, int nx, int ny, int ns, const double *w1, const double *w2, double *wl_admit,
                            double *ewl_admit, double *wl_coh, double *ewl_coh)
{
    PFX_ARG(w1 && w2 && wl_admit && ewl_admit && wl_coh && ewl_coh, "null pointer");
    return cube_reduce_host(device, RED_ADMIT_COH, nx, ny, ns, w1, w2, wl_admit, ewl_admit, wl_coh, ewl_coh);
}

}  // extern "C"
