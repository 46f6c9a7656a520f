// pfx_plan.hpp -- the CWT plan object behind the opaque pfx_plan handle.
#pragma once
#include <vector>

#include "pfx_common.cuh"

// Per-(scale) constants of the fan of daughters (cpwt.f90:131,148-161; cpwt_sub.f90:85-87).
struct PfxScaleConsts {
    double s;              // scale = k0 / (k*1e3)   [km]
    double norm;           // s / sqrt(pi)
    double kx0[PFX_NA];    // k0*cos(angle)
    double ky0[PFX_NA];    // k0*sin(angle)
    double e0[PFX_NA];     // exp(-0.5*(kx0^2 + ky0^2))
};

// Kernel classes of the per-kernel device timers (pfx_plan_profile_*).
enum PfxKernelClass {
    PFX_KC_FORWARD = 0,  // mean + mirror/pad + forward cuFFT, per grid
    PFX_KC_PREPARE,      // admissibility transform Z (pruned engine), per call
    PFX_KC_PASS1,        // pruned engine, y-direction transforms of one scale
    PFX_KC_PASS2,        // pruned engine, x-direction transforms of one scale
    PFX_KC_REDUCE,       // azimuth reduction + jackknife of one scale
    PFX_KC_DENSE_MUL,    // dense engine: tables + daughter multiply of one scale
    PFX_KC_DENSE_FFT,    // dense engine: batched inverse cuFFT of one scale
    PFX_KC_CROP,         // wl_trans materialiser of one scale
    PFX_KC_COUNT
};

struct PfxProfSpan {
    int cls;
    cudaEvent_t a, b;
};

struct pfx_plan {
    int nx = 0, ny = 0, ns = 0, NX = 0, NY = 0;
    int engine = PFX_ENGINE_PRUNED, device = 0;
    double dx = 0, dy = 0, k0 = 0, cutoff = 9.5;
    std::vector<double> k;               // rad/m
    std::vector<PfxScaleConsts> consts;  // per scale
    cudaStream_t stream = nullptr;
    long long launches = 0;
    size_t device_bytes = 0;

    // wavenumber axes (defk, cpwt_sub.f90:41-62) and the scale-independent
    // admissibility factors exp(-kx^2/2), exp(-ky^2/2) (cpwt_sub.f90:93)
    double *d_kx = nullptr, *d_ky = nullptr, *d_u2 = nullptr, *d_v2 = nullptr;
    // padded spectra of up to two grids: layout [b][a] (a = x-wavenumber fastest) for the dense engine,
    // [a][b] (spec_t: y-wavenumber fastest, so that pass 1 of the pruned engine reads its columns
    // contiguously) for the pruned engine -- the 2-D transform of the transposed grid
    double2 *d_spec[2] = {nullptr, nullptr};
    bool spec_t = false;
    double *d_partial = nullptr;  // deterministic two-stage sum for the mean
    int *d_flag_words = nullptr;  // [2] flag words of the reduction, one per coefficient-plane buffer
    cufftHandle fft_fwd = 0;
    bool fft_fwd_ok = false;

    // dense engine: per-scale separable daughter tables u1[11][NX], v1[11][NY], and 22 work planes
    double *d_tab = nullptr;
    double2 *d_planes = nullptr;
    cufftHandle fft_inv[2] = {0, 0};  // batch 11 and batch 22
    bool fft_inv_ok[2] = {false, false};
    void *d_fft_work = nullptr;
    size_t fft_work_bytes = 0;

    void *pruned = nullptr;  // PrunedState of pfx_cwt_pruned.cu

    // Internal streams: the azimuth reduction of scale s runs on s_aux while the transform passes of
    // scale s+1 run on the caller's stream (two plane buffers); the host entry points drain finished
    // scales to the host on s_copy.  Everything is joined back into `stream` before a call returns.
    cudaStream_t s_aux = nullptr, s_copy = nullptr;
    cudaEvent_t ev_planes[2] = {nullptr, nullptr}, ev_reduced[2] = {nullptr, nullptr}, ev_join = nullptr;
    // Pruned engine: pass 1 of scale s+1 (short, latency-bound launches) runs on s_p1 while pass 2 of scale s runs
    // on the caller's stream; two intermediate buffers T, guarded by these events
    cudaStream_t s_p1 = nullptr;
    cudaEvent_t ev_t_ready[2] = {nullptr, nullptr}, ev_t_free[2] = {nullptr, nullptr}, ev_spec = nullptr;

    // grow-only scratch used by the _host entry points (inputs / outputs staging)
    void *d_scratch[10] = {};
    size_t scratch_bytes[10] = {};
    unsigned long scratch_call[10] = {};  // host call that last used the slot
    unsigned long call_id = 0;

    std::vector<void *> owned;  // everything to cudaFree on destroy

    // optional per-kernel device timers: an event pair around every launch of a class, on the stream
    // the kernel is launched on; read back (after a synchronise) by pfx_plan_profile_read
    bool profiling = false;
    std::vector<PfxProfSpan> spans;
    std::vector<cudaEvent_t> spare_events;
    cudaEvent_t prof_event();

    int alloc(void **p, size_t bytes);
    int scratch(int slot, size_t bytes, void **p);
};

namespace pfx {

// RAII event pair around the launches of one kernel class (no-op unless plan->profiling).
struct ProfScope {
    pfx_plan *pl;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    ProfScope(pfx_plan *p, int cls, cudaStream_t s) : pl(p), st(s)
    {
        if (!pl->profiling) return;
        cudaEvent_t a = pl->prof_event();
        b = pl->prof_event();
        if (!a || !b) {
            b = nullptr;
            return;
        }
        cudaEventRecord(a, st);
        pl->spans.push_back({cls, a, b});
    }
    ~ProfScope()
    {
        if (b) cudaEventRecord(b, st);
    }
};

// K1: mirror + de-mean + zero-pad (cpwt.f90:100-116, cpwt_sub.f90:435-449) then one forward FFT
// (cpwt.f90:122) of grid g (device pointer, C order) into plan->d_spec[slot].
int forward_spectrum(pfx_plan *pl, const double *grid_dev, int slot);
int grid_mean(pfx_plan *pl, const double *grid_dev, const double **mean_dev);

// Dense engine: all 11 azimuth planes of scale `is` for ngrids spectra -> plan->d_planes
// (plane p = ia + 11*g), as un-normalised, un-shifted inverse FFTs.
int dense_scale_planes(pfx_plan *pl, int is, int ngrids);
PlaneView dense_view(const pfx_plan *pl, int g);

int plan_init_dense(pfx_plan *pl);
int plan_init_axes(pfx_plan *pl);

// Pruned engine (pfx_cwt_pruned.cu)
int plan_init_pruned(pfx_plan *pl);
void plan_free_pruned(pfx_plan *pl);
int pruned_prepare(pfx_plan *pl, int ngrids);          // admissibility transform Z, once per set of spectra
// forward spectrum of grid `slot` for the pruned engine: the cosine-transform form (pfx_dct.cu) on power-of-two
// grids, else forward_spectrum
int pruned_forward(pfx_plan *pl, const double *grid_dev, int slot);
int pruned_scale_planes(pfx_plan *pl, int is, int ngrids);  // -> compact W planes of scale `is`
// the two halves of pruned_scale_planes on a chosen stream and intermediate buffer tb (0 | 1)
int pruned_pass1(pfx_plan *pl, int is, int ngrids, int tb, cudaStream_t stream);
int pruned_pass2(pfx_plan *pl, int is, int ngrids, int tb, cudaStream_t stream);
PlaneView pruned_view(const pfx_plan *pl, int g);
int pruned_num_buffers(const pfx_plan *pl);
double pruned_scale_cost(const pfx_plan *pl, int is);
void pruned_select_buffer(pfx_plan *pl, int b);

}  // namespace pfx
