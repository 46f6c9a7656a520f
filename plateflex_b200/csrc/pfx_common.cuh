// pfx_common.cuh -- shared declarations of libplateflex_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "plateflex_b200.h"

namespace pfx {

void set_error(const char *fmt, ...);

// NVTX range around the host-side enqueue of one stage (header-only NVTX 3: a no-op unless a profiler injects
// its library).  Stage names: pfx:forward, pfx:prepare, pfx:scale<is>:passes, pfx:scale<is>:reduce, pfx:l2_fit, ...
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    NvtxRange(const char *prefix, int n, const char *suffix)
    {
        char buf[64];
        snprintf(buf, sizeof buf, "%s%d%s", prefix, n, suffix);
        nvtxRangePushA(buf);
    }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define PFX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            pfx::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            return e_ == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;                    \
        }                                                                                           \
    } while (0)

#define PFX_CUFFT(call)                                                                             \
    do {                                                                                            \
        cufftResult r_ = (call);                                                                    \
        if (r_ != CUFFT_SUCCESS) {                                                                  \
            pfx::set_error("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r_);        \
            return r_ == CUFFT_ALLOC_FAILED ? PFX_ERR_OOM : PFX_ERR_CUFFT;                          \
        }                                                                                           \
    } while (0)

#define PFX_CHECK(expr)                 \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != PFX_OK) return rc_;  \
    } while (0)

#define PFX_ARG(cond, msg)                                    \
    do {                                                      \
        if (!(cond)) {                                        \
            pfx::set_error("invalid argument: %s", msg);      \
            return PFX_ERR_ARG;                               \
        }                                                     \
    } while (0)

// A set of complex wavelet-coefficient planes W[p](i, j) as the reduction
// kernels see them: value = scale * base[p*plane_stride + (j+off_j)*ld + (i+off_i)].
// The dense engine hands over whole padded cuFFT planes (off = +1: the one-sample
// shift of cpwt.f90:173-185, scale = 1/(nnx*nny) of cpwt.f90:177); the pruned engine
// and the cube-input forms hand over compact (nx, ny) planes.  When zc is set the planes hold only the
// main Gaussian term of the daughter and the coefficient is  W[p](i,j) - ce[p] * Z(i,j)  with the real
// Z = zc[j*ld + i].x or .y (zsel: the grid)  -- the scale- and azimuth-independent admissibility term of
// cpwt_sub.f90:93-94, see pfx_cwt_pruned.cu.
struct PlaneView {
    const double2 *base;
    size_t plane_stride;
    int ld;
    int off_i, off_j;
    double scale;
    const double2 *zc;
    int zsel;
    double ce[PFX_NA];
};

enum ReduceMode { RED_SCALOGRAM = 0, RED_XSCALOGRAM = 1, RED_ADMIT_COH = 2 };

// Row-sharded destination of the admittance / coherence epilogue (multi-GPU): grid row j
// belongs to shard q = j / rows_per_shard, whose buffer -- possibly the memory of a peer GPU mapped through
// CUDA IPC, written with plain stores over NVLink -- holds four maps laid out [map][scale][row in shard][x]
// for all ns scales, i.e. exactly the (nx, rows, ns) input of pfx_l2_fit_rows_dev on that GPU.
struct ShardOut {
    int nshards = 0;  // 0: not sharded
    int rows_per_shard = 0;
    int ns = 0;      // scales of the whole transform (slab stride of the destination maps)
    int scale = 0;   // global index of the scale being reduced
    double *base[PFX_MAX_SHARDS] = {};
};

// K3: azimuth reduction + running-prefix jackknife for one scale.
//   RED_SCALOGRAM : w1 -> out0 = wl_sg, out1 = ewl_sg                       (cpwt.f90:200-241)
//   RED_XSCALOGRAM: w1,w2 -> out0 = wl_xsg (complex), out1 = ewl_xsg        (cpwt.f90:250-293)
//   RED_ADMIT_COH : w1,w2 -> out0..3 = wl_admit, ewl_admit, wl_coh, ewl_coh (cpwt.f90:303-364)
// Each out* points at the (nx, ny) plane of this scale (x fastest); with f32 they are float arrays.
int launch_reduce(int mode, int nx, int ny, PlaneView w1, PlaneView w2, double *out0, double *out1, double *out2,
                  double *out3, cudaStream_t stream, const ShardOut *shards = nullptr, bool f32 = false,
                  int *flag_word = nullptr);
// flag_word (RED_ADMIT_COH): a device int owned by the caller that no other launch in flight uses -- the "some cell
// needs the exact jackknife" word shared by the kernel and its fix-up pass; when null one is allocated and freed in
// stream order.

// Copies between the device and pageable host memory through pinned chunks and a thread team (pfx_hostio.cu)
bool host_is_pageable(const void *p);
int staged_d2h(void *dst, const void *src_dev, size_t bytes, cudaStream_t st);
int staged_h2d(void *dst_dev, const void *src, size_t bytes, cudaStream_t st);

// K4: copy 11 planes of a view into a compact (nx, ny, 11) cube slab of double2, or float2 with c64.
int launch_crop(int nx, int ny, PlaneView w, void *slab, bool c64, cudaStream_t stream);

}  // namespace pfx
