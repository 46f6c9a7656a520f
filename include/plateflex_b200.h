/*
 * plateflex_b200.h -- C ABI of libplateflex_b200.so
 *
 * The drop-in boundary for PlateFlex's data-parallel hot path on B200
 * (sm_100a).  It replaces the two numpy.f2py extension modules that the
 * reference builds from Fortran (reference meson.build:46-75):
 *
 *     plateflex.cpwt   (src/cpwt/cpwt.f90, src/cpwt/cpwt_sub.f90)
 *     plateflex.flex   (src/flex/flex.f90)
 *
 * plus the per-cell SciPy fit loop of plateflex/estimate.py:227-409 driven by
 * plateflex/classes.py:1125-1332, which has no native counterpart in the
 * reference.  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every function returns PFX_OK (0) or a PFX_ERR_* code; the message is
 *     available from pfx_last_error() (thread-local).
 *   - "_host" entry points take HOST pointers and perform the host<->device
 *     copies themselves; "_dev" entry points take DEVICE pointers that live on
 *     the plan's device and enqueue all work on the plan's stream
 *     (pfx_plan_set_stream) without synchronising.
 *   - input grids are (nx, ny) float64, C order (y fastest), exactly what
 *     numpy hands to f2py in classes.py:225.
 *   - all outputs use the reference's Fortran layout (x fastest):
 *         maps   (nx, ny, ns)        -> out[ix + nx*(iy + ny*is)]
 *         cubes  (nx, ny, 11, ns)    -> out[ix + nx*(iy + ny*(ia + 11*is))]
 *     complex values are interleaved (re, im) float64 pairs.
 *   - all arithmetic is float64.  The reference is float32 (default-kind REAL,
 *     cpwt.f90:36,41,72-74); float64 outputs are a superset.
 */
#ifndef PLATEFLEX_B200_H
#define PLATEFLEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFX_NA 11 /* azimuths of the fan wavelet, cpwt.f90:37 */
#define PFX_MAX_SHARDS 16 /* row shards (GPUs) a sharded epilogue can address */

enum pfx_status {
    PFX_OK = 0,
    PFX_ERR_ARG = 1,   /* bad argument (shape, null pointer, range) */
    PFX_ERR_CUDA = 2,  /* CUDA runtime error */
    PFX_ERR_CUFFT = 3, /* cuFFT error */
    PFX_ERR_OOM = 4,   /* device allocation failed */
    PFX_ERR_STATE = 5  /* call sequence error */
};

enum pfx_engine {
    PFX_ENGINE_PRUNED = 0, /* hand-written band-pruned inverse transform (default) */
    PFX_ENGINE_DENSE = 1   /* daughter multiply + batched cuFFT Z2Z inverse */
};

enum pfx_atype { /* estimate.py:267,308,349 */
    PFX_ATYPE_ADMIT = 0,
    PFX_ATYPE_COH = 1,
    PFX_ATYPE_JOINT = 2
};

typedef struct pfx_plan pfx_plan; /* opaque CWT plan: sizes, tables, workspaces */

/* conf_flex module variables (flex.f90:47-48).  rhof is not here: the Fortran
 * overwrites it from wd on every call (flex.f90:207-211) and so do we.  wd,
 * and optionally rhoc / zc, are per-cell (classes.py:1087-1091) and passed as
 * arrays where relevant. */
typedef struct pfx_flex_conf {
    double rhoc; /* crustal density   (default 2700, __init__.py:153) */
    double rhom; /* mantle density    (default 3200) */
    double rhow; /* water density     (default 1030) */
    double rhoa; /* air density       (default 0) */
    double zc;   /* crustal thickness (default 35e3 m) */
    int32_t boug; /* 1 = Bouguer (A=0), 0 = free air (A=1), flex.f90:200-204 */
    int32_t pad_;
} pfx_flex_conf;

/* ---- library ---------------------------------------------------------- */
int pfx_version(void);
const char *pfx_last_error(void);
int pfx_device_count(int *count);

/* cpwt_sub.f90:26-33 (npow2): smallest power of two >= n, at least 2. */
int pfx_npow2(int n);

/* Measured FP64 peak of the device: a kernel of independent DFMA chains (8 per thread, all SMs full), timed
 * with CUDA events; *tflops = 2 * fused multiply-adds / second / 1e12.  The denominator of the FP64 rooflines
 * bench.py reports (the driver's MEASURED_PEAKS.json holds no FP64 figure). */
int pfx_fp64_peak(int device, double *tflops);

/* ---- device buffers ---------------------------------------------------- */
/* For callers that have no device allocator of their own: the Python layer keeps the four (nx,ny,ns)
 * maps of Project.wlet_admit_coh (classes.py:916-975) on the device so that Project.estimate_grid
 * (classes.py:1125-1332) fits them in place, and copies a map to the host only when its attribute is
 * read.  Copies are synchronous; pageable host memory (plain numpy arrays) is staged through pinned
 * chunks by a small thread team. */
int pfx_dev_alloc(int device, size_t bytes, void **dev_ptr);
int pfx_dev_free(int device, void *dev_ptr);
int pfx_dev_h2d(int device, void *dst_dev, const void *src_host, size_t bytes);
/* Upload of a C-order (rows, cols) host array of 8- or 1-byte elements as its Fortran-order image
 * (dst[j*rows + i] = src[i*cols + j]): the (nx, ny) grids numpy hands over (water depth, rhoc, zc, mask --
 * classes.py:1087-1091, 1228) reach the x-fastest layout of the maps without a strided copy on the host. */
int pfx_dev_h2d_transposed(int device, void *dst_dev, const void *src_host, int rows, int cols, int elem_bytes);
int pfx_dev_d2h(int device, void *dst_host, const void *src_dev, size_t bytes);

/* ---- CWT plan --------------------------------------------------------- */
/* Fixes (nx, ny, dx, dy, k[ns], k0) -- the arguments of cpwt.wlet_transform
 * (cpwt.f90:65) other than the data -- allocates device workspaces and builds
 * the wavenumber / daughter tables (defk cpwt_sub.f90:41-62, wave_function
 * cpwt_sub.f90:73-99).  dx, dy in km; k in rad/m (kf, cpwt.f90:148); k0 is
 * conf_cpwt.k0 (cpwt.f90:41).  device = CUDA ordinal. */
int pfx_plan_create(pfx_plan **plan, int nx, int ny, double dx_km, double dy_km, const double *k_rad_m, int ns,
                    double k0, int engine, int device);
int pfx_plan_destroy(pfx_plan *plan);
/* stream is a cudaStream_t (NULL = legacy default stream). */
int pfx_plan_set_stream(pfx_plan *plan, void *stream);
/* padded sizes npow2(2nx), npow2(2ny) (cpwt.f90:94-95) and bytes of device memory held. */
int pfx_plan_info(const pfx_plan *plan, int *nnx, int *nny, size_t *device_bytes);
/* number of kernels / library launches this plan has enqueued so far (for bench.py's gpu_launches). */
int pfx_plan_launch_count(const pfx_plan *plan, long long *kernels);
/* Relative cost estimate of each scale (cost[ns], arbitrary units): small scales have wide k-space
 * supports and cost more.  Lets a multi-GPU caller cut the scale axis into balanced blocks. */
int pfx_plan_scale_cost(const pfx_plan *plan, double *cost);
/* Per-kernel device timers.  When enabled, every launch of the plan's kernels is bracketed by a CUDA
 * event pair on the stream it is launched on; pfx_plan_profile_read waits for the recorded spans,
 * returns the summed milliseconds and launch counts per kernel class and clears them.  While enabled
 * the plan does not overlap kernels on its internal streams, so each span times one kernel alone.
 * Classes (index): 0 forward spectrum, 1 admissibility transform, 2 pruned pass 1 (y transforms),
 * 3 pruned pass 2 (x transforms), 4 azimuth reduction + jackknife, 5 dense daughter multiply,
 * 6 dense inverse cuFFT, 7 wl_trans crop. */
#define PFX_KERNEL_CLASSES 8
int pfx_plan_profile_enable(pfx_plan *plan, int on);
int pfx_plan_profile_read(pfx_plan *plan, double *ms, long long *count, int n);
/* pruned engine: support half-width of the daughter in units of its k-space sigma (default 9.5). */
int pfx_plan_set_cutoff(pfx_plan *plan, double nsigma);

/* ---- cpwt.wlet_transform  (cpwt.f90:65-191) --------------------------- */
/* grid (nx,ny) -> cube (nx,ny,11,ns) complex.  Scales [is0, is1) only; the
 * output holds is1-is0 scale slabs, the first one being scale is0. */
int pfx_wlet_transform_host(pfx_plan *plan, const double *grid, int is0, int is1, double *cube);
int pfx_wlet_transform_dev(pfx_plan *plan, const double *grid_dev, int is0, int is1, double *cube_dev);

/* ---- cpwt.wlet_scalogram  (cpwt.f90:200-241), fused with the transform - */
int pfx_wlet_scalogram_host(pfx_plan *plan, const double *grid, int is0, int is1, double *wl_sg, double *ewl_sg);
int pfx_wlet_scalogram_dev(pfx_plan *plan, const double *grid_dev, int is0, int is1, double *wl_sg_dev,
                           double *ewl_sg_dev);

/* ---- cpwt.wlet_xscalogram (cpwt.f90:250-293), fused -------------------- */
int pfx_wlet_xscalogram_host(pfx_plan *plan, const double *grid1, const double *grid2, int is0, int is1,
                             double *wl_xsg /* complex */, double *ewl_xsg);
int pfx_wlet_xscalogram_dev(pfx_plan *plan, const double *grid1_dev, const double *grid2_dev, int is0, int is1,
                            double *wl_xsg_dev, double *ewl_xsg_dev);

/* ---- cpwt.wlet_admit_coh  (cpwt.f90:303-364), fused -------------------- */
/* grid1 = topography, grid2 = gravity (classes.py:966-967). */
int pfx_wlet_admit_coh_host(pfx_plan *plan, const double *topo, const double *grav, int is0, int is1, double *wl_admit,
                            double *ewl_admit, double *wl_coh, double *ewl_coh);
int pfx_wlet_admit_coh_dev(pfx_plan *plan, const double *topo_dev, const double *grav_dev, int is0, int is1,
                           double *wl_admit_dev, double *ewl_admit_dev, double *wl_coh_dev, double *ewl_coh_dev);

/* ---- float32 outputs ---------------------------------------------------- */
/* The reference's REAL arrays reach Python as float32 (numpy.f2py).  These forms compute in
 * float64 like everything else and round each result once on the device, halving the device-to-host
 * copy that bounds the host-pointer calls (671 MB -> 336 MB per 1024^2 x 20 transform). */
int pfx_wlet_transform_host_c64(pfx_plan *plan, const double *grid, int is0, int is1, float *cube /* complex64 */);
int pfx_wlet_scalogram_host_f32(pfx_plan *plan, const double *grid, int is0, int is1, float *wl_sg, float *ewl_sg);
int pfx_wlet_admit_coh_host_f32(pfx_plan *plan, const double *topo, const double *grav, int is0, int is1,
                                float *wl_admit, float *ewl_admit, float *wl_coh, float *ewl_coh);

/* ---- multi-GPU: admittance / coherence written straight into row shards ---- */
/* The scale-sharded CWT of one GPU feeding the row-sharded fit of all GPUs without a
 * separate exchange step: the epilogue kernel of every scale in [is0, is1) stores its four output rows
 * directly into the buffer of the shard that owns each grid row -- shard q owns rows
 * [q*R, min((q+1)*R, ny)), R = ceil(ny / nshards) -- at the position pfx_l2_fit_rows_dev expects:
 *     shard_bases[q][ ((map*ns + scale)*rows_q + (row - q*R))*nx + x ],  map = admit, eadmit, coh, ecoh.
 * shard_bases[q] is a device pointer valid on this GPU: local memory for the own shard, memory of a peer
 * GPU mapped with pfx_ipc_open for the others (plain stores over NVLink; the transfer overlaps the
 * transform scale by scale).  Every rank must synchronise its stream and meet the others at a barrier
 * before any shard is read. */
int pfx_wlet_admit_coh_sharded_dev(pfx_plan *plan, const double *topo_dev, const double *grav_dev, int is0, int is1,
                                   int nshards, void *const *shard_bases);
/* Same for an arbitrary set of scales (nscales indices into the plan's wavenumbers, transformed in the given
 * order): with the one-sided exchange a rank's scales need not be contiguous, so they can be dealt out by
 * cost (pfx_plan_scale_cost) instead of cut into blocks. */
int pfx_wlet_admit_coh_sharded_list_dev(pfx_plan *plan, const double *topo_dev, const double *grav_dev,
                                        const int *scales, int nscales, int nshards, void *const *shard_bases);
/* bytes of shard q's buffer: 4 * ns * rows_q * nx * sizeof(double) */
size_t pfx_shard_bytes(int nx, int ny, int ns, int nshards, int q);
/* Device memory that other processes on the node can map (cudaMalloc + cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle).  handle is PFX_IPC_HANDLE_BYTES opaque bytes to be passed between processes. */
#define PFX_IPC_HANDLE_BYTES 64
int pfx_ipc_alloc(int device, size_t bytes, void **dev_ptr, unsigned char *handle);
int pfx_ipc_free(int device, void *dev_ptr);
int pfx_ipc_open(int device, const unsigned char *handle, void **peer_ptr);
int pfx_ipc_close(int device, void *peer_ptr);

/* ---- cube-input forms: the literal f2py signatures --------------------- */
/* For callers that already hold wl_trans cubes (classes.py:277, 966).  Host
 * pointers; streamed through the device one scale slab at a time. */
int pfx_cube_scalogram_host(int device, int nx, int ny, int ns, const double *wl_trans, double *wl_sg, double *ewl_sg);
int pfx_cube_xscalogram_host(int device, int nx, int ny, int ns, const double *wl_trans1, const double *wl_trans2,
                             double *wl_xsg, double *ewl_xsg);
int pfx_cube_admit_coh_host(int device, int nx, int ny, int ns, const double *wl_trans1, const double *wl_trans2,
                            double *wl_admit, double *ewl_admit, double *wl_coh, double *ewl_coh);

/* ---- flex.real_xspec_functions (flex.f90:178-235), batched ------------- */
/* ncurves parameter sets (te[c] km, f[c], alpha[c] rad, wd[c] m) evaluated on
 * k[ns] (rad/m): admit/coh are [ncurves][ns] row-major.  wd may be NULL (0).
 * jac_admit / jac_coh (may be NULL) receive the analytic derivatives with
 * respect to (Te, F, alpha) that the fit uses, as [ncurves][3][ns]. */
int pfx_flex_xspec_host(int device, const pfx_flex_conf *conf, const double *k, int ns, int ncurves, const double *te,
                        const double *f, const double *alpha, const double *wd, double *admit, double *coh,
                        double *jac_admit, double *jac_coh);

/* ---- helper routines that numpy.f2py also exports (last rows) ------------------ */
/* flex.f90:74-169 on ns-vectors: op 0 flexfilter_top(psi) -> o0; 1 flexfilter_bot(psi) -> o0; 2 decon(theta =
 * in0, phi = in1, k = in2, A = p0) -> mu_h, mu_w, nu_h, nu_w; 3 tr_func(mu_h, mu_w, nu_h, nu_w, F = p0, alpha = p1)
 * -> Re admit, Im admit, coh.  rhof and wd are the conf_flex module values (flex.f90:47-48). */
int pfx_flex_helper_host(int device, const pfx_flex_conf *conf, double rhof, double wd, int op, int ns, const double *in0,
                         const double *in1, const double *in2, const double *in3, double p0, double p1, double *o0,
                         double *o1, double *o2, double *o3);
/* cpwt_sub.f90:41-62 (defk) and :73-99 (wave_function, (nx, ny) complex, x fastest) in float64. */
int pfx_defk_host(int device, int nx, int ny, double dx, double dy, double *kx, double *ky);
int pfx_wave_function_host(int device, int nx, int ny, const double *kx, const double *ky, double k0, double scale,
                           double angle, double *wave);

/* ---- per-cell L2 fit: estimate.L2_estimate_cell over a grid ------------ */
/* Replaces the Python loop of Project.estimate_grid (classes.py:1218-1304)
 * and the SciPy call of estimate.py:274-280 (+5 siblings).
 *
 * Inputs are (nx,ny,ns) maps in the layout above; wd is (nx,ny) (x fastest)
 * and required; rhoc_map / zc_map / mask may be NULL (conf values / no mask;
 * mask is uint8, non-zero = skip, classes.py:1228-1231).  Cells (i, j) with
 * i in range(0, nx-nn, nn), j in range(0, ny-nn, nn) are fitted
 * (classes.py:1218-1219); outputs are (mx, my) = (nx/nn, ny/nn) maps, x
 * fastest, zero where not fitted (classes.py:1189-1204).  out_alpha /
 * out_alpha_std may be NULL when alph == 0.  status (int32, may be NULL):
 * low byte 0 not fitted, 1 converged, 2 iteration limit, 3 bad data (sigma<=0 / NaN:
 * where SciPy would raise, estimate.py:274 -> ValueError); the upper 24 bits hold the
 * number of model evaluations the cell used (least_squares' nfev).
 */
int pfx_l2_fit_host(int device, const pfx_flex_conf *conf, const double *k, int ns, int nx, int ny,
                    const double *wl_admit, const double *ewl_admit, const double *wl_coh, const double *ewl_coh,
                    const double *wd, const double *rhoc_map, const double *zc_map, const uint8_t *mask, int nn,
                    int alph, int atype, double *out_te, double *out_te_std, double *out_f, double *out_f_std,
                    double *out_alpha, double *out_alpha_std, double *out_chi2, int32_t *status);
int pfx_l2_fit_dev(int device, void *stream, const pfx_flex_conf *conf, const double *k_dev, int ns, int nx, int ny,
                   const double *wl_admit, const double *ewl_admit, const double *wl_coh, const double *ewl_coh,
                   const double *wd, const double *rhoc_map, const double *zc_map, const uint8_t *mask, int nn,
                   int alph, int atype, double *out_te, double *out_te_std, double *out_f, double *out_f_std,
                   double *out_alpha, double *out_alpha_std, double *out_chi2, int32_t *status);

/* Row-sharded form for the multi-GPU inversion (cells shard by grid rows): the input
 * maps hold only grid rows [row0, row0+nrows) of a grid with ny_total rows -- (nx, nrows, ns) maps and
 * (nx, nrows) wd / rhoc / zc / mask -- and the call fits exactly the cells Project.estimate_grid
 * (classes.py:1218-1219) would visit inside that row range.  Outputs are (nx/nn, out_rows) maps
 * holding global output rows [first_out_row, first_out_row + out_rows) as given by
 * pfx_l2_fit_rows_shape; concatenating the shards' outputs along y gives the full-grid maps.
 * pfx_l2_fit_dev is this call with row0 = 0 and nrows = ny_total. */
int pfx_l2_fit_rows_shape(int ny_total, int row0, int nrows, int nn, int *first_out_row, int *out_rows);
int pfx_l2_fit_rows_dev(int device, void *stream, const pfx_flex_conf *conf, const double *k_dev, int ns, int nx,
                        int ny_total, int row0, int nrows, const double *wl_admit, const double *ewl_admit,
                        const double *wl_coh, const double *ewl_coh, const double *wd, const double *rhoc_map,
                        const double *zc_map, const uint8_t *mask, int nn, int alph, int atype, double *out_te,
                        double *out_te_std, double *out_f, double *out_f_std, double *out_alpha, double *out_alpha_std,
                        double *out_chi2, int32_t *status);

/* ---- the step in front of the path: grid ingest -------------------------- */
/* GMT ".xyz" dumps as the reference's notebooks read them (docs/getting_started.rst:50-56, Ex1 cells 2-3): a
 * first line "# name<TAB>xmin xmax ymin ymax zmin zmax dx dy nx ny" (gmt grdinfo -C) followed by one
 * "x y z" line per node.  pfx_xyz_header returns those ten numbers; pfx_xyz_read parses the nvalues = nx*ny
 * z values in file order (NaN accepted) with nthreads host threads (0 = automatic). */
int pfx_xyz_header(const char *path, double *hdr10);
int pfx_xyz_read(const char *path, long nvalues, double *z, int nthreads);
/* Grid.__init__ (classes.py:143-153): every non-finite cell of the (nx, ny) C-order grid takes the value of
 * the nearest finite cell (Euclidean distance in index units -- scipy.interpolate.griddata(method='nearest');
 * equidistant candidates: smallest row, then smallest column).  In place; *nfilled (optional) = cells filled. */
int pfx_nan_fill_nearest_host(int device, int nx, int ny, double *grid, long *nfilled);
/* TopoGrid.filter_water_depth (classes.py:538-541): skimage.filters.gaussian(image, sigma) for a float image,
 * i.e. scipy.ndimage.gaussian_filter(mode='nearest', truncate=4.0); (nx, ny) C order, in != out allowed. */
int pfx_gaussian_filter_host(int device, int nx, int ny, double sigma, const double *in, double *out);

#ifdef __cplusplus
}
#endif
#endif /* PLATEFLEX_B200_H */
