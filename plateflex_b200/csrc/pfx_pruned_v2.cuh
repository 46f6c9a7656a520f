// pfx_pruned_v2.cuh -- second-generation kernels of the band-pruned engine for padded axes of
// N >= 2048 samples (grids from 512 cells per side up).  Same mathematics as pfx_pruned_kernels.cuh
// (see pfx_cwt_pruned.cu), different mapping:
//
//   * a line of N = L*r output samples is produced in parts of CH = max(2048, L) elements, each part the
//     L-point FFTs of RC = CH/L consecutive residues.  Parts are independent given the K support samples,
//     so chunking costs no arithmetic and keeps a line part at 32 KB of shared memory whatever N is
//     (3 CTAs of 256 threads per SM instead of 1 at N = 8192);
//   * with CH fixed every thread owns the same radix-8 butterfly (u, rho) in every part of every line:
//     all shared-memory indices, strides and digit reversals are compile-time constants, and the K
//     support samples a thread needs are the same for every part, so they are fetched from global
//     memory once per line straight into registers -- optionally one line group ahead (PREFETCH) so the
//     load latency hides behind the remaining stages of the current group;
//   * (v2h) the same kernels with 128 threads per CTA and line parts of 1024 elements: twice as many, smaller
//     barrier domains per SM (pfx_pruned_v2h_p2.cu), chosen per scale by the planner;
//   * the first-stage twiddles w_N^(a (rho+1)) advance from part to part by one complex multiply
//     (rho -> rho + RC) instead of table reads.
#pragma once
#include "pfx_pruned.hpp"
#include "pfx_pruned_kernels.cuh"

#define PFX_V2_NS v2
#define PFX_V2_NT 256
#define PFX_V2_LOGCH0 11
#include "pfx_pruned_v2_body.inc"
#undef PFX_V2_NS
#undef PFX_V2_NT
#undef PFX_V2_LOGCH0
#define PFX_V2_NS v2h
#define PFX_V2_NT 128
#define PFX_V2_LOGCH0 10
#include "pfx_pruned_v2_body.inc"
#undef PFX_V2_NS
#undef PFX_V2_NT
#undef PFX_V2_LOGCH0

namespace pfx {

// launchers (pfx_pruned_v2_p1.cu, pfx_pruned_v2_p2.cu).  variant = cpc | (prefetch << 4) | (multi << 5)
// resident: receives the number of CTAs of this kernel that fit on one SM
int v2_configure_p1(int logL, int variant, int *resident);
int v2_configure_p2(int logL, int variant, int *resident);
int v2_launch_p1(const v2::V2Args &a, int logL, int variant, dim3 grid, cudaStream_t st);
int v2_launch_p2(const v2::V2Args &a, int logL, int variant, dim3 grid, cudaStream_t st);
size_t v2_smem_bytes(int logL, int cpc);
// 128-thread form of pass 2 (FFT lengths up to 1024)
int v2h_configure_p2(int logL, int multi, int *resident);
int v2h_launch_p2(const v2h::V2Args &a, int logL, int multi, dim3 grid, cudaStream_t st);

}  // namespace pfx
