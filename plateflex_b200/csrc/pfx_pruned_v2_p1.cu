// pfx_pruned_v2_p1.cu -- instantiations and launcher of the second-generation pass-1 kernels.
#include "pfx_pruned_v2.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(v2::V2Args);

// variant = cpc | (prefetch << 4) | (multi << 5) | (dct << 6); two lines per CTA only without MULTI / PREFETCH
// (registers); dct: the spectrum is read in its cosine-transform form (pfx_dct.cu), single line, no prefetch
template <int LOGL>
KernelFn pick(int variant)
{
    const int cpc = variant & 15, pre = (variant >> 4) & 1, multi = (variant >> 5) & 1, dct = (variant >> 6) & 1;
    if (dct) {
        if (cpc != 1 || pre) return nullptr;
        return multi ? v2::pass1_kernel<LOGL, 1, true, false, true> : v2::pass1_kernel<LOGL, 1, false, false, true>;
    }
    if (cpc == 2) return (pre || multi) ? nullptr : v2::pass1_kernel<LOGL, 2, false, false>;
    if (multi) return pre ? v2::pass1_kernel<LOGL, 1, true, true> : v2::pass1_kernel<LOGL, 1, true, false>;
    return pre ? v2::pass1_kernel<LOGL, 1, false, true> : v2::pass1_kernel<LOGL, 1, false, false>;
}

KernelFn kernel_for(int logL, int variant)
{
    switch (logL) {
        case 5: return pick<5>(variant);
        case 6: return pick<6>(variant);
        case 7: return pick<7>(variant);
        case 8: return pick<8>(variant);
        case 9: return pick<9>(variant);
        case 10: return pick<10>(variant);
        case 11: return pick<11>(variant);
        default: return nullptr;
    }
}
}  // namespace

int v2_configure_p1(int logL, int variant, int *resident)
{
    KernelFn f = kernel_for(logL, variant);
    PFX_ARG(f != nullptr, "pruned v2 pass 1: unsupported FFT length or variant");
    const size_t smem = v2_smem_bytes(logL, variant & 15);
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (resident) PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, f, v2::NT, smem));
    return PFX_OK;
}

int v2_launch_p1(const v2::V2Args &a, int logL, int variant, dim3 grid, cudaStream_t st)
{
    KernelFn f = kernel_for(logL, variant);
    PFX_ARG(f != nullptr, "pruned v2 pass 1: unsupported FFT length or variant");
    f<<<grid, v2::NT, v2_smem_bytes(logL, variant & 15), st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

size_t v2_smem_bytes(int logL, int cpc)
{
    switch (logL) {
        case 5: return v2::smem_bytes<5>(cpc);
        case 6: return v2::smem_bytes<6>(cpc);
        case 7: return v2::smem_bytes<7>(cpc);
        case 8: return v2::smem_bytes<8>(cpc);
        case 9: return v2::smem_bytes<9>(cpc);
        case 10: return v2::smem_bytes<10>(cpc);
        case 11: return v2::smem_bytes<11>(cpc);
        default: return 0;
    }
}

}  // namespace pfx
