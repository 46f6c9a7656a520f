// pfx_pruned_v4_p1.cu -- instantiations and launcher of the fourth-generation (radix-16) pass 1 kernels.
#include "pfx_pruned_v4.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(v4::V4Args);

template <int LOGL>
KernelFn pick(int multi)
{
    if (multi & 2) return (multi & 1) ? v4::pass1_kernel<LOGL, true, true> : v4::pass1_kernel<LOGL, false, true>;
    return (multi & 1) ? v4::pass1_kernel<LOGL, true, false> : v4::pass1_kernel<LOGL, false, false>;
}

KernelFn kernel_for(int logL, int multi)  // multi | (dct << 1)
{
    switch (logL) {
        case 5: return pick<5>(multi);
        case 6: return pick<6>(multi);
        case 7: return pick<7>(multi);
        case 8: return pick<8>(multi);
        case 9: return pick<9>(multi);
        case 10: return pick<10>(multi);
        case 11: return pick<11>(multi);
        default: return nullptr;
    }
}

size_t smem_for(int logL)
{
    switch (logL) {
        case 5: return v4::smem_bytes<5>();
        case 6: return v4::smem_bytes<6>();
        case 7: return v4::smem_bytes<7>();
        case 8: return v4::smem_bytes<8>();
        case 9: return v4::smem_bytes<9>();
        case 10: return v4::smem_bytes<10>();
        case 11: return v4::smem_bytes<11>();
        default: return 0;
    }
}
}  // namespace

int v4_configure_p1(int logL, int multi, int dct, int *resident)
{
    KernelFn f = kernel_for(logL, (multi ? 1 : 0) | (dct ? 2 : 0));
    PFX_ARG(f != nullptr, "pruned v4 pass 1: unsupported FFT length");
    const size_t smem = smem_for(logL);
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (resident) PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, f, v4::NT, smem));
    return PFX_OK;
}

int v4_launch_p1(const v4::V4Args &a, int logL, int multi, int dct, dim3 grid, cudaStream_t st)
{
    KernelFn f = kernel_for(logL, (multi ? 1 : 0) | (dct ? 2 : 0));
    PFX_ARG(f != nullptr, "pruned v4 pass 1: unsupported FFT length");
    f<<<grid, v4::NT, smem_for(logL), st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
