// pfx_pruned_v2h_p2.cu -- instantiations and launcher of the 128-thread form of the second-generation pass-2
// kernels (namespace v2h of pfx_pruned_v2.cuh: line parts of 1024 elements, FFT lengths 32..1024).
#include "pfx_pruned_v2.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(v2h::V2Args);

template <int LOGL>
KernelFn pick(int multi)
{
    return multi ? v2h::pass2_kernel<LOGL, 1, true, false> : v2h::pass2_kernel<LOGL, 1, false, false>;
}

KernelFn kernel_for(int logL, int multi)
{
    switch (logL) {
        case 5: return pick<5>(multi);
        case 6: return pick<6>(multi);
        case 7: return pick<7>(multi);
        case 8: return pick<8>(multi);
        case 9: return pick<9>(multi);
        case 10: return pick<10>(multi);
        default: return nullptr;
    }
}

size_t smem_for(int logL)
{
    switch (logL) {
        case 5: return v2h::smem_bytes<5>(1);
        case 6: return v2h::smem_bytes<6>(1);
        case 7: return v2h::smem_bytes<7>(1);
        case 8: return v2h::smem_bytes<8>(1);
        case 9: return v2h::smem_bytes<9>(1);
        case 10: return v2h::smem_bytes<10>(1);
        default: return 0;
    }
}
}  // namespace

int v2h_configure_p2(int logL, int multi, int *resident)
{
    KernelFn f = kernel_for(logL, multi);
    PFX_ARG(f != nullptr, "pruned v2h pass 2: unsupported FFT length");
    const size_t smem = smem_for(logL);
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (resident) PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, f, v2h::NT, smem));
    return PFX_OK;
}

int v2h_launch_p2(const v2h::V2Args &a, int logL, int multi, dim3 grid, cudaStream_t st)
{
    KernelFn f = kernel_for(logL, multi);
    PFX_ARG(f != nullptr, "pruned v2h pass 2: unsupported FFT length");
    f<<<grid, v2h::NT, smem_for(logL), st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
