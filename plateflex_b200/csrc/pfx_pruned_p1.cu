// pfx_pruned_p1.cu -- instantiations and launcher of the pass-1 kernels (y-direction transforms).
#include "pfx_pruned_kernels.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(PrunedArgs);

template <int LOGL>
KernelFn pick(int pad)
{
    if constexpr (LOGL >= 8) {
        if (pad) return prn::pass1_kernel<LOGL, true>;
    }
    return prn::pass1_kernel<LOGL, false>;
}

KernelFn kernel_for(int logL, int pad)
{
    switch (logL) {
        case 1: return pick<1>(pad);
        case 2: return pick<2>(pad);
        case 3: return pick<3>(pad);
        case 4: return pick<4>(pad);
        case 5: return pick<5>(pad);
        case 6: return pick<6>(pad);
        case 7: return pick<7>(pad);
        case 8: return pick<8>(pad);
        case 9: return pick<9>(pad);
        case 10: return pick<10>(pad);
        case 11: return pick<11>(pad);
        case 12: return pick<12>(pad);
        case 13: return pick<13>(pad);
        case 14: return pick<14>(pad);
        default: return nullptr;
    }
}
}  // namespace

int prn_configure_pass1(int logL, int pad, size_t smem)
{
    KernelFn f = kernel_for(logL, pad);
    PFX_ARG(f != nullptr, "pruned pass 1: unsupported FFT length");
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return PFX_OK;
}

int prn_launch_pass1(const PrunedArgs &a, dim3 grid, size_t smem, cudaStream_t st)
{
    KernelFn f = kernel_for(a.sd.ay.logL, a.sd.ay.pad);
    PFX_ARG(f != nullptr, "pruned pass 1: unsupported FFT length");
    f<<<grid, PRN_NT, smem, st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
