// pfx_pruned_p2.cu -- instantiations and launcher of the pass-2 kernels (x-direction transforms),
// one or two output rows per CTA.
#include "pfx_pruned_kernels.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(PrunedArgs);

template <int LOGL>
KernelFn pick(int pad, int cpc)
{
    if constexpr (LOGL >= 8) {
        if (pad) return cpc == 2 ? prn::pass2_kernel<LOGL, 2, true> : prn::pass2_kernel<LOGL, 1, true>;
    }
    return cpc == 2 ? prn::pass2_kernel<LOGL, 2, false> : prn::pass2_kernel<LOGL, 1, false>;
}

KernelFn kernel_for(int logL, int pad, int cpc)
{
    switch (logL) {
        case 1: return pick<1>(pad, cpc);
        case 2: return pick<2>(pad, cpc);
        case 3: return pick<3>(pad, cpc);
        case 4: return pick<4>(pad, cpc);
        case 5: return pick<5>(pad, cpc);
        case 6: return pick<6>(pad, cpc);
        case 7: return pick<7>(pad, cpc);
        case 8: return pick<8>(pad, cpc);
        case 9: return pick<9>(pad, cpc);
        case 10: return pick<10>(pad, cpc);
        case 11: return pick<11>(pad, cpc);
        case 12: return pick<12>(pad, cpc);
        case 13: return pick<13>(pad, cpc);
        case 14: return pick<14>(pad, cpc);
        default: return nullptr;
    }
}
}  // namespace

int prn_configure_pass2(int logL, int pad, int cpc, size_t smem)
{
    KernelFn f = kernel_for(logL, pad, cpc);
    PFX_ARG(f != nullptr, "pruned pass 2: unsupported FFT length");
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return PFX_OK;
}

int prn_launch_pass2(const PrunedArgs &a, int cpc, dim3 grid, size_t smem, cudaStream_t st)
{
    KernelFn f = kernel_for(a.sd.ax.logL, a.sd.ax.pad, cpc);
    PFX_ARG(f != nullptr, "pruned pass 2: unsupported FFT length");
    f<<<grid, PRN_NT, smem, st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
