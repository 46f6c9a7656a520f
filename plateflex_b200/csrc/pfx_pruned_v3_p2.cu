// pfx_pruned_v3_p2.cu -- instantiations and launcher of the third-generation pass-2 kernels
// (bulk-copy staged support samples, pfx_pruned_v3.cuh).
#include "pfx_pruned_v3.cuh"

namespace pfx {

namespace {
using KernelFn = void (*)(v3::V3Args);

template <int LOGL, int OCC>
KernelFn pick(int multi, int half)
{
    if (multi) return half ? v3::pass2_kernel<LOGL, true, true, OCC> : v3::pass2_kernel<LOGL, true, false, OCC>;
    return half ? v3::pass2_kernel<LOGL, false, true, OCC> : v3::pass2_kernel<LOGL, false, false, OCC>;
}

template <int OCC>
KernelFn kernel_occ(int logL, int multi, int half)
{
    switch (logL) {
        case 5: return pick<5, OCC>(multi, half);
        case 6: return pick<6, OCC>(multi, half);
        case 7: return pick<7, OCC>(multi, half);
        case 8: return pick<8, OCC>(multi, half);
        case 9: return pick<9, OCC>(multi, half);
        case 10: return pick<10, OCC>(multi, half);
        case 11: return pick<11, OCC>(multi, half);
        default: return nullptr;
    }
}

KernelFn kernel_for(int logL, int multi, int half, int occ)
{
    return occ == 2 ? kernel_occ<2>(logL, multi, half) : kernel_occ<3>(logL, multi, half);
}
}  // namespace

int v3_line_slots(int logL)
{
    switch (logL) {
        case 5: return v3::Geo3<5>::LP;
        case 6: return v3::Geo3<6>::LP;
        case 7: return v3::Geo3<7>::LP;
        case 8: return v3::Geo3<8>::LP;
        case 9: return v3::Geo3<9>::LP;
        case 10: return v3::Geo3<10>::LP;
        case 11: return v3::Geo3<11>::LP;
        default: return 0;
    }
}

int v3_pad(int logL) { return logL >= 5 && logL <= 11 && v3_line_slots(logL) != (1 << logL); }

size_t v3_smem_bytes(int logL, int multi, int nbuf)
{
    switch (logL) {
        case 5: return v3::smem_bytes<5>(multi, nbuf);
        case 6: return v3::smem_bytes<6>(multi, nbuf);
        case 7: return v3::smem_bytes<7>(multi, nbuf);
        case 8: return v3::smem_bytes<8>(multi, nbuf);
        case 9: return v3::smem_bytes<9>(multi, nbuf);
        case 10: return v3::smem_bytes<10>(multi, nbuf);
        case 11: return v3::smem_bytes<11>(multi, nbuf);
        default: return 0;
    }
}

int v3_configure_p2(int logL, int multi, int half, int occ, int nbuf, int *resident)
{
    KernelFn f = kernel_for(logL, multi, half, occ);
    PFX_ARG(f != nullptr, "pruned v3 pass 2: unsupported FFT length");
    const size_t smem = v3_smem_bytes(logL, multi, nbuf);
    PFX_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (resident) PFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(resident, f, v3::NT, smem));
    return PFX_OK;
}

int v3_launch_p2(const v3::V3Args &a, int logL, int multi, int half, int occ, dim3 grid, cudaStream_t st)
{
    KernelFn f = kernel_for(logL, multi, half, occ);
    PFX_ARG(f != nullptr, "pruned v3 pass 2: unsupported FFT length");
    f<<<grid, v3::NT, v3_smem_bytes(logL, multi, a.nbuf), st>>>(a);
    PFX_CUDA(cudaGetLastError());
    return PFX_OK;
}

}  // namespace pfx
