#include <cstdio>
#include <cuda_runtime.h>
// A: a = fma(a, x, y) (two shared operands)   B: a[j] = fma(b[j], c[j], a[j]) three distinct register operands
// C: butterfly-like: complex multiply-add chains with 4-way ILP
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double x, double y)
{
    double a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = (threadIdx.x + j) * 1e-3; b[j] = x + j * 1e-9; c[j] = y + j * 1e-9; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (MODE == 0) a[j] = fma(a[j], x, y);
                if (MODE == 1) a[j] = fma(b[j], c[j], a[j]);
                if (MODE == 2) a[j] = fma(b[(j + r) & 7], c[(j + 2 * r + 1) & 7], a[j]);
                if (MODE == 3) { a[j] = fma(b[j], c[(j+r)&7], a[j]); b[j] = fma(a[(j+1)&7], c[j], b[j]); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j] + b[j] + c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int occ_blocks)
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int blocks = nsm * occ_blocks, iters = 4096; double *out; cudaMalloc(&out, blocks * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30;
    for (int rep = 0; rep < 4; rep++) { cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms; }
    double per = (MODE == 3) ? 64.0 : 32.0;
    printf("%s blocks/SM %d: %.2f TFLOP/s\n", name, occ_blocks, 2.0 * per * iters * blocks * 256 / (best * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    for (int occ : {8, 3, 2, 1}) { run<0>("A shared operands   ", occ); run<1>("B 3 distinct regs    ", occ); run<2>("C rotating operands  ", occ); run<3>("D dependent pairs    ", occ); }
    return 0;
}
