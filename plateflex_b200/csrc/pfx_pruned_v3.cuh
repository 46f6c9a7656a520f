// pfx_pruned_v3.cuh -- third-generation pass 2 (x-direction transforms) of the band-pruned engine.
//
// Same mathematics and the same shared-memory FFT as pfx_pruned_v2.cuh; what changes is how the K support
// samples of a line reach the first stage.  In v2 every thread fetched its eight samples from T[m][j] with
// register loads: 16-byte reads at a stride of ny*16 bytes, each one a half-used sector, and their latency
// (the top stall in the ncu source view) sat in front of every line because all eight warps of a CTA move in
// lock step between barriers.  Here
//   * pass 1 writes the intermediate transposed, T[t][j][slot]: one contiguous line of slots per output row,
//     slot = wavenumber index modulo L, padded (n + n/8) for L >= 512 exactly as the first stage wants to
//     read it from shared memory without bank conflicts;
//   * one elected thread fetches the support range of the NEXT line(s) with cp.async.bulk (the 1-D form of
//     TMA: UBLKCP in SASS) into a ring of NBUF line buffers, completion signalled on an mbarrier
//     (SYNCS), while the eight warps run the stages of the current line -- no registers, no address
//     arithmetic per sample, whole 128-byte lines from L2 / HBM;
//   * the samples no longer live in registers across the parts of a long line (N > 2048), so those kernels
//     are back at 3 CTAs per SM;
//   * the first stage folds its input twiddles w_N^(a_k (rho+1)) = t0 g^k into a twiddled radix-8
//     butterfly (v2::tw_dft), which removes the serial chain of seven complex multiplies.
#pragma once
#include "pfx_pruned_v2.cuh"

namespace pfx {
namespace v3 {

using v2::Geo;
using v2::NT;
using v2::padded;
using v2::tw_dft;
using v2::bitrev;
using prn::cmul;

struct V3Args {
    PrunedArgs p;
    int logr;    // log2(N / L) of the transformed axis
    int nparts;  // N / CH
    int nbuf;    // line buffers in the ring (1 or 2), each Geo3::LP slots
};

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int LOGL>
struct Geo3 : Geo<LOGL> {
    using G = Geo<LOGL>;
    static constexpr int LP = G::PAD ? (G::L + (G::L >> 3)) : G::L;  // slots per line of T (and of a full buffer)
};

// later stages without the leading barrier of v2::run_stages (the caller has just synchronised)
template <int LOGL, int S, bool HALF, class Store>
__device__ __forceinline__ void later_stages(double2 *sm, const double2 *stw, const Store &store)
{
    v2::stage<LOGL, S, 1, HALF, Store>(sm, stw, store);
    v2::run_stages<LOGL, S + 1, 1, HALF, Store>(sm, stw, store);
}

// Pass 2: transform t = blockIdx.y, output rows j = blockIdx.x, + gridDim.x, ...  W[t][j][i], i < nx.
// HALF: power-of-two grid (N = 2 nx): outputs q >= L/2 are never stored, so the last stage skips them and
// every stored index is inside the row.
template <int LOGL, bool MULTI, bool HALF, int OCC>
__global__ void __launch_bounds__(NT, OCC) pass2_kernel(V3Args a)
{
    using G = Geo3<LOGL>;
    constexpr int L = G::L, STEP = G::STEP, LP = G::LP;
    constexpr int STEP_P = G::PAD ? STEP + (STEP >> 3) : STEP;  // padded(n + STEP) - padded(n)
    static_assert(G::IPT0 == 1, "one first-stage butterfly per thread (L <= 2048)");
    static_assert(!G::PAD || (STEP & 7) == 0, "padding is additive over first-stage input strides");
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ax;
    const int t = blockIdx.y, ia = t % PFX_NA;
    const int lo = sd.alo[ia], K = sd.ka[ia];
    const int ny = a.p.ny, nx = a.p.nx, logr = a.logr;
    const int nparts = MULTI ? a.nparts : 1;

    double2 *stw = smem + G::PLANE;
    double2 *adv = stw + G::STW;                 // MULTI: per-thread twiddle advances (t0, t0 h) and the common one of g
    double2 *bufs = adv + (MULTI ? 2 * NT + 1 : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bufs + (size_t)a.nbuf * LP);
    v2::load_stage_twiddles<LOGL>(stw, A.logN, a.p.rootx);
    // Slots outside the support range are never written by the copies: zero the line buffers once, so that the
    // first stage needs no "inside the support" test.
    for (int i = threadIdx.x; i < a.nbuf * LP; i += NT) bufs[i] = make_double2(0., 0.);
    if (threadIdx.x == 0) {
        for (int b = 0; b < a.nbuf; b++) mbar_init(bars + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zeros before the bulk copies (async proxy)

    // this thread's first-stage butterfly (u, rho): fixed for the whole kernel
    const int Nm = A.N - 1;
    const int rho = threadIdx.x & (G::RC - 1), u = threadIdx.x >> G::LOGRC;
    const int nlow = prn::digit_reverse<LOGL>(u);  // natural index (wavenumber index mod L) of input 0
    const int m0 = (nlow - lo) & (L - 1);          // its support index
    const int a0 = lo + m0;                        // its wavenumber index
    // inputs k >= kw have wrapped modulo L: their wavenumber index is a0 + k STEP - L
    const int kw = (L - m0 + STEP - 1) >> G::LOGBF;
    double2 *const out0 = smem + (padded<G::PAD>(u << 3) << G::LOGRC) + rho;
    const int s0 = padded<G::PAD>(nlow);  // slot of input 0; input k sits STEP_P slots further
    // twiddles: t0 = w_N^(a0 (rho+1)), its wrapped form t0 h with h = w_N^(-L (rho+1)), g = w_N^(STEP (rho+1))
    const double2 t0i = __ldg(a.p.rootx + (int)(((long long)a0 * (rho + 1)) & Nm));
    const double2 g0i = __ldg(a.p.rootx + (((rho + 1) << G::LOGBF) & Nm));
    const double2 t0hi = cmul(t0i, __ldg(a.p.rootx + ((A.N - (((rho + 1) << LOGL) & Nm)) & Nm)));
    if (MULTI) {
        // advance from part to part (rho -> rho + RC), kept in shared memory: registers are what limits occupancy
        const double2 dt0 = __ldg(a.p.rootx + (int)(((long long)a0 * G::RC) & Nm));
        const double2 dh = __ldg(a.p.rootx + ((A.N - ((G::RC << LOGL) & Nm)) & Nm));
        adv[threadIdx.x] = dt0;
        adv[NT + threadIdx.x] = cmul(dt0, dh);
        if (threadIdx.x == 0) adv[2 * NT] = __ldg(a.p.rootx + ((G::RC << G::LOGBF) & Nm));
    }
    __syncthreads();

    // support range in slot space: n in [n0, n0 + K) modulo L -> padded slots [pa0, pa1) and, when the range
    // wraps, [0, pb1); both segments land at their own slots of the buffer.
    const double2 *Tt = a.p.T + (size_t)t * ny * LP;  // T[t][j][slot]
    double2 *Wt = a.p.W + (size_t)t * ny * nx;
    auto issue = [&](int j, int b) {  // one thread: fetch the support range of row j into buffer b
        const int n0 = lo & (L - 1);
        const int nA = min(n0 + K, L);  // end of the first segment
        const int pa0 = padded<G::PAD>(n0), pa1 = padded<G::PAD>(nA - 1) + 1;
        const int pb1 = (n0 + K > L) ? padded<G::PAD>(n0 + K - L - 1) + 1 : 0;
        const uint32_t bytesA = (uint32_t)(pa1 - pa0) * 16u, bytesB = (uint32_t)pb1 * 16u;
        uint64_t *bar = bars + b;
        double2 *dst = bufs + b * LP;
        const double2 *src = Tt + (size_t)j * LP;
        mbar_arrive_expect_tx(bar, bytesA + bytesB);
        bulk_g2s(dst + pa0, src + pa0, bytesA, bar);
        if (bytesB) bulk_g2s(dst, src, bytesB, bar);
    };
    if (threadIdx.x == 0) {
        int j = blockIdx.x;
        for (int b = 0; b < a.nbuf && j < ny; b++, j += gridDim.x) issue(j, b);
    }

    int it = 0;
    for (int j = blockIdx.x; j < ny; j += gridDim.x, it++) {
        const int b = (a.nbuf == 2) ? (it & 1) : 0;
        const uint32_t parity = (a.nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
        const double2 *x = bufs + b * LP + s0;
        mbar_wait(bars + b, parity);
        double2 t0 = t0i, t0h = t0hi, g = g0i;
        double2 *Wrow = Wt + (size_t)j * nx;
        for (int part = 0; part < nparts; part++) {
            // ---- first stage: inputs from the line buffer, twiddled radix-8 butterfly into the plane
            {
                double2 v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = cmul(x[k * STEP_P], (k >= kw) ? t0h : t0);
                tw_dft<8, false>(v, g);
#pragma unroll
                for (int k = 0; k < 8; k++) out0[k << G::LOGRC] = v[bitrev<8>(k)];
            }
            if (MULTI && part + 1 < nparts) {  // rho -> rho + RC
                t0 = cmul(t0, adv[threadIdx.x]);
                t0h = cmul(t0h, adv[NT + threadIdx.x]);
                g = cmul(g, adv[2 * NT]);
            }
            __syncthreads();
            if (part == nparts - 1 && threadIdx.x == 0) {
                // every warp has read buffer b for the last time: refill it with the line nbuf iterations ahead
                const int jn = j + a.nbuf * gridDim.x;
                if (jn < ny) issue(jn, b);
            }
            const int rho0 = part << G::LOGRC;
            auto store = [&](int, int q, int rr, double2 val) {
                const int i = (q << logr) + rho0 + rr;
                if (HALF || i < nx) Wrow[i] = val;
            };
            later_stages<LOGL, 1, HALF>(smem, stw, store);
            __syncthreads();  // the next first stage overwrites the plane
        }
    }
}

// Pass 1: transform t = blockIdx.y, spectrum columns m = blockIdx.x, + gridDim.x, ... (wavenumber index alo + m):
// the y-direction transform of v1(b) X[a][b] over the support rows, outputs j < ny, times (c/P) u1(a), into T.
// The spectrum is y-fastest, so the support range of a column is contiguous: the same bulk-copy ring as pass 2
// fetches the next column(s) while this one is transformed (the v2 kernel's global loads were 56 % of its stall
// samples).  The buffer is the raw spectrum (slot = wavenumber index mod L, no padding); the v1 weights sit in
// shared memory in the same order.
template <int LOGL, bool MULTI, bool HALF, int OCC>
__global__ void __launch_bounds__(NT, OCC) pass1_kernel(V3Args a)
{
    using G = Geo3<LOGL>;
    constexpr int L = G::L, STEP = G::STEP;
    extern __shared__ double2 smem[];
    const ScaleDesc &sd = a.p.sd;
    const AxisDesc &A = sd.ay;
    const int t = blockIdx.y, ia = t % PFX_NA, gsel = t / PFX_NA;
    const int lo = sd.blo[ia], K = sd.kb[ia], ka = sd.ka[ia], alo = sd.alo[ia];
    if ((int)blockIdx.x >= ka) return;
    const int ny = a.p.ny, logr = a.logr, NYm = a.p.NY - 1;
    const int nparts = MULTI ? a.nparts : 1;

    double2 *stw = smem + G::PLANE;
    double2 *adv = stw + G::STW;
    double *wv = reinterpret_cast<double *>(adv + (MULTI ? 2 * NT + 1 : 0));  // v1 weights in slot order [L]
    double2 *bufs = reinterpret_cast<double2 *>(wv + L);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bufs + (size_t)a.nbuf * L);
    v2::load_stage_twiddles<LOGL>(stw, A.logN, a.p.rooty);
    {
        const double *tv = a.p.tabv + sd.tabv + ((size_t)ia << LOGL);  // by support index, 0 beyond the support
        for (int n = threadIdx.x; n < L; n += NT) wv[n] = tv[(n - lo) & (L - 1)];
    }
    for (int i = threadIdx.x; i < a.nbuf * L; i += NT) bufs[i] = make_double2(0., 0.);
    if (threadIdx.x == 0) {
        for (int b = 0; b < a.nbuf; b++) mbar_init(bars + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    const int Nm = A.N - 1;
    const int rho = threadIdx.x & (G::RC - 1), u = threadIdx.x >> G::LOGRC;
    const int nlow = prn::digit_reverse<LOGL>(u);
    const int m0 = (nlow - lo) & (L - 1);
    const int a0 = lo + m0;
    const int kw = (L - m0 + STEP - 1) >> G::LOGBF;
    double2 *const out0 = smem + (padded<G::PAD>(u << 3) << G::LOGRC) + rho;
    const double2 t0i = __ldg(a.p.rooty + (int)(((long long)a0 * (rho + 1)) & Nm));
    const double2 g0i = __ldg(a.p.rooty + (((rho + 1) << G::LOGBF) & Nm));
    const double2 t0hi = cmul(t0i, __ldg(a.p.rooty + ((A.N - (((rho + 1) << LOGL) & Nm)) & Nm)));
    if (MULTI) {
        const double2 dt0 = __ldg(a.p.rooty + (int)(((long long)a0 * G::RC) & Nm));
        const double2 dh = __ldg(a.p.rooty + ((A.N - ((G::RC << LOGL) & Nm)) & Nm));
        adv[threadIdx.x] = dt0;
        adv[NT + threadIdx.x] = cmul(dt0, dh);
        if (threadIdx.x == 0) adv[2 * NT] = __ldg(a.p.rooty + ((G::RC << G::LOGBF) & Nm));
    }
    __syncthreads();

    const double2 *spec = gsel ? a.p.spec1 : a.p.spec0;  // padded spectrum [a][b], y-wavenumber fastest
    const double *tu = a.p.tabu + sd.tabu + (size_t)ia * sd.ax.L;
    auto issue = [&](int m, int b) {  // one thread: the support range of column m -> buffer b, slot = index mod L
        const int n0 = lo & (L - 1);
        const int lenA = min(K, L - n0), lenB = K - lenA;
        const double2 *src = spec + (size_t)((alo + m) & (a.p.NX - 1)) * (NYm + 1);
        uint64_t *bar = bars + b;
        double2 *dst = bufs + b * L;
        mbar_arrive_expect_tx(bar, (uint32_t)K * 16u);
        bulk_g2s(dst + n0, src + (lo & NYm), (uint32_t)lenA * 16u, bar);
        if (lenB) bulk_g2s(dst, src + ((lo + lenA) & NYm), (uint32_t)lenB * 16u, bar);
    };
    if (threadIdx.x == 0) {
        int m = blockIdx.x;
        for (int b = 0; b < a.nbuf && m < ka; b++, m += gridDim.x) issue(m, b);
    }

    const double *wvp = wv + nlow;
    int it = 0;
    for (int m = blockIdx.x; m < ka; m += gridDim.x, it++) {
        const int b = (a.nbuf == 2) ? (it & 1) : 0;
        const uint32_t parity = (a.nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
        const double2 *x = bufs + b * L + nlow;
        mbar_wait(bars + b, parity);
        double2 t0 = t0i, t0h = t0hi, g = g0i;
        size_t tjs;
        double2 *Tcol = a.p.T + prn_t_column(a.p, t, m, alo, tjs);
        const double wu = tu[m];
        for (int part = 0; part < nparts; part++) {
            {
                double2 v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const double2 xin = x[k * STEP];
                    const double w = wvp[k * STEP];
                    v[k] = cmul(make_double2(xin.x * w, xin.y * w), (k >= kw) ? t0h : t0);
                }
                tw_dft<8, false>(v, g);
#pragma unroll
                for (int k = 0; k < 8; k++) out0[k << G::LOGRC] = v[bitrev<8>(k)];
            }
            if (MULTI && part + 1 < nparts) {
                t0 = cmul(t0, adv[threadIdx.x]);
                t0h = cmul(t0h, adv[NT + threadIdx.x]);
                g = cmul(g, adv[2 * NT]);
            }
            __syncthreads();
            if (part == nparts - 1 && threadIdx.x == 0) {
                const int mn = m + a.nbuf * gridDim.x;
                if (mn < ka) issue(mn, b);
            }
            const int rho0 = part << G::LOGRC;
            auto store = [&](int, int q, int rr, double2 val) {
                const int j = (q << logr) + rho0 + rr;
                if (HALF || j < ny) Tcol[(size_t)j * tjs] = make_double2(val.x * wu, val.y * wu);
            };
            later_stages<LOGL, 1, HALF>(smem, stw, store);
            __syncthreads();
        }
    }
}

template <int LOGL>
inline size_t smem_bytes(int multi, int nbuf)
{
    return ((size_t)Geo3<LOGL>::PLANE + Geo3<LOGL>::STW + (multi ? 2 * NT + 1 : 0) + (size_t)nbuf * Geo3<LOGL>::LP) * sizeof(double2) +
           2 * sizeof(uint64_t);
}

template <int LOGL>
inline size_t smem_bytes_p1(int multi, int nbuf)
{
    return ((size_t)Geo3<LOGL>::PLANE + Geo3<LOGL>::STW + (multi ? 2 * NT + 1 : 0) + Geo3<LOGL>::L / 2 +
            (size_t)nbuf * Geo3<LOGL>::L) * sizeof(double2) + 2 * sizeof(uint64_t);
}

}  // namespace v3

// launchers (pfx_pruned_v3_p2.cu, pfx_pruned_v3_p1.cu)
int v3_line_slots(int logL);  // LP: slots per line of the transposed T for this FFT length
int v3_pad(int logL);         // 1 when slots are padded (n + n/8)
size_t v3_smem_bytes(int logL, int multi, int nbuf);
// occ: CTAs per SM the kernel is compiled for (3: 80 registers, 2: 128 registers)
int v3_configure_p2(int logL, int multi, int half, int occ, int nbuf, int *resident);
int v3_launch_p2(const v3::V3Args &a, int logL, int multi, int half, int occ, dim3 grid, cudaStream_t st);
size_t v3_smem_bytes_p1(int logL, int multi, int nbuf);
int v3_configure_p1(int logL, int multi, int half, int occ, int nbuf, int *resident);
int v3_launch_p1(const v3::V3Args &a, int logL, int multi, int half, int occ, dim3 grid, cudaStream_t st);

}  // namespace pfx
