// tests/v4_emul.cu -- host emulation of one line of the fourth-generation pass kernels (pfx_pruned_v4.cuh).
//
// Test infrastructure (CPU only): replays, thread by thread and stage by stage, exactly the index arithmetic of
// v4::run_lines / v4::stage with the kernel's own device/host helpers (radix plan, digit reversal, padding,
// dft16, tw_dft, bitrev), on a plain array standing in for shared memory, and compares the line with the direct
// sum  out[i] = sum_m x[m] w_N^{(lo+m)(i+1)}.  Built and run by tests/test_v4_emulation.py (nvcc, no GPU needed).
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../plateflex_b200/csrc/pfx_pruned_v4.cuh"

using namespace pfx;
using namespace pfx::v4;

template <int LOGL, int S, bool HALF>
static void emu_stage(std::vector<double2> &sm, const std::vector<double2> &stw_all, int logr, int rho0, std::vector<double2> &out,
                      std::vector<char> &seen)
{
    using G = Geo<LOGL>;
    constexpr int LR = lrad(LOGL, S), R = 1 << LR, LOGM = logm(LOGL, S), M = 1 << LOGM;
    constexpr bool LAST = (S == G::NST - 1);
    constexpr int NBF = (1 << LOGCH) >> LR;
    constexpr int IPT = NBF / NT;
    static_assert(IPT >= 1, "");
    const double2 *stw = stw_all.data() + stw_off(LOGL, S);
    for (int tid = 0; tid < NT; tid++)
        for (int it = 0; it < IPT; it++) {
            const int b = it * NT + tid;
            const int rho = b & (G::RC - 1);
            const int u = b >> G::LOGRC;
            const int q0 = u & (M - 1);
            const int base = ((u >> LOGM) << (LOGM + LR)) + q0;
            double2 pw[LR];
            pw[0] = stw[q0];
            for (int i = 1; i < LR; i++) pw[i] = csqr(pw[i - 1]);
            double2 *p = sm.data() + rho;
            double2 v[R];
            for (int k = 0; k < R; k++) v[k] = p[padded<G::PAD>(base + k * M) << G::LOGRC];
            tw_dft<R, HALF && LAST>(v, pw);
            const int nk = (HALF && LAST) ? R / 2 : R;
            for (int k = 0; k < nk; k++) {
                if (LAST) {
                    const int q = base + k * M;
                    const int i = (q << logr) + rho0 + rho;
                    out[i] = v[bitrev<R>(k)];
                    seen[i]++;
                } else {
                    p[padded<G::PAD>(base + k * M) << G::LOGRC] = v[bitrev<R>(k)];
                }
            }
        }
}

template <int LOGL, int S, bool HALF>
static void emu_stages(std::vector<double2> &sm, const std::vector<double2> &stw, int logr, int rho0, std::vector<double2> &out,
                       std::vector<char> &seen)
{
    if constexpr (S < Geo<LOGL>::NST) {
        emu_stage<LOGL, S, HALF>(sm, stw, logr, rho0, out, seen);
        emu_stages<LOGL, S + 1, HALF>(sm, stw, logr, rho0, out, seen);
    }
}

template <int LOGL>
static double run(int logN, int lo, int K, bool half)
{
    using G = Geo<LOGL>;
    constexpr int STEP = G::STEP, L = G::L;
    const int N = 1 << logN, Nm = N - 1, logr = logN - LOGL, nparts = N >> LOGCH;
    const double pi = 3.14159265358979323846;
    std::vector<double2> root(N);
    for (int m = 0; m < N; m++) root[m] = make_double2(std::cos(2 * pi * m / N), std::sin(2 * pi * m / N));
    std::vector<double2> x(K);
    for (int m = 0; m < K; m++) x[m] = make_double2(std::sin(0.37 * m + 0.1 * lo) + 0.01 * m, std::cos(1.3 * m) - 0.5);
    std::vector<double2> stw(G::STW + 1), sm(G::PLANE, make_double2(NAN, NAN));
    for (int s = 1; s < nstages(LOGL); s++) {
        const int lm = logm(LOGL, s), rs = logN - (lm + lrad(LOGL, s));
        for (int q0 = 0; q0 < (1 << lm); q0++) stw[stw_off(LOGL, s) + q0] = root[(q0 << rs) & Nm];
    }
    std::vector<double2> out(N, make_double2(NAN, NAN));
    std::vector<char> seen(N, 0);
    const int l0 = lo & (L - 1), lh = l0 >> G::LOGBF, ll = l0 & (STEP - 1);
    for (int part = 0; part < nparts; part++) {
        for (int tid = 0; tid < NT; tid++) {
            const int rho = tid & (G::RC - 1), u = tid >> G::LOGRC;
            const int nlow = digit_reverse<LOGL>(u);
            const int m0 = (nlow - lo) & (L - 1);
            const bool flag = nlow >= ll;
            const int a0 = lo + m0;
            const int rg = part * G::RC + rho;  // global residue
            double2 t0 = root[(int)(((long long)a0 * (rg + 1)) & Nm)];
            const double2 g = root[((rg + 1) << G::LOGBF) & Nm];
            const double2 h = root[(N - (((rg + 1) << LOGL) & Nm)) & Nm];
            double2 v[16], tw[16];
            first_stage_twiddles(t0, g, h, lh, flag, tw);
            for (int k = 0; k < 16; k++) {
                const int m = (m0 + k * STEP) & (L - 1);
                const double2 xin = (m < K) ? x[m] : make_double2(0., 0.);
                v[k] = cmul(xin, tw[k]);
            }
            dft16(v);
            double2 *out0 = sm.data() + (padded<G::PAD>(u << 4) << G::LOGRC) + rho;
            for (int k = 0; k < 16; k++) out0[k << G::LOGRC] = v[k];
        }
        if (half)
            emu_stages<LOGL, 1, true>(sm, stw, logr, part << G::LOGRC, out, seen);
        else
            emu_stages<LOGL, 1, false>(sm, stw, logr, part << G::LOGRC, out, seen);
    }
    // direct sum
    const int nout = half ? N / 2 : N;
    double worst = 0., scale = 0.;
    for (int i = 0; i < nout; i += (nout > 512 ? 7 : 1)) {
        std::complex<double> acc(0., 0.);
        for (int m = 0; m < K; m++) {
            const long long e = ((long long)(lo + m) * (i + 1)) % N;
            const double ang = 2 * pi * (double)((e + N) % N) / N;
            acc += std::complex<double>(x[m].x, x[m].y) * std::complex<double>(std::cos(ang), std::sin(ang));
        }
        if (seen[i] != 1) return 1e300;
        worst = std::fmax(worst, std::abs(acc - std::complex<double>(out[i].x, out[i].y)));
        scale = std::fmax(scale, std::abs(acc));
    }
    return worst / scale;
}

template <int LOGL>
static int check(int logN)
{
    int bad = 0;
    const int L = 1 << LOGL;
    const int los[] = {-3, 5, -L / 2 - 1, L + 7, -(1 << logN) / 2 + 1, 17 * L + L / 16 + 1};
    for (int lo : los)
        for (int K : {L, L / 2 + 1, 5})
            for (int half = 0; half < 2; half++) {
                const double e = run<LOGL>(logN, lo, K, half);
                if (!(e < 1e-11)) {
                    std::printf("FAIL logL=%d logN=%d lo=%d K=%d half=%d err=%.3e\n", LOGL, logN, lo, K, half, e);
                    bad++;
                }
            }
    return bad;
}

int main()
{
    int bad = 0;
    for (int logN : {11, 13}) {
        bad += check<5>(logN);
        bad += check<6>(logN);
        bad += check<7>(logN);
        bad += check<8>(logN);
        bad += check<9>(logN);
        bad += check<10>(logN);
        bad += check<11>(logN);
    }
    std::printf(bad ? "v4 emulation: %d failures\n" : "v4 emulation: ok\n", bad);
    return bad ? 1 : 0;
}
