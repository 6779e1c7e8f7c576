// CPU emulation of the device FFT (brutefir_b200/csrc/bf_fft.cuh): runs the per-thread phases of
// every "thread" of a block in lock step and checks the real forward / inverse transforms against a
// double-precision direct DFT.  Built and run by tests/test_host_emulation.py (no GPU needed).
#define BF_HOST_EMULATION 1
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../brutefir_b200/csrc/bf_fft.cuh"

using namespace bf;

template <typename T>
static std::vector<T> make_table(int N)
{
    std::vector<T> tw(N);   // N/2 complex
    for (int j = 0; j < N / 2; j++) {
        long double a = -2.0L * M_PIl * (long double)j / (long double)N;
        tw[2 * j] = (T)cosl(a);
        tw[2 * j + 1] = (T)sinl(a);
    }
    return tw;
}

template <typename T, int E, bool INV>
static void emul_cfft(std::vector<T> &sre, std::vector<T> &sim, const T *tw, int M, int nt)
{
    std::vector<FftRegs<T, E>> regs(nt);
    int lg = 0;
    while ((1 << lg) < M) lg++;
    int Ns = 1;
    const int rem = lg % 3;
    if (rem == 1) {
        for (int t = 0; t < nt; t++) fft_pass_read<T, E, 2, INV>(sre.data(), sim.data(), tw, M, Ns, t, nt, regs[t]);
        for (int t = 0; t < nt; t++) fft_pass_write<T, E, 2>(sre.data(), sim.data(), M, Ns, t, nt, regs[t]);
        Ns *= 2;
    } else if (rem == 2) {
        for (int t = 0; t < nt; t++) fft_pass_read<T, E, 4, INV>(sre.data(), sim.data(), tw, M, Ns, t, nt, regs[t]);
        for (int t = 0; t < nt; t++) fft_pass_write<T, E, 4>(sre.data(), sim.data(), M, Ns, t, nt, regs[t]);
        Ns *= 4;
    }
    while (Ns < M) {
        for (int t = 0; t < nt; t++) fft_pass_read<T, E, 8, INV>(sre.data(), sim.data(), tw, M, Ns, t, nt, regs[t]);
        for (int t = 0; t < nt; t++) fft_pass_write<T, E, 8>(sre.data(), sim.data(), M, Ns, t, nt, regs[t]);
        Ns *= 8;
    }
}

template <typename T, int E>
static int check(int N, double tol)
{
    const int M = N / 2;
    int nt = fft_threads(M);
    if (E == 16) nt = M / 16;
    std::vector<T> tw = make_table<T>(N);
    std::vector<double> x(N);
    for (int j = 0; j < N; j++) x[j] = (double)rand() / RAND_MAX - 0.5;
    // forward
    std::vector<T> sre(fft_smem_reals(M) / 2), sim(fft_smem_reals(M) / 2);
    for (int j = 0; j < M; j++) {
        sre[fft_pad(j)] = (T)x[2 * j];
        sim[fft_pad(j)] = (T)x[2 * j + 1];
    }
    emul_cfft<T, E, false>(sre, sim, tw.data(), M, nt);
    std::vector<double> S(N);   // planar spectrum
    S[0] = (double)(sre[0] + sim[0]);
    S[M] = (double)(sre[0] - sim[0]);
    for (int k = 1; k <= M / 2; k++) {
        T xkr, xki, xmr, xmi, wr, wi;
        fft_twiddle<T>(tw.data(), M, k, false, wr, wi);
        fft_split_pair<T>(sre[fft_pad(k)], sim[fft_pad(k)], sre[fft_pad(M - k)], sim[fft_pad(M - k)], wr, wi,
                          xkr, xki, xmr, xmi);
        S[k] = xkr; S[M + k] = xki;
        S[M - k] = xmr; S[M + (M - k)] = xmi;
    }
    double emax = 0, smax = 0;
    const int step = N > 2048 ? 37 : 1;     // sample the bins for big N (direct DFT is O(N^2))
    for (int k = 0; k <= M; k += step) {
        long double re = 0, im = 0;
        for (int j = 0; j < N; j++) {
            long double a = -2.0L * M_PIl * (long double)(((long)j * k) % N) / N;
            re += (T)x[j] * cosl(a);
            im += (T)x[j] * sinl(a);
        }
        double gr = k < M ? S[k] : S[M];
        double gi = (k == 0 || k == M) ? 0.0 : S[M + k];
        emax = fmax(emax, fabs((double)re - gr));
        if (k != 0 && k != M) emax = fmax(emax, fabs((double)im - gi));
        smax = fmax(smax, fabs((double)re));
    }
    // inverse of the exact-ish spectrum we just produced: must give N * x back
    for (int k = 1; k <= M / 2; k++) {
        T zkr, zki, zmr, zmi, wr, wi;
        fft_twiddle<T>(tw.data(), M, k, false, wr, wi);
        fft_merge_pair<T>((T)S[k], (T)S[M + k], (T)S[M - k], (T)S[M + (M - k)], wr, wi, zkr, zki, zmr, zmi);
        sre[fft_pad(k)] = zkr; sim[fft_pad(k)] = zki;
        sre[fft_pad(M - k)] = zmr; sim[fft_pad(M - k)] = zmi;
    }
    sre[0] = (T)(S[0] + S[M]);
    sim[0] = (T)(S[0] - S[M]);
    emul_cfft<T, E, true>(sre, sim, tw.data(), M, nt);
    double rmax = 0;
    for (int j = 0; j < M; j++) {
        rmax = fmax(rmax, fabs((double)sre[fft_pad(j)] / N - (double)(T)x[2 * j]));
        rmax = fmax(rmax, fabs((double)sim[fft_pad(j)] / N - (double)(T)x[2 * j + 1]));
    }
    const double rel = emax / (smax > 0 ? smax : 1.0);
    const int ok = rel < tol && rmax < tol;
    printf("%s N=%6d E=%2d nt=%4d  fwd rel err %.3e  roundtrip err %.3e  %s\n", sizeof(T) == 4 ? "f32" : "f64", N, E,
           nt, rel, rmax, ok ? "ok" : "FAIL");
    return ok ? 0 : 1;
}

int main()
{
    int bad = 0;
    for (int N = 8; N <= 16384; N *= 2) {     // 8 points per thread covers M <= 8192 (1024 threads)
        bad += check<float, 8>(N, 2e-6);
        if (N <= 16384) bad += check<double, 8>(N, 1e-13);
    }
    bad += check<float, 16>(32768, 2e-6);
    bad += check<float, 16>(16384, 2e-6);
    bad += check<double, 16>(16384, 1e-13);
    return bad;
}
