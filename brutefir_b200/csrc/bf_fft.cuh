// bf_fft.cuh -- one real FFT of N = 2M points per thread block, staged in shared memory.
//
// Replaces the FFTW r2r plans of the reference (fftw_convolver.c:98-126: FFTW_R2HC / FFTW_HC2R of
// size n_fft, unnormalised).  FFTW's codelet order is not reproducible, only its definition is:
//     forward  X_k = sum_j x_j e^{-2 pi i jk/N},   inverse  y_j = sum_k X_k e^{+2 pi i jk/N}.
//
// Method: the N-point real transform is an M-point complex transform of z_j = x_2j + i x_2j+1 plus
// an O(N) split (forward) / merge (inverse) pass.  The complex transform is a Stockham autosort FFT,
// radix 8 with one leading radix-2 or radix-4 pass, done IN PLACE in shared memory: every thread
// pulls its butterflies' inputs into registers, the block synchronises, every thread writes its
// outputs.  Shared memory is split re[] / im[] and padded one word per 32 so that the strided
// writes of the first passes spread over the banks.
//
// Twiddles come from a table W[j] = e^{-2 pi i j/N}, j in [0, N/2), computed on the host in long
// double and rounded once; the upper half circle is the negated lower half.
//
// All functions are written per thread ("tid of nt") with the block-wide synchronisation left to the
// caller, so the same code runs under the CPU emulation harness (tests/host_emul).
#pragma once

#include "bf_common.cuh"

namespace bf {

BF_HD int fft_pad(int i) { return i + (i >> 5); }
BF_HD int fft_smem_reals(int M) { return 2 * (M + (M >> 5) + 1); }

// number of threads a block uses for an M-point complex transform: 8 points per thread, capped
BF_HD int fft_threads(int M)
{
    int nt = M / 8;
    if (nt > 1024) nt = 1024;
    if (nt < 32) nt = 32;
    return nt;
}

template <typename T>
BF_HD void fft_twiddle(const T *__restrict__ tw, int halfN, int idx, bool inverse, T &wr, T &wi)
{
    bool neg = false;
    if (idx >= halfN) {
        idx -= halfN;
        neg = true;
    }
    // one vector load for the (cos, -sin) pair
    struct alignas(2 * sizeof(T)) Pair { T c, s; };
    const Pair pr = reinterpret_cast<const Pair *>(tw)[idx];
    T c = pr.c, s = pr.s;
    if (neg) {
        c = -c;
        s = -s;
    }
    wr = c;
    wi = inverse ? -s : s;
}

// ---- butterflies (forward: multiply by -i where marked; inverse: +i) ---------------------------
template <typename T, bool INV>
BF_HD void mul_mi(T &re, T &im)   // (re + i im) * (-i) forward, * (+i) inverse
{
    const T r = re;
    if (INV) {
        re = -im;
        im = r;
    } else {
        re = im;
        im = -r;
    }
}

template <typename T, bool INV>
BF_HD void dft4(T &r0, T &i0, T &r1, T &i1, T &r2, T &i2, T &r3, T &i3)
{
    const T s0r = r0 + r2, s0i = i0 + i2, d0r = r0 - r2, d0i = i0 - i2;
    const T s1r = r1 + r3, s1i = i1 + i3;
    T d1r = r1 - r3, d1i = i1 - i3;
    mul_mi<T, INV>(d1r, d1i);
    r0 = s0r + s1r; i0 = s0i + s1i;
    r2 = s0r - s1r; i2 = s0i - s1i;
    r1 = d0r + d1r; i1 = d0i + d1i;
    r3 = d0r - d1r; i3 = d0i - d1i;
}

template <typename T, bool INV>
BF_HD void dft8(T *r, T *i)   // in: v0..v7, out: X0..X7 in natural order
{
    const T h = (T)0.70710678118654752440;
    T ar[4], ai[4], br[4], bi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ar[j] = r[j] + r[j + 4];
        ai[j] = i[j] + i[j + 4];
        br[j] = r[j] - r[j + 4];
        bi[j] = i[j] - i[j + 4];
    }
    // b1 *= W8^1, b2 *= W8^2, b3 *= W8^3   (forward W8 = e^{-i pi/4}; inverse conjugate)
    {
        T t;
        if (INV) {
            t = br[1]; br[1] = h * (t - bi[1]); bi[1] = h * (t + bi[1]);          // * (1+i)/sqrt2
            t = br[3]; br[3] = h * (-t - bi[3]); bi[3] = h * (t - bi[3]);          // * (-1+i)/sqrt2
        } else {
            t = br[1]; br[1] = h * (t + bi[1]); bi[1] = h * (bi[1] - t);          // * (1-i)/sqrt2
            t = br[3]; br[3] = h * (bi[3] - t); bi[3] = h * (-t - bi[3]);          // * (-1-i)/sqrt2
        }
        mul_mi<T, INV>(br[2], bi[2]);
    }
    dft4<T, INV>(ar[0], ai[0], ar[1], ai[1], ar[2], ai[2], ar[3], ai[3]);
    dft4<T, INV>(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        r[2 * j] = ar[j];
        i[2 * j] = ai[j];
        r[2 * j + 1] = br[j];
        i[2 * j + 1] = bi[j];
    }
}

// Registers of one thread across the read->sync->write of a pass: E complex values.
template <typename T, int E>
struct FftRegs {
    T re[E], im[E];
};

// Read phase of one Stockham pass of radix R with Ns points already combined.
template <typename T, int E, int R, bool INV>
BF_HD void fft_pass_read(const T *sre, const T *sim, const T *__restrict__ tw, int M, int Ns, int tid,
                         int nt, FftRegs<T, E> &g)
{
    const int nb = M / R;               // butterflies in this pass
    const int tstep = (2 * M) / (Ns * R);   // table index step of the W_N table for this pass
#pragma unroll
    for (int b = 0; b < E / R; b++) {
        const int j = tid + b * nt;
        if (j < nb) {
            const int k = j & (Ns - 1);
            T *r = &g.re[b * R], *i = &g.im[b * R];
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int a = fft_pad(j + q * nb);
                r[q] = sre[a];
                i[q] = sim[a];
            }
            if (Ns > 1) {
                // Twiddles w^q = e^{-+2 pi i q k / (Ns R)}, q = 1..R-1.  -DBF_TWIDDLE_DERIVED fetches only w^1, w^2,
                // w^4 and forms the others as products of two table entries: ~7 us faster per FFT stage at the
                // headline shape, but the extra rounding measurably widens the distance to the reference
                // (tools/diag_accuracy.py), so it is off.
                T wr[R], wi[R];
                const int i1 = k * tstep;
#ifndef BF_TWIDDLE_DERIVED
                // every power straight from the table: one rounding each (the accurate default)
#pragma unroll
                for (int q = 1; q < R; q++) {
                    fft_twiddle<T>(tw, M, q * i1, INV, wr[q], wi[q]);
                }
#else
                wr[1] = tw[2 * i1];
                wi[1] = INV ? -tw[2 * i1 + 1] : tw[2 * i1 + 1];
                if (R >= 4) {
                    wr[2] = tw[4 * i1];
                    wi[2] = INV ? -tw[4 * i1 + 1] : tw[4 * i1 + 1];
                    wr[3] = wr[1] * wr[2] - wi[1] * wi[2];
                    wi[3] = wr[1] * wi[2] + wi[1] * wr[2];
                }
                if (R == 8) {
                    wr[4] = tw[8 * i1];
                    wi[4] = INV ? -tw[8 * i1 + 1] : tw[8 * i1 + 1];
                    wr[5] = wr[1] * wr[4] - wi[1] * wi[4];
                    wi[5] = wr[1] * wi[4] + wi[1] * wr[4];
                    wr[6] = wr[2] * wr[4] - wi[2] * wi[4];
                    wi[6] = wr[2] * wi[4] + wi[2] * wr[4];
                    wr[7] = wr[3] * wr[4] - wi[3] * wi[4];
                    wi[7] = wr[3] * wi[4] + wi[3] * wr[4];
                }
#endif
#pragma unroll
                for (int q = 1; q < R; q++) {
                    const T xr = r[q], xi = i[q];
                    r[q] = xr * wr[q] - xi * wi[q];
                    i[q] = xr * wi[q] + xi * wr[q];
                }
            }
            if (R == 8) {
                dft8<T, INV>(r, i);
            } else if (R == 4) {
                dft4<T, INV>(r[0], i[0], r[1], i[1], r[2], i[2], r[3], i[3]);
            } else {
                const T xr = r[0], xi = i[0];
                r[0] = xr + r[1]; i[0] = xi + i[1];
                r[1] = xr - r[1]; i[1] = xi - i[1];
            }
        }
    }
}

template <typename T, int E, int R>
BF_HD void fft_pass_write(T *sre, T *sim, int M, int Ns, int tid, int nt, const FftRegs<T, E> &g)
{
    const int nb = M / R;
#pragma unroll
    for (int b = 0; b < E / R; b++) {
        const int j = tid + b * nt;
        if (j < nb) {
            const int k = j & (Ns - 1);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int a = fft_pad(j0 + q * Ns);
                sre[a] = g.re[b * R + q];
                sim[a] = g.im[b * R + q];
            }
        }
    }
}

// The block-wide synchronisation primitive is a template parameter so that the emulation harness can
// run the phases of all "threads" in lock step.
#if defined(__CUDACC__) && !defined(BF_HOST_EMULATION)
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
#endif

// Complex M-point FFT in shared memory (padded SoA), in place.  M = 2^m, 4 <= M; E = points per
// thread (8 or 16); nt threads with nt * E >= M.  Must be called by all nt threads; `sync` is
// called between phases.  On return the data is in natural order and a final sync has been issued.
template <typename T, int E, bool INV, typename Sync>
BF_D void fft_complex_inplace(T *sre, T *sim, const T *__restrict__ tw, int M, int tid, int nt, Sync sync)
{
    FftRegs<T, E> g;
    int lg = 0;
    while ((1 << lg) < M) {
        lg++;
    }
    int Ns = 1;
    const int rem = lg % 3;
    if (rem == 1) {
        fft_pass_read<T, E, 2, INV>(sre, sim, tw, M, Ns, tid, nt, g);
        sync();
        fft_pass_write<T, E, 2>(sre, sim, M, Ns, tid, nt, g);
        sync();
        Ns *= 2;
    } else if (rem == 2) {
        fft_pass_read<T, E, 4, INV>(sre, sim, tw, M, Ns, tid, nt, g);
        sync();
        fft_pass_write<T, E, 4>(sre, sim, M, Ns, tid, nt, g);
        sync();
        Ns *= 4;
    }
    while (Ns < M) {
        fft_pass_read<T, E, 8, INV>(sre, sim, tw, M, Ns, tid, nt, g);
        sync();
        fft_pass_write<T, E, 8>(sre, sim, M, Ns, tid, nt, g);
        sync();
        Ns *= 8;
    }
}

// ---- real <-> complex glue ----------------------------------------------------------------------
// Forward split: from Z (M-point FFT of the packed sequence) produce X_k and X_{M-k}:
//   E = (Z_k + conj Z_{M-k})/2, O = (Z_k - conj Z_{M-k})/(2i), X_k = E + W_N^k O,
//   X_{M-k} = conj(E) - conj(W_N^k O);  X_0 = Re Z_0 + Im Z_0, X_M = Re Z_0 - Im Z_0.
template <typename T>
BF_HD void fft_split_pair(T zkr, T zki, T zmr, T zmi, T wr, T wi, T &xkr, T &xki, T &xmr, T &xmi)
{
    const T er = (T)0.5 * (zkr + zmr), ei = (T)0.5 * (zki - zmi);
    const T orr = (T)0.5 * (zki + zmi), oi = (T)-0.5 * (zkr - zmr);
    const T tr = orr * wr - oi * wi, ti = orr * wi + oi * wr;
    xkr = er + tr;
    xki = ei + ti;
    xmr = er - tr;
    xmi = ti - ei;
}

// Inverse merge: Z_k = (X_k + conj X_{M-k}) + i W_N^{-k} (X_k - conj X_{M-k}), and Z_{M-k}.
// (wr, wi) is the FORWARD root W_N^k.
template <typename T>
BF_HD void fft_merge_pair(T xkr, T xki, T xmr, T xmi, T wr, T wi, T &zkr, T &zki, T &zmr, T &zmi)
{
    const T sr = xkr + xmr, si = xki - xmi;
    const T dr = xkr - xmr, di = xki + xmi;
    const T pr = dr * wr + di * wi, pi_ = di * wr - dr * wi;   // conj(w) * d
    zkr = sr - pi_;
    zki = si + pr;
    zmr = sr + pi_;
    zmi = pr - si;
}

}  // namespace bf
