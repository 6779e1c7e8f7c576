// bf_fft2.cuh -- the size-specialised real FFT core (one N = 2M point real transform per thread block).
//
// Same mathematics and the same definition as bf_fft.cuh (FFTW's unnormalised R2HC / HC2R,
// /root/reference/fftw_convolver.c:98-126), restructured so that everything the compiler can know at
// compile time IS known at compile time: M = 2^LOG2M, the radix sequence, every shared-memory offset
// and every twiddle-table offset are constants.  The generic kernels spent ~85 % of their issue slots on
// index arithmetic and waited on twiddle loads from global memory; here
//   * the first pass takes its 16 inputs straight from global memory (no staging round trip),
//   * passes are radix 16 / radix 8 Stockham passes with 16 points per thread (M/16 threads),
//   * twiddles come from per-pass tables T_p[q-1][k] = e^{-2 pi i q k / (Ns R)} laid out so that a warp
//     reads consecutive entries (host-computed in long double, one rounding each), normally resident in
//     shared memory (bulk-copied once per block),
//   * shared memory holds complex pairs (8-byte accesses) and every access pattern of every pass is
//     bank-conflict free: consecutive butterflies read consecutive elements; writes of a pass with
//     Ns >= 16 are consecutive in k; only the buffer written by the first pass (Ns = 1, stride-16 writes)
//     is XOR-swizzled, element a living at a ^ ((a >> 4) & 15).
//
// Radix sequence for M = 2^m, 10 <= m <= 14: one radix-16 pass from registers, then as many further
// radix-16 passes as needed to make the rest a power of 8, then radix-8 passes:
//   m = 10: 16 8 8      m = 11: 16 16 8      m = 12: 16 16 16      m = 13: 16 8 8 8      m = 14: 16 16 8 8
// and below that m = 6: 16 4, m = 7: 16 8, m = 8: 16 16, m = 9: 16 8 4 (4 .. 32 threads per transform, 64 .. 8
// transforms per block).
//
// All functions are per thread ("tid"), synchronisation is the caller's job, and the file compiles as
// plain C++ (tests/host_emul/emul_fft2.cpp runs the phases of all threads in lock step).
#pragma once

#include "bf_common.cuh"
#include "bf_fft.cuh"

#if defined(__CUDACC__) && !defined(BF_HOST_EMULATION)
#define BF_CE static __host__ __device__ constexpr
#else
#define BF_CE static constexpr
#endif

namespace bf {

template <typename T>
struct alignas(2 * sizeof(T)) cpx {
    T x, y;
};

template <int LOG2M>
struct Fft2 {
    static_assert(LOG2M >= 6 && LOG2M <= 14, "specialised FFT sizes: 64 <= M <= 16384");
    static constexpr int M = 1 << LOG2M;
    static constexpr int NT = M / 16;                   // threads per transform
    // Small transforms share a block: SUBS transforms side by side, each in its own region of shared memory, all
    // walking the same passes in lock step (the block-wide barriers serve them all).  Regions are 8 elements apart
    // from a multiple of the bank period so that two 8-thread transforms in one half-warp do not collide.
    static constexpr int SUBS = NT >= 256 ? 1 : 256 / NT;
    static constexpr int CTA = NT * SUBS;               // threads per block
    static constexpr int SUB_STRIDE = M + (SUBS > 1 ? 8 : 0);
    static constexpr int REST = LOG2M - 4;
    // radix 16 first, then as many more radix-16 passes as make the rest a power of 8; where that is impossible
    // (m = 6, 9) one closing radix-4 pass takes the two odd bits
    static constexpr int N4 = (REST % 3 == 2 && REST < 8) ? 1 : 0;
    static constexpr int REST8 = REST - 2 * N4;
    static constexpr int N16 = (REST8 % 3 == 0) ? 0 : (REST8 % 3 == 1 ? 1 : 2);
    static constexpr int N8 = (REST8 - 4 * N16) / 3;
    static constexpr int NP = 1 + N16 + N8 + N4;        // passes
    BF_CE int radix(int p) { return p <= N16 ? 16 : (p <= N16 + N8 ? 8 : 4); }
    BF_CE int ns(int p)                      // points already combined before pass p
    {
        int n = 1;
        for (int i = 0; i < p; i++) n *= radix(i);
        return n;
    }
    BF_CE int tw_off(int p)                  // first entry of pass p's table (pass 0 has none)
    {
        int o = 0;
        for (int i = 1; i < p; i++) o += (radix(i) - 1) * ns(i);
        return o;
    }
    static constexpr int TW_SPLIT = tw_off(NP);         // W_N^k, k = 0 .. M/2 (real split / merge)
    static constexpr int TW_TOTAL = (TW_SPLIT + M / 2 + 1 + 1) & ~1;    // complex entries, even
    static_assert(ns(NP) == M, "radix sequence does not multiply to M");
};

// element a of the buffer written by pass 0
BF_HD int fft2_swz(int a) { return a ^ ((a >> 4) & 15); }

// host: fill the table of an M-point plan (N = 2M real points).  `out` receives TW_TOTAL complex entries.
template <typename T, int LOG2M>
inline void fft2_fill_table(cpx<T> *out)
{
    typedef Fft2<LOG2M> F;
    const long double pi2 = 2.0L * 3.14159265358979323846264338327950288L;
    for (int p = 1; p < F::NP; p++) {
        const int R = F::radix(p), Ns = F::ns(p);
        for (int q = 1; q < R; q++) {
            for (int k = 0; k < Ns; k++) {
                // exact quadrant values where the angle is a multiple of pi/2
                const long qk = (long)q * k, den = (long)Ns * R;
                cpx<T> w;
                if ((4 * qk) % den == 0) {
                    const int quad = (int)((4 * qk) / den) & 3;
                    w.x = (T)(quad == 0 ? 1 : (quad == 2 ? -1 : 0));
                    w.y = (T)(quad == 1 ? -1 : (quad == 3 ? 1 : 0));
                } else {
                    const long double a = -pi2 * (long double)(qk % den) / (long double)den;
                    w.x = (T)cosl(a);
                    w.y = (T)sinl(a);
                }
                out[F::tw_off(p) + (q - 1) * Ns + k] = w;
            }
        }
    }
    const int N = 2 * F::M;
    for (int k = 0; k <= F::M / 2; k++) {
        cpx<T> w;
        if (k == F::M / 2) {
            w.x = (T)0;     // W_N^{N/4} = -i exactly
            w.y = (T)-1;
        } else {
            const long double a = -pi2 * (long double)k / (long double)N;
            w.x = (T)cosl(a);
            w.y = (T)sinl(a);
        }
        out[F::TW_SPLIT + k] = w;
    }
    for (int i = F::TW_SPLIT + F::M / 2 + 1; i < F::TW_TOTAL; i++) {
        out[i].x = (T)0;
        out[i].y = (T)0;
    }
}

// ---- butterflies on complex registers ------------------------------------------------------------------
template <typename T> BF_HD cpx<T> cadd(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename T> BF_HD cpx<T> csub(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename T> BF_HD cpx<T> cmul(cpx<T> a, cpx<T> w)
{
    cpx<T> r;
    r.x = a.x * w.x - a.y * w.y;
    r.y = a.x * w.y + a.y * w.x;
    return r;
}
// multiply by -i (forward) / +i (inverse)
template <typename T, bool INV> BF_HD cpx<T> cmul_mi(cpx<T> a)
{
    cpx<T> r;
    if (INV) {
        r.x = -a.y;
        r.y = a.x;
    } else {
        r.x = a.y;
        r.y = -a.x;
    }
    return r;
}
// multiply by the constant (c, -s) forward / (c, +s) inverse, i.e. by e^{-+ i theta}
template <typename T, bool INV> BF_HD cpx<T> cmul_const(cpx<T> a, T c, T s)
{
    cpx<T> r;
    if (INV) {
        r.x = a.x * c - a.y * s;
        r.y = a.y * c + a.x * s;
    } else {
        r.x = a.x * c + a.y * s;
        r.y = a.y * c - a.x * s;
    }
    return r;
}

template <typename T, bool INV>
BF_HD void cdft4(cpx<T> &a0, cpx<T> &a1, cpx<T> &a2, cpx<T> &a3)      // natural order in, natural order out
{
    const cpx<T> s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const cpx<T> s1 = cadd(a1, a3), d1 = cmul_mi<T, INV>(csub(a1, a3));
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    a1 = cadd(d0, d1);
    a3 = csub(d0, d1);
}

template <typename T, bool INV>
BF_HD void cdft8(cpx<T> *v)
{
    const T h = (T)0.70710678118654752440;
    cpx<T> a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        a[j] = cadd(v[j], v[j + 4]);
        b[j] = csub(v[j], v[j + 4]);
    }
    b[1] = cmul_const<T, INV>(b[1], h, h);          // W8^1
    b[2] = cmul_mi<T, INV>(b[2]);                   // W8^2
    b[3] = cmul_const<T, INV>(b[3], -h, h);         // W8^3 = (-h, -h) forward
    cdft4<T, INV>(a[0], a[1], a[2], a[3]);
    cdft4<T, INV>(b[0], b[1], b[2], b[3]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        v[2 * j] = a[j];
        v[2 * j + 1] = b[j];
    }
}

// 16 = 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2
template <typename T, bool INV>
BF_HD void cdft16(cpx<T> *v)
{
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    cpx<T> y[4][4];     // y[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        y[n2][0] = v[n2];
        y[n2][1] = v[4 + n2];
        y[n2][2] = v[8 + n2];
        y[n2][3] = v[12 + n2];
        cdft4<T, INV>(y[n2][0], y[n2][1], y[n2][2], y[n2][3]);
    }
    // W16^{n2 k1}
    y[1][1] = cmul_const<T, INV>(y[1][1], c1, s1);      // W16^1
    y[1][2] = cmul_const<T, INV>(y[1][2], h, h);        // W16^2
    y[1][3] = cmul_const<T, INV>(y[1][3], s1, c1);      // W16^3
    y[2][1] = cmul_const<T, INV>(y[2][1], h, h);        // W16^2
    y[2][2] = cmul_mi<T, INV>(y[2][2]);                 // W16^4
    y[2][3] = cmul_const<T, INV>(y[2][3], -h, h);       // W16^6
    y[3][1] = cmul_const<T, INV>(y[3][1], s1, c1);      // W16^3
    y[3][2] = cmul_const<T, INV>(y[3][2], -h, h);       // W16^6
    y[3][3] = cmul_const<T, INV>(y[3][3], -c1, -s1);    // W16^9 = (-c1, +s1) forward
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        cdft4<T, INV>(y[0][k1], y[1][k1], y[2][k1], y[3][k1]);
        v[k1] = y[0][k1];
        v[k1 + 4] = y[1][k1];
        v[k1 + 8] = y[2][k1];
        v[k1 + 12] = y[3][k1];
    }
}

template <typename T, int R, bool INV>
BF_HD void cdft(cpx<T> *v)
{
    if (R == 16) {
        cdft16<T, INV>(v);
    } else if (R == 8) {
        cdft8<T, INV>(v);
    } else {
        cdft4<T, INV>(v[0], v[1], v[2], v[3]);
    }
}

// ---- passes ------------------------------------------------------------------------------------------------
// Pass 0: v[q] = z[tid + q * NT], q = 0..15 (already in registers) -> radix-16 butterfly without twiddles ->
// shared memory, swizzled.
template <typename T, int LOG2M, bool INV>
BF_D void fft2_pass0(cpx<T> *s, int tid, cpx<T> *v)
{
    cdft16<T, INV>(v);
    const int base = 16 * tid, x = tid & 15;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        s[base + (q ^ x)] = v[q];
    }
}

// Pass P >= 1, read half: gathers this thread's 16 / R butterflies, applies the twiddles and the radix-R DFT.
template <typename T, int LOG2M, int P, bool INV>
BF_D void fft2_pass_read(const cpx<T> *s, const cpx<T> *tw, int tid, cpx<T> *v)
{
    typedef Fft2<LOG2M> F;
    constexpr int R = F::radix(P), Ns = F::ns(P), nb = F::M / R, BPT = 16 / R;
    const cpx<T> *tp = tw + F::tw_off(P);
#pragma unroll
    for (int b = 0; b < BPT; b++) {
        const int j = tid + b * F::NT;
        const int k = j & (Ns - 1);
        cpx<T> *r = v + b * R;
        if (P == 1) {
            // the buffer pass 0 wrote: element a at a ^ ((a >> 4) & 15); the low four bits of j are the lane's own
            const int lo = j & 15, hi = j & ~15;
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int a = hi + q * nb;      // multiple of 16
                r[q] = s[a + (lo ^ ((a >> 4) & 15))];
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) {
                r[q] = s[j + q * nb];
            }
        }
#pragma unroll
        for (int q = 1; q < R; q++) {
            cpx<T> w = tp[(q - 1) * Ns + k];
            if (INV) {
                w.y = -w.y;
            }
            r[q] = cmul(r[q], w);
        }
        cdft<T, R, INV>(r);
    }
}

// Pass P >= 1, write half (after a block-wide synchronisation): Stockham autosort placement, linear layout.
template <typename T, int LOG2M, int P>
BF_D void fft2_pass_write(cpx<T> *s, int tid, const cpx<T> *v)
{
    typedef Fft2<LOG2M> F;
    constexpr int R = F::radix(P), Ns = F::ns(P), BPT = 16 / R;
#pragma unroll
    for (int b = 0; b < BPT; b++) {
        const int j = tid + b * F::NT;
        const int k = j & (Ns - 1);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; q++) {
            s[j0 + q * Ns] = v[b * R + q];
        }
    }
}

// Passes 1 .. NP-1 in place; v holds pass 0's input on entry.  With LAST_IN_REGS the final pass stops after its
// butterflies: v[b * 8 + q] then holds element (tid + b * NT) + q * (M / 8) of the result (natural order) and shared
// memory is not written (the inverse transform's epilogue consumes registers).  Otherwise the result is in shared
// memory, natural order, linear, and a final synchronisation has been issued.
template <typename T, int LOG2M, bool INV, bool LAST_IN_REGS, int P, typename Sync>
struct Fft2Passes {
    static BF_D void run(cpx<T> *s, const cpx<T> *tw, int tid, cpx<T> *v, Sync sync)
    {
        typedef Fft2<LOG2M> F;
        sync();         // the previous pass's writes are visible
        fft2_pass_read<T, LOG2M, P, INV>(s, tw, tid, v);
        if (LAST_IN_REGS && P == F::NP - 1) {
            return;
        }
        sync();         // everybody has read
        fft2_pass_write<T, LOG2M, P>(s, tid, v);
        if (P == F::NP - 1) {
            sync();
        }
        Fft2Passes<T, LOG2M, INV, LAST_IN_REGS, (P + 1 < F::NP ? P + 1 : -1), Sync>::run(s, tw, tid, v, sync);
    }
};
template <typename T, int LOG2M, bool INV, bool LAST_IN_REGS, typename Sync>
struct Fft2Passes<T, LOG2M, INV, LAST_IN_REGS, -1, Sync> {
    static BF_D void run(cpx<T> *, const cpx<T> *, int, cpx<T> *, Sync) {}
};

template <typename T, int LOG2M, bool INV, bool LAST_IN_REGS, typename Sync>
BF_D void fft2_complex(cpx<T> *s, const cpx<T> *tw, int tid, cpx<T> *v, Sync sync)
{
    fft2_pass0<T, LOG2M, INV>(s, tid, v);
    Fft2Passes<T, LOG2M, INV, LAST_IN_REGS, 1, Sync>::run(s, tw, tid, v, sync);
}

// ---- real <-> complex glue ------------------------------------------------------------------------------------
// Forward: after fft2_complex (result in shared memory) every thread emits its share of the M/2 + 1 bin pairs.
// emit(k, re, im): bin k of the real transform; the Nyquist value rides as the imaginary part of bin 0.
template <typename T, int LOG2M, typename Emit>
BF_D void fft2_split_emit(const cpx<T> *s, const cpx<T> *tw, int tid, Emit emit)
{
    typedef Fft2<LOG2M> F;
    const cpx<T> *ts = tw + F::TW_SPLIT;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int k = tid + b * F::NT;          // 0 .. M/2 - 1
        if (b == 0 && tid == 0) {
            const cpx<T> z = s[0];
            emit(0, z.x + z.y, z.x - z.y);
            // bin M/2 pairs with itself
            const cpx<T> zh = s[F::M / 2], w = ts[F::M / 2];
            T xkr, xki, xmr, xmi;
            fft_split_pair<T>(zh.x, zh.y, zh.x, zh.y, w.x, w.y, xkr, xki, xmr, xmi);
            emit(F::M / 2, xkr, xki);
        } else {
            const cpx<T> zk = s[k], zm = s[F::M - k], w = ts[k];
            T xkr, xki, xmr, xmi;
            fft_split_pair<T>(zk.x, zk.y, zm.x, zm.y, w.x, w.y, xkr, xki, xmr, xmi);
            emit(k, xkr, xki);
            emit(F::M - k, xmr, xmi);
        }
    }
}

// Inverse: load(i) returns planar element i of the spectrum (S[k] = Re X_k, S[M + k] = Im X_k, S[M] = Re X_M);
// writes the merged complex sequence Z into shared memory (linear).  The caller synchronises, then gathers
// v[q] = s[tid + q * NT] and runs fft2_complex<INV = true>.
template <typename T, int LOG2M, typename Load>
BF_D void fft2_merge_load(cpx<T> *s, const cpx<T> *tw, int tid, Load load)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M;
    const cpx<T> *ts = tw + F::TW_SPLIT;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int k = tid + b * F::NT;
        if (b == 0 && tid == 0) {
            const T x0 = load(0), xm = load(M);
            cpx<T> z;
            z.x = x0 + xm;
            z.y = x0 - xm;
            s[0] = z;
            const T hr = load(M / 2), hi = load(M + M / 2);
            const cpx<T> w = ts[M / 2];
            T zkr, zki, zmr, zmi;
            fft_merge_pair<T>(hr, hi, hr, hi, w.x, w.y, zkr, zki, zmr, zmi);
            z.x = zkr;
            z.y = zki;
            s[M / 2] = z;
        } else {
            const T xkr = load(k), xki = load(M + k), xmr = load(M - k), xmi = load(2 * M - k);
            const cpx<T> w = ts[k];
            cpx<T> zk, zm;
            fft_merge_pair<T>(xkr, xki, xmr, xmi, w.x, w.y, zk.x, zk.y, zm.x, zm.y);
            s[k] = zk;
            s[M - k] = zm;
        }
    }
}

// The same in two halves, so that the loads of the NEXT transform can be in flight while the current one finishes:
// fetch the 32 spectrum values this thread merges (x[4b .. 4b+3] = S[k], S[M+k], S[M-k], S[2M-k], k = tid + b NT;
// thread 0's first group is S[0], S[M], S[M/2], S[M + M/2]) ...
template <typename T, int LOG2M, typename Load>
BF_D void fft2_merge_fetch(int tid, T *x, Load load)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int k = tid + b * F::NT;
        if (b == 0 && tid == 0) {
            x[0] = load(0);
            x[1] = load(M);
            x[2] = load(M / 2);
            x[3] = load(M + M / 2);
        } else {
            x[4 * b] = load(k);
            x[4 * b + 1] = load(M + k);
            x[4 * b + 2] = load(M - k);
            x[4 * b + 3] = load(2 * M - k);
        }
    }
}

// ... and merge them into shared memory.
template <typename T, int LOG2M>
BF_D void fft2_merge_store(cpx<T> *s, const cpx<T> *tw, int tid, const T *x)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M;
    const cpx<T> *ts = tw + F::TW_SPLIT;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int k = tid + b * F::NT;
        if (b == 0 && tid == 0) {
            cpx<T> z;
            z.x = x[0] + x[1];
            z.y = x[0] - x[1];
            s[0] = z;
            const cpx<T> w = ts[M / 2];
            T zkr, zki, zmr, zmi;
            fft_merge_pair<T>(x[2], x[3], x[2], x[3], w.x, w.y, zkr, zki, zmr, zmi);
            z.x = zkr;
            z.y = zki;
            s[M / 2] = z;
        } else {
            const cpx<T> w = ts[k];
            cpx<T> zk, zm;
            fft_merge_pair<T>(x[4 * b], x[4 * b + 1], x[4 * b + 2], x[4 * b + 3], w.x, w.y, zk.x, zk.y, zm.x, zm.y);
            s[k] = zk;
            s[M - k] = zm;
        }
    }
}

}  // namespace bf
