// bf_common.cuh -- shared device/host helpers of the B200 convolution engine.
//
// Everything in the *.cuh headers is written so that it also compiles as plain C++ (g++, with
// BF_HOST_EMULATION defined): tests/host_emul/ runs the very same FFT passes, sample conversion and
// quantiser on the CPU against the oracle before any GPU time is spent.
//
// Data layout on the device ("planar" spectrum), chosen for coalesced float4 streaming:
//   a spectrum of the N-point real transform (N = 2L, M = N/2 bins) is N reals
//     S[k]     = Re X_k            0 <= k < M
//     S[M + k] = Im X_k            1 <= k < M
//     S[M]     = Re X_{N/2}        (Nyquist rides in the imaginary slot of DC)
//   This is a pure permutation of the reference's blocked layout (fftw_convfuns.h:25-42):
//     R[8*(k/4) + k%4] = S[k],  R[8*(k/4) + 4 + k%4] = S[M + k]
//   and of FFTW's half-complex order: hc[k] = S[k] (k < M), hc[M] = S[M], hc[N-k] = S[M+k].
//   The host-visible layout of coefficients stays the reference's blocked one; the permutation is
//   applied on upload / download.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <math.h>

#if defined(__CUDACC__) && !defined(BF_HOST_EMULATION)
#define BF_HD __host__ __device__ __forceinline__
#define BF_D __device__ __forceinline__
// out-of-line on the device: the sample conversion routines are format-generic and large; inlining them at
// every use made the fused FFT kernels 5-13 k instructions and instruction-fetch bound (ncu: 34 % no_inst)
#define BF_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define BF_HD inline
#define BF_D inline
#define BF_HD_NOINLINE inline
#endif

namespace bf {

// ---- exactly rounded arithmetic -------------------------------------------------------------
// The reference computes every product and sum of the hot loops as separate IEEE operations
// (gcc -O2 on x86-64 without FMA: fftw_convfuns.h:534-590, convolver_xmm.c:25-30).  nvcc would
// contract a*b+c into one FMA; these wrappers pin the reference's roundings so that the MAC, mix and
// scale stages are BIT-EXACT against it.  The kernels are HBM-bound (0.5 flop/byte), the extra
// instructions are free.
#if defined(__CUDA_ARCH__) && !defined(BF_HOST_EMULATION)
BF_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
BF_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
BF_D float sub_rn(float a, float b) { return __fsub_rn(a, b); }
BF_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
BF_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
BF_D double sub_rn(double a, double b) { return __dsub_rn(a, b); }
#else
// host build is compiled with -ffp-contract=off
inline float mul_rn(float a, float b) { return a * b; }
inline float add_rn(float a, float b) { return a + b; }
inline float sub_rn(float a, float b) { return a - b; }
inline double mul_rn(double a, double b) { return a * b; }
inline double add_rn(double a, double b) { return a + b; }
inline double sub_rn(double a, double b) { return a - b; }
#endif

// ---- layout permutations ----------------------------------------------------------------------
// planar index -> blocked index (reference layout), for a spectrum of N = 2M reals
BF_HD int planar_to_blocked(int i, int M)
{
    const int k = i < M ? i : i - M;
    return ((k >> 2) << 3) + (k & 3) + (i < M ? 0 : 4);
}
// planar index -> FFTW half-complex index
BF_HD int planar_to_hc(int i, int M)
{
    if (i <= M) {
        return i;           // Re X_0..Re X_{M-1}, and S[M] = Re X_M = hc[M]
    }
    return 2 * M - (i - M); // Im X_k = hc[N - k]
}

struct SampleFormat {   // struct sample_format + buffer_format, dai.h:21-34 (hot fields only)
    int isfloat;
    int swap;
    int bytes;
    int sbytes;
    int sample_spacing;
    int byte_offset;
};

struct Overflow {       // struct bfoverflow, bfmod.h:99-104
    unsigned int n_overflows;
    int32_t intlargest;
    double largest;
    double max;
};

}  // namespace bf
