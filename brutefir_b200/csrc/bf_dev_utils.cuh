// bf_dev_utils.cuh -- device-only helpers shared by the kernel translation units (bf_kernels.cu,
// bf_fft2_kernels.cu, bf_mac_tma.cu): dynamic shared memory base, the reference's complex product, the
// quantiser-statistics reduction, the crossfade ramp, the output mix, and the mbarrier / bulk-copy PTX.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bf_kernels.h"
#include "bf_sample.cuh"

namespace bf {

template <typename T>
__device__ __forceinline__ T *smem_re()
{
    extern __shared__ __align__(16) unsigned char bf_smem_raw[];
    return reinterpret_cast<T *>(bf_smem_raw);
}

// the reference's blocked complex product for one bin: (re, im) = b (*) c with separate roundings
// (fftw_convfuns.h:548-556 / convolver_xmm.c:25-30)
template <typename T>
__device__ __forceinline__ void cprod(T br, T bi, T cr, T ci, T &re, T &im)
{
    re = sub_rn(mul_rn(br, cr), mul_rn(bi, ci));
    im = add_rn(mul_rn(br, ci), mul_rn(bi, cr));
}

// block-wide reduction of the quantiser statistics into overflow[o] / status
__device__ __forceinline__ void reduce_stats(QuantStats st, Overflow *of, unsigned int *status, void *smem, int tid,
                                             int nt)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        st.n_overflows += __shfl_down_sync(full, st.n_overflows, d);
        st.intlargest = max(st.intlargest, __shfl_down_sync(full, st.intlargest, d));
        st.largest = fmax(st.largest, __shfl_down_sync(full, st.largest, d));
        st.status |= __shfl_down_sync(full, st.status, d);
    }
    QuantStats *w = reinterpret_cast<QuantStats *>(smem);
    __syncthreads();        // shared memory is about to be reused
    if ((tid & 31) == 0) {
        w[tid >> 5] = st;
    }
    __syncthreads();
    if (tid == 0) {
        const int nw = (nt + 31) >> 5;
        for (int i = 1; i < nw; i++) {
            st.n_overflows += w[i].n_overflows;
            st.intlargest = max(st.intlargest, w[i].intlargest);
            st.largest = fmax(st.largest, w[i].largest);
            st.status |= w[i].status;
        }
        // atomics: the blocks of one batch update the same output's counters concurrently (sum / max
        // commute, so the result equals the reference's sequential running count / maximum)
        if (st.n_overflows != 0) {
            atomicAdd(&of->n_overflows, st.n_overflows);
        }
        atomicMax(&of->intlargest, st.intlargest);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(&of->largest),
                  (unsigned long long)__double_as_longlong(st.largest));
        if (st.status != 0) {
            atomicOr(status, st.status);
        }
    }
}

// powersave: is the frame [previous block | block blk] of input c silent?  (test_silent, bfrun.c:722-772: exact zeros,
// or a peak below the analog level)
__device__ __forceinline__ bool frame_silent(const ForwardArgs &a, int c, int blk)
{
    if (a.powersave == 0) {
        return false;
    }
    const unsigned int cur = a.amax_cur[(size_t)blk * a.n_in + c];
    const unsigned int prv = blk == 0 ? a.amax_prev[c] : a.amax_cur[(size_t)(blk - 1) * a.n_in + c];
    const float m = fmaxf(__uint_as_float(cur), __uint_as_float(prv));
    return a.powersave == 1 ? (m == 0.f) : (m < a.ps_thr[c]);
}

template <typename T> struct Vec2;
template <> struct Vec2<float> {
    typedef float2 type;
    static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <> struct Vec2<double> {
    typedef double2 type;
    static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
};

template <typename T>
__device__ __forceinline__ T xfade(T old, T nw, int n, int L);
template <>
__device__ __forceinline__ float xfade<float>(float old, float nw, int n, int L)
{
    // fftw_convolver.c:349-355, literally: f and f*n in float, the old term in double
    const float f = (float)(1.0 / (double)(float)(L - 1));
    const float fn = __fmul_rn(f, (float)n);
    const double a = __dmul_rn((double)old, __dsub_rn(1.0, (double)fn));
    const float b = __fmul_rn(__fmul_rn(nw, f), (float)n);
    return (float)__dadd_rn(a, (double)b);
}
template <>
__device__ __forceinline__ double xfade<double>(double old, double nw, int n, int L)
{
    // the float branch's formula in double (the reference's own double branch is broken)
    const double d = 1.0 / (double)(L - 1);
    const double a = __dmul_rn(old, __dsub_rn(1.0, __dmul_rn(d, (double)n)));
    const double b = __dmul_rn(__dmul_rn(nw, d), (double)n);
    return __dadd_rn(a, b);
}

template <typename T>
__device__ __forceinline__ T mix_terms(const T *__restrict__ Y, const MixTerm *__restrict__ terms, int first, int n,
                                       int n_slots, int split, int N, int i)
{
    T acc = (T)0;
    for (int j = 0; j < n; j++) {
        const MixTerm tm = terms[first + j];
        T y = Y[(size_t)tm.index * N + i];
        for (int z = 1; z < split; z++) {
            y = add_rn(y, Y[((size_t)z * n_slots + tm.index) * N + i]);
        }
        const T v = mul_rn(y, (T)tm.scale);
        acc = j == 0 ? v : add_rn(acc, v);
    }
    return acc;
}

// ---- mbarrier + 1-D bulk copy (TMA) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

}  // namespace bf
