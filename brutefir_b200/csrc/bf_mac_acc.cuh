// bf_mac_acc.cuh -- the exact complex multiply-accumulate of the batched MAC kernels (bf_mac_batch.cu,
// bf_mac_tile.cu): vector types, cp.async helpers, and the accumulator classes that keep every rounding of
// convolver_xmm.c:25-30 / fftw_convfuns.h:548-556 while using the packed FP32 pair instructions of sm_100.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bf_kernels.h"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"

namespace bf {

template <typename T, int W> struct VecB;
template <> struct VecB<float, 4> { typedef float4 type; };
template <> struct VecB<float, 2> { typedef float2 type; };
template <> struct VecB<float, 1> { typedef float type; };
template <> struct VecB<double, 2> { typedef double2 type; };
template <> struct VecB<double, 1> { typedef double type; };

template <typename T, int W>
struct __align__(sizeof(T) * W) LanesB {
    T v[W];
};

// cp.async with a run-time source size: src_bytes = BYTES copies, src_bytes = 0 writes zeros and does not touch
// global memory (the "zfill" form) -- predication without a branch.
template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src, unsigned int src_bytes)
{
    if (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                     : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES),
                     "r"(src_bytes)
                     : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory");
}

template <typename V>
__device__ __forceinline__ V ldg_once(const V *p)
{
    return __ldg(p);
}

// ---- one complex multiply-accumulate ------------------------------------------------------------------------------
// acc += b (*) c with the reference's roundings: four separately rounded products, re = p1 - p2, im = p3 + p4, then
// the two accumulations (convolver_xmm.c:25-30 / fftw_convfuns.h:548-556).  p1 - p2 is computed as p1 + (-bi) ci:
// negating a factor negates the rounded product exactly.  The kernel is FP32-issue bound at B = 8, so the instruction
// count per complex MAC is what matters; sm_100 has packed FP32 pairs (FADD2, FFMA2) and two ways to use them exactly:
//
//  * PairAcc (any type and width): accumulators are (re, im) pairs; the four sums of a MAC are TWO packed adds, the
//    four products stay scalar -- 6 instructions per MAC (8 unpacked).
//  * BinPairAcc (float, even W): accumulators pair two neighbouring BINS; the products are packed too -- 4
//    instructions per MAC.  A packed multiply cannot be written as such: ptxas contracts mul.rn.f32x2 (and an FFMA2
//    with a literal -0.0 addend) feeding a packed add into one FFMA2, which would drop a rounding.  So the product is
//    fma.rn.f32x2(x, y, nz) with nz = (-0.0, -0.0) arriving as a KERNEL PARAMETER: the compiler cannot fold what it
//    does not know, and rn(x*y + -0.0) is rn(x*y) bit for bit for every x*y (including +-0: +0 + -0 = +0,
//    -0 + -0 = -0 in round-to-nearest; NaN and infinities pass through as in a multiply).
template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<double> { typedef double2 type; };

__device__ __forceinline__ float2 add_pair(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ double2 add_pair(double2 a, double2 b)
{
    return make_double2(add_rn(a.x, b.x), add_rn(a.y, b.y));
}
__device__ __forceinline__ float2 make_pair(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ double2 make_pair(double x, double y) { return make_double2(x, y); }
// rn(a * b) per half, as FFMA2 with the run-time (-0.0, -0.0) addend
__device__ __forceinline__ float2 mul_pair_exact(float2 a, float2 b, unsigned long long nz)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(nz));
    return *reinterpret_cast<float2 *>(&rd);
}

template <typename T>
__device__ __forceinline__ typename Pair<T>::type cprod_pair(T br, T bi, T cr, T ci)
{
    return add_pair(make_pair(mul_rn(br, cr), mul_rn(br, ci)), make_pair(mul_rn(-bi, ci), mul_rn(bi, cr)));
}

template <typename T, int W>
struct PairAcc {
    typedef typename VecB<T, W>::type V;
    typedef LanesB<T, W> L;
    typename Pair<T>::type a[W];        // (re, im) per bin
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int l = 0; l < W; l++) {
            a[l] = make_pair((T)0, (T)0);
        }
    }
    template <bool ADD>
    __device__ __forceinline__ void step(const V &wr, const V &wi, const V &hr, const V &hi, unsigned long long)
    {
        const L br = *reinterpret_cast<const L *>(&wr), bi = *reinterpret_cast<const L *>(&wi);
        const L cr = *reinterpret_cast<const L *>(&hr), ci = *reinterpret_cast<const L *>(&hi);
#pragma unroll
        for (int l = 0; l < W; l++) {
            const typename Pair<T>::type p = cprod_pair<T>(br.v[l], bi.v[l], cr.v[l], ci.v[l]);
            a[l] = ADD ? add_pair(a[l], p) : p;
        }
    }
    __device__ __forceinline__ void set(int l, T re, T im) { a[l] = make_pair(re, im); }
    __device__ __forceinline__ void get(V &re, V &im) const
    {
        L ore, oim;
#pragma unroll
        for (int l = 0; l < W; l++) {
            ore.v[l] = a[l].x;
            oim.v[l] = a[l].y;
        }
        re = *reinterpret_cast<V *>(&ore);
        im = *reinterpret_cast<V *>(&oim);
    }
};

template <int W>
struct BinPairAcc {
    static_assert(W % 2 == 0, "pairs of neighbouring bins");
    typedef typename VecB<float, W>::type V;
    struct __align__(sizeof(float) * W) H2 {
        float2 h[W / 2];
    };
    float2 re[W / 2], im[W / 2];        // planar, two bins per register pair
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int h = 0; h < W / 2; h++) {
            re[h] = make_float2(0.0f, 0.0f);
            im[h] = make_float2(0.0f, 0.0f);
        }
    }
    template <bool ADD>
    __device__ __forceinline__ void step(const V &wr, const V &wi, const V &hr, const V &hi, unsigned long long nz)
    {
        const H2 br = *reinterpret_cast<const H2 *>(&wr), bi = *reinterpret_cast<const H2 *>(&wi);
        const H2 cr = *reinterpret_cast<const H2 *>(&hr), ci = *reinterpret_cast<const H2 *>(&hi);
#pragma unroll
        for (int h = 0; h < W / 2; h++) {
            const float2 nbi = make_float2(-bi.h[h].x, -bi.h[h].y);     // folds into the operand's negate modifier
            const float2 pr = add_pair(mul_pair_exact(br.h[h], cr.h[h], nz), mul_pair_exact(nbi, ci.h[h], nz));
            const float2 pi = add_pair(mul_pair_exact(br.h[h], ci.h[h], nz), mul_pair_exact(bi.h[h], cr.h[h], nz));
            re[h] = ADD ? add_pair(re[h], pr) : pr;
            im[h] = ADD ? add_pair(im[h], pi) : pi;
        }
    }
    __device__ __forceinline__ void set(int l, float r, float i)     // l is a compile-time constant at every call site
    {
        if (l & 1) {
            re[l / 2].y = r;
            im[l / 2].y = i;
        } else {
            re[l / 2].x = r;
            im[l / 2].x = i;
        }
    }
    __device__ __forceinline__ void get(V &r, V &i) const
    {
        H2 ore, oim;
#pragma unroll
        for (int h = 0; h < W / 2; h++) {
            ore.h[h] = re[h];
            oim.h[h] = im[h];
        }
        r = *reinterpret_cast<V *>(&ore);
        i = *reinterpret_cast<V *>(&oim);
    }
};

template <typename T, int W> struct AccSel { typedef PairAcc<T, W> type; };
#ifndef BF_MAC_NO_BINPAIR
template <> struct AccSel<float, 2> { typedef BinPairAcc<2> type; };
template <> struct AccSel<float, 4> { typedef BinPairAcc<4> type; };
#endif

struct DcTrue { static constexpr bool value = true; };
struct DcFalse { static constexpr bool value = false; };

}  // namespace bf
