// bf_sample.cuh -- raw PCM <-> real conversion: the device form of raw2real.h / real2raw.h.
//
//   raw_to_real : RAW2REAL_NAME, /root/reference/raw2real.h:7-160.  Integers are NOT scaled
//                 (the 2^-(bits-1) factor is folded into the input mix, bfrun.c:1664-1665).
//   real_to_int : ditherd_real2int_no_dither, /root/reference/dither_funs.h:70-114.  The DOUBLE
//                 instance serves both precisions (fftw_convolver.c:447-449, 470-472).
//   store_sample / float stats : REAL2RAW_NAME no-dither instance, /root/reference/real2raw.h:61-251,
//                 REAL_OVERFLOW_UPDATE real2raw.h:44-59, sample test real2raw.h:24-42.
//
// Overflow accounting is kept as per-thread candidates and reduced per output channel afterwards; the
// reference's sequential "running maximum / counter" over the block equals max / sum over the block.
#pragma once

#include "bf_common.cuh"

namespace bf {

// ---- raw bytes <-> 64-bit little-endian word -------------------------------------------------
// A sample of `bytes` bytes is moved as ONE access of its natural width when its address allows it
// (the common interleaved / planar layouts of 2-, 4- and 8-byte formats do), otherwise byte by byte.
// No arrays: everything stays in registers.
BF_HD_NOINLINE uint64_t load_raw_le(const uint8_t *p, int bytes)
{
    const uintptr_t a = (uintptr_t)p;
    if (bytes == 4 && (a & 3) == 0) {
        return (uint64_t)*reinterpret_cast<const uint32_t *>(p);
    }
    if (bytes == 2 && (a & 1) == 0) {
        return (uint64_t)*reinterpret_cast<const uint16_t *>(p);
    }
    if (bytes == 8 && (a & 7) == 0) {
        return *reinterpret_cast<const uint64_t *>(p);
    }
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < bytes) {
            v |= (uint64_t)p[i] << (8 * i);
        }
    }
    return v;
}

BF_HD_NOINLINE void store_raw_le(uint8_t *p, uint64_t v, int bytes)
{
    const uintptr_t a = (uintptr_t)p;
    if (bytes == 4 && (a & 3) == 0) {
        *reinterpret_cast<uint32_t *>(p) = (uint32_t)v;
        return;
    }
    if (bytes == 2 && (a & 1) == 0) {
        *reinterpret_cast<uint16_t *>(p) = (uint16_t)v;
        return;
    }
    if (bytes == 8 && (a & 7) == 0) {
        *reinterpret_cast<uint64_t *>(p) = v;
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < bytes) {
            p[i] = (uint8_t)(v >> (8 * i));
        }
    }
}

// reverse the low `bytes` bytes of v (the reference's SWAP16/SWAP32/SWAP64 and the 3-byte cases)
BF_HD uint64_t swap_bytes(uint64_t v, int bytes)
{
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < bytes) {
            r |= ((v >> (8 * i)) & 0xffu) << (8 * (bytes - 1 - i));
        }
    }
    return r;
}

// sample bits (as fetched by load_raw_le) -> real, unscaled
template <typename T>
BF_HD_NOINLINE T decode_sample(uint64_t bits, int bytes, int isfloat, int swap)
{
    if (swap && bytes > 1) {
        bits = swap_bytes(bits, bytes);
    }
    if (isfloat) {
        if (bytes == 4) {
            union { uint32_t u; float f; } v;
            v.u = (uint32_t)bits;
            return (T)v.f;
        }
        union { uint64_t u; double f; } v;
        v.u = bits;
        return (T)v.f;
    }
    switch (bytes) {
    case 1:
        return (T)(int8_t)(uint8_t)bits;
    case 2:
        return (T)(int16_t)(uint16_t)bits;
    case 3:
        // three bytes into the top of an int32, arithmetic shift down (raw2real.h:106-142)
        return (T)((int32_t)((uint32_t)bits << 8) >> 8);
    default:
        return (T)(int32_t)(uint32_t)bits;
    }
}

template <typename T>
BF_HD T raw_to_real(const uint8_t *p, int bytes, int isfloat, int swap)
{
    return decode_sample<T>(load_raw_le(p, bytes), bytes, isfloat, swap);
}

// status bits raised by the output stage (the reference abort()s / bf_exit()s instead)
#define BF_STATUS_NONFINITE 1u
#define BF_STATUS_SAFETY 2u

struct QuantStats {
    unsigned int n_overflows;   // to add
    int32_t intlargest;         // candidate maximum (>= 0)
    double largest;             // candidate maximum (>= 0)
    unsigned int status;
};

BF_HD void quant_stats_init(QuantStats &s)
{
    s.n_overflows = 0;
    s.intlargest = 0;
    s.largest = 0.0;
    s.status = 0;
}

// real2raw.h:24-42
template <typename T>
BF_HD void sample_test(T v, double safety_limit, double of_max, QuantStats &s)
{
    if (!isfinite((double)v)) {
        s.status |= BF_STATUS_NONFINITE;
    }
    if (safety_limit != 0.0 && ((double)v < -safety_limit * of_max || (double)v > safety_limit * of_max)) {
        s.status |= BF_STATUS_SAFETY;
    }
}

// dither_funs.h:70-114.  rmin/rmax are (double)(T)imin / (double)(T)imax, i.e. rounded through the
// real type first (real2raw.h:165-168).
template <typename T>
BF_HD int32_t real_to_int(T sample, double rmin, double rmax, int32_t imin, int32_t imax, QuantStats &s)
{
    double y = (double)sample;
    int32_t q;
    y += 0.5;
    if (y < 0) {
        if (y <= rmin) {
            q = imin;
            s.n_overflows++;
            if (-y > s.largest) {
                s.largest = -y;
            }
        } else {
            q = (int32_t)y;
            q--;
            if (-q > s.intlargest) {
                s.intlargest = -q;
            }
        }
    } else {
        if (y > rmax) {
            q = imax;
            s.n_overflows++;
            if (y > s.largest) {
                s.largest = y;
            }
        } else {
            q = (int32_t)y;
            if (q > s.intlargest) {
                s.intlargest = q;
            }
        }
    }
    return q;
}

// float output formats: REAL_OVERFLOW_UPDATE (real2raw.h:44-59) with rmin = (T)-max, rmax = (T)max
// The same quantiser for float samples well inside the clip range, in single precision and without the sum x + 0.5
// (which is NOT exact in float for small x): with fl = floor(x) and fr = x - fl (exact), y = x + 0.5 has
// floor(y) = fl + (fr >= 0.5), y < 0 <=> x < -0.5, and y is an integer <=> fr == 0.5.  The double chain of
// dither_funs.h:70-114 gives y >= 0 -> (int32)y = floor(y); y < 0 -> (int32)y - 1 = floor(y), or y - 1 where y is an
// integer (the reference's off-by-one for negative exact integers).  `cand` is what the branch taken would compare with
// intlargest (q, or -q on the negative side).  Returns false where the full chain must run (near or beyond full scale,
// non-finite).  thr = min(2^22 - 1, imax - 1).
BF_HD bool real_to_int_fast(float x, float thr, int32_t &q, int32_t &cand)
{
    if (!(fabsf(x) <= thr)) {
        return false;
    }
    const float fl = floorf(x);
    const float fr = x - fl;
    q = (int32_t)fl + (fr >= 0.5f ? 1 : 0);
    const bool neg = x < -0.5f;
    if (neg && fr == 0.5f) {
        q--;
    }
    cand = neg ? -q : q;
    return true;
}
BF_HD bool real_to_int_fast(double, float, int32_t &, int32_t &) { return false; }

template <typename T>
BF_HD void float_overflow_update(T v, T rmin, T rmax, QuantStats &s)
{
    if (v < (T)0.0) {
        if (v < rmin) {
            s.n_overflows++;
        }
        if ((double)-v > s.largest) {
            s.largest = (double)-v;
        }
    } else {
        if (v > rmax) {
            s.n_overflows++;
        }
        if ((double)v > s.largest) {
            s.largest = (double)v;
        }
    }
}

// One output sample: test, quantise or copy, account; returns the sample's bytes as a little-endian word
// (already byte-swapped for the _BE formats), ready for store_raw_le.
template <typename T>
BF_HD_NOINLINE uint64_t encode_sample(T v, int bytes, int sbytes, int isfloat, int swap, double safety_limit, double of_max,
                             QuantStats &s)
{
    uint64_t bits;
    sample_test<T>(v, safety_limit, of_max, s);
    if (isfloat) {
        float_overflow_update<T>(v, (T)-of_max, (T)of_max, s);
        if (bytes == 4) {
            union { uint32_t u; float f; } c;
            c.f = (float)v;
            bits = c.u;
        } else {
            union { uint64_t u; double f; } c;
            c.f = (double)v;
            bits = c.u;
        }
    } else {
        const int bits_n = sbytes << 3;
        const int32_t imin = (int32_t)(-((uint64_t)1 << (bits_n - 1)));
        const int32_t imax = (int32_t)(((uint64_t)1 << (bits_n - 1)) - 1);
        const double rmin = (double)(T)imin, rmax = (double)(T)imax;
        const int32_t q = real_to_int<T>(v, rmin, rmax, imin, imax, s);
        // 1: (int8_t)q, 2: (int16_t)q, 3: low three bytes, 4: all of it -- all are the low `bytes` bytes
        bits = (uint64_t)(uint32_t)q;
        if (bytes < 4) {
            bits &= ((uint64_t)1 << (8 * bytes)) - 1;
        }
    }
    if (swap && bytes > 1) {
        bits = swap_bytes(bits, bytes);
    }
    return bits;
}

template <typename T>
BF_HD void real_to_raw(T v, uint8_t *p, int bytes, int sbytes, int isfloat, int swap, double safety_limit,
                       double of_max, QuantStats &s)
{
    store_raw_le(p, encode_sample<T>(v, bytes, sbytes, isfloat, swap, safety_limit, of_max, s), bytes);
}

}  // namespace bf
