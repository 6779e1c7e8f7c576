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

template <typename T>
BF_HD T raw_to_real(const uint8_t *p, int bytes, int isfloat, int swap)
{
    uint8_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < bytes) {
            t[i] = swap ? p[bytes - 1 - i] : p[i];
        }
    }
    if (isfloat) {
        if (bytes == 4) {
            union { uint32_t u; float f; } v;
            v.u = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            return (T)v.f;
        }
        union { uint64_t u; double f; } v;
        v.u = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v.u |= (uint64_t)t[i] << (8 * i);
        }
        return (T)v.f;
    }
    switch (bytes) {
    case 1:
        return (T)(int8_t)t[0];
    case 2:
        return (T)(int16_t)((uint16_t)t[0] | ((uint16_t)t[1] << 8));
    case 3: {
        // three bytes into the top of an int32, arithmetic shift down (raw2real.h:106-142)
        const uint32_t u = ((uint32_t)t[0] << 8) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 24);
        return (T)((int32_t)u >> 8);
    }
    default:
        return (T)(int32_t)((uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) |
                            ((uint32_t)t[3] << 24));
    }
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

BF_HD void store_bytes(uint8_t *p, const uint8_t *t, int bytes, int swap)
{
    for (int i = 0; i < bytes; i++) {
        p[i] = swap ? t[bytes - 1 - i] : t[i];
    }
}

// One output sample: test, quantise or copy, account, pack.  `p` points at the sample's first byte.
template <typename T>
BF_HD void real_to_raw(T v, uint8_t *p, int bytes, int sbytes, int isfloat, int swap,
                       double safety_limit, double of_max, QuantStats &s)
{
    uint8_t t[8];
    sample_test<T>(v, safety_limit, of_max, s);
    if (isfloat) {
        float_overflow_update<T>(v, (T)-of_max, (T)of_max, s);
        if (bytes == 4) {
            union { uint32_t u; float f; } c;
            c.f = (float)v;
            t[0] = (uint8_t)c.u; t[1] = (uint8_t)(c.u >> 8); t[2] = (uint8_t)(c.u >> 16); t[3] = (uint8_t)(c.u >> 24);
        } else {
            union { uint64_t u; double f; } c;
            c.f = (double)v;
            for (int i = 0; i < 8; i++) {
                t[i] = (uint8_t)(c.u >> (8 * i));
            }
        }
    } else {
        const int bits = sbytes << 3;
        const int32_t imin = (int32_t)(-((uint64_t)1 << (bits - 1)));
        const int32_t imax = (int32_t)(((uint64_t)1 << (bits - 1)) - 1);
        const double rmin = (double)(T)imin, rmax = (double)(T)imax;
        const int32_t q = real_to_int<T>(v, rmin, rmax, imin, imax, s);
        // 1: (int8_t)q, 2: (int16_t)q, 3: low three bytes, 4: all of it -- all are the low `bytes` bytes
        const uint32_t u = (uint32_t)q;
        t[0] = (uint8_t)u; t[1] = (uint8_t)(u >> 8); t[2] = (uint8_t)(u >> 16); t[3] = (uint8_t)(u >> 24);
    }
    store_bytes(p, t, bytes, swap);
}

}  // namespace bf
