// bf_fft2_kernels.cu -- the forward and inverse stages on the size-specialised FFT core (bf_fft2.cuh).
//
//   k_unpack     raw2real: raw PCM block -> planar reals (K1; /root/reference/raw2real.h, fftw_convolver.c:170-194)
//   k_forward2   frame assembly + R2HC + input scale into the delay line
//                (K2, K3; bfrun.c:1494-1560, 1671; fftw_convolver.c:180-214)
//   k_inverse2   output mix + HC2R + overlap-save discard + crossfade
//                (K6..K8; bfrun.c:1847-1936; fftw_convolver.c:330-368, 391-409)
//   k_pack       real2raw: planar reals -> raw PCM block, overflow accounting (K9; real2raw.h, fftw_convolver.c:482-518)
//
// Same results contract as k_forward / k_inverse in bf_kernels.cu (which stay as the path for float_bits 64 and
// for partitions shorter than 1024 samples); what changes is the cost:
//   * persistent blocks: grid = min(transforms, resident blocks), each block walks its transforms, the twiddle
//     table is bulk-copied (cp.async.bulk + mbarrier) into shared memory once per block;
//   * forward: the first radix-16 pass reads its samples straight from the raw block / the previous-block buffer;
//   * inverse: the last pass leaves its results in registers -- only the L valid samples of the overlap-save
//     frame are produced -- and the crossfade / quantise / pack epilogue consumes them there;
//   * sample conversion lives in two transposing kernels (k_unpack before the forward stage, k_pack after the
//     inverse one), so the transforms read and write planar reals with fully coalesced accesses; the next
//     transform's operands are fetched while the current one's results are stored.
// Roofline: per transform N*4 bytes of spectrum + L*bytes of samples; ~2.5 N log2 N flop.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "bf_kernels.h"
#include "bf_fft2.cuh"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"

namespace bf {

struct BlockSync2 {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// shared memory: [M complex data][TW_TOTAL complex twiddles (TWS only)][32 QuantStats][mbarrier]
template <typename T, int LOG2M, bool TWS>
struct Smem2 {
    typedef Fft2<LOG2M> F;
    static constexpr size_t data_bytes = (size_t)F::SUBS * F::SUB_STRIDE * sizeof(cpx<T>);
    static constexpr size_t tw_bytes = TWS ? (size_t)F::TW_TOTAL * sizeof(cpx<T>) : 0;
    static constexpr size_t stats_off = data_bytes + tw_bytes;
    static constexpr size_t bar_off = stats_off + 32 * sizeof(QuantStats);
    static constexpr size_t total = bar_off + 16;
};

template <typename T, int LOG2M, bool TWS>
__device__ __forceinline__ const cpx<T> *stage_twiddles(unsigned char *smem, const cpx<T> *tw_global, int tid)
{
    typedef Smem2<T, LOG2M, TWS> S;
    if (!TWS) {
        return tw_global;
    }
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + S::bar_off);
    unsigned char *dst = smem + S::data_bytes;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)S::tw_bytes);
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
        constexpr uint32_t CH = 32768;
        for (uint32_t off = 0; off < (uint32_t)S::tw_bytes; off += CH) {
            const uint32_t n = (uint32_t)S::tw_bytes - off < CH ? (uint32_t)S::tw_bytes - off : CH;
            bulk_g2s(dst + off, reinterpret_cast<const unsigned char *>(tw_global) + off, n, bar, policy);
        }
    }
    __syncthreads();        // the barrier is initialised before anybody polls it
    return reinterpret_cast<const cpx<T> *>(dst);
}

template <typename T, int LOG2M, bool TWS>
__device__ __forceinline__ void wait_twiddles(unsigned char *smem)
{
    if (TWS) {
        mbar_wait(reinterpret_cast<uint64_t *>(smem + Smem2<T, LOG2M, TWS>::bar_off), 0);
    }
}

// ======================================================================================================
// k_unpack / k_pack -- raw PCM <-> planar reals, transposed through shared memory
//
// The dai block layouts interleave the channels (dai.c:537-576: byte_offset = channel * bytes, sample_spacing =
// channels), so the samples of ONE channel are a 4-byte access every 256 bytes at the headline shape: a transform
// reading them directly touches 8192 sectors for 32 KB of payload and the LSU serialises every warp access into 32
// requests (ncu: the forward stage spent ~half its time there).  These two kernels do the layout change at full
// coalescing instead: a warp moves a 32 channel x 8 sample tile, lanes along the channels on the raw side and
// along time on the planar side.  They also carry the whole sample conversion (raw2real.h / real2raw.h), so the
// transforms themselves are format-free.
// ======================================================================================================

// A warp moves a tile of 32 channels x RW samples; RW = 8 keeps every planar-side access a full 32-byte sector
// (8 consecutive floats of one channel) while giving four times more warps than a square tile -- the quantiser is a
// dependent double-precision chain per sample, so it is parallelism, not bytes, that the small launches lack.
#ifndef BF_PACK_RW
#define BF_PACK_RW 8
#endif
constexpr int RW = BF_PACK_RW;  // samples per tile row group
constexpr int WPB = 8;          // warps per block
constexpr int TS = RW + 1;      // padded row stride of the shared tile

__device__ __forceinline__ float peak_as_float(float v) { return fabsf(v); }
__device__ __forceinline__ float peak_as_float(double v) { return __double2float_ru(fabs(v)); }

template <typename T>
__global__ void __launch_bounds__(32 * WPB) k_unpack(UnpackArgs a)
{
    __shared__ T tile[WPB][32 * TS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * WPB + w) * RW, c0 = blockIdx.y * 32, blk = blockIdx.z;
    if (n0 >= a.L) {
        return;
    }
    const uint8_t *raw = a.raw_in + (size_t)blk * a.in_stride;
    if (a.fast_fmt == 3) {
        // packed 24-bit little-endian, all channels interleaved (massive_config's own format): a tile row is 3 * nc
        // contiguous bytes (nc = channels of this tile, a multiple of 4); 3 nc / 4 lanes fetch the words, every lane picks
        // its three bytes out of two of them (raw2real.h:106-142: into the top of an int32, arithmetic shift down).
        // All 32 lanes take part in the shuffles, whatever nc is.
        const int nc = min(32, a.n_in - c0);
        const size_t stride = (size_t)a.n_in * 3;
        const int w0 = (3 * lane) >> 2, sh = ((3 * lane) & 3) * 8;
        const uint8_t *rowp = raw + (size_t)n0 * stride + (size_t)c0 * 3;
        T *row = &tile[w][lane * TS];
#pragma unroll
        for (int r = 0; r < RW; r++) {
            const uint32_t wv = 4 * lane < 3 * nc ? reinterpret_cast<const uint32_t *>(rowp + r * stride)[lane] : 0u;
            const uint32_t lo = __shfl_sync(0xffffffffu, wv, w0), hi = __shfl_sync(0xffffffffu, wv, (w0 + 1) & 31);
            const uint32_t v3 = __funnelshift_r(lo, hi, sh);
            row[r] = (T)((int32_t)(v3 << 8) >> 8);
        }
    }
    {
        const int c = c0 + lane;
        if (c < a.n_in) {
            const SampleFormat f = a.fmt[c];
            const size_t stride = (size_t)f.sample_spacing * f.bytes;
            const uint8_t *p = raw + f.byte_offset + (size_t)n0 * stride;
            T *row = &tile[w][lane * TS];
            if (a.fast_fmt == 1) {
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    row[r] = (T)*reinterpret_cast<const int32_t *>(p + r * stride);
                }
            } else if (a.fast_fmt == 2) {
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    row[r] = (T)*reinterpret_cast<const float *>(p + r * stride);
                }
            } else if (a.fast_fmt != 3) {
                for (int r = 0; r < RW; r++) {
                    row[r] = decode_sample<T>(load_raw_le(p + r * stride, f.bytes), f.bytes, f.isfloat, f.swap);
                }
            }
            if (a.muted != nullptr && a.muted[c]) {
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    row[r] = (T)0;          // a muted input reads as silence (bfrun.c:1523-1525)
                }
            }
            if (a.amax != nullptr) {
                // powersave: the block's peak, rounded up to float so that "not zero" stays "not zero"
                float m = 0.f;
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    m = fmaxf(m, peak_as_float(row[r]));
                }
                unsigned int *slot = a.amax + (size_t)blk * a.n_in + c;
                const unsigned int bits = __float_as_uint(m);       // non-negative floats order like their bit patterns
                if (bits > *slot) {
                    atomicMax(slot, bits);
                }
            }
        }
    }
    __syncwarp();
    // planar side: lanes j = lane % RW along time, RW-lane groups over 32 / RW channels at a time
    const int j = lane % RW, r0 = lane / RW;
    T *xt = reinterpret_cast<T *>(a.xt) + ((size_t)blk * a.n_in + c0) * a.L + n0 + j;
    const int nc = min(32, a.n_in - c0);
#pragma unroll
    for (int i = 0; i < RW; i++) {
        const int r = r0 + (32 / RW) * i;
        if (r < nc) {
            xt[(size_t)r * a.L] = tile[w][r * TS + j];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(32 * WPB) k_pack(InverseArgs a, int L)
{
    __shared__ T tile[WPB][32 * TS];
    __shared__ QuantStats wstats[WPB][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n0 = (blockIdx.x * WPB + w) * RW, c0 = blockIdx.y * 32, blk = blockIdx.z;
    const int nc = min(32, a.n_out - c0);
    QuantStats st;
    quant_stats_init(st);
    if (n0 < L) {
        const int j = lane % RW, r0 = lane / RW;
        const T *src = reinterpret_cast<const T *>(a.out_time) + ((size_t)blk * a.n_out + c0) * L + n0 + j;
#pragma unroll
        for (int i = 0; i < RW; i++) {
            const int r = r0 + (32 / RW) * i;
            if (r < nc) {
                tile[w][r * TS + j] = src[(size_t)r * L];
            }
        }
        __syncwarp();
        const int c = c0 + lane;
        if (a.fast_fmt == 3) {
            // packed 24-bit little-endian on all (undithered, ungrouped) outputs: quantise per lane, then 3 nc / 4 lanes
            // assemble and store the tile row -- word w holds the tail of channel 4w/3 and the head of the next one.
            // All 32 lanes take part in the shuffles; lanes beyond the last channel carry zeros.
            const bool live = c < a.n_out;
            const int32_t imin = -(1 << 23), imax = (1 << 23) - 1;
            const double rmin = (double)(T)imin, rmax = (double)(T)imax;
            const double of_max = live ? a.overflow[c].max : 1.0;
            const float thr = 4194303.0f;
            const bool limit = a.safety_limit != 0.0;
            const int ca = (4 * lane) / 3, sh = ((4 * lane) % 3) * 8;
            const size_t stride = (size_t)a.n_out * 3;
            uint8_t *rowp = a.raw_out + (size_t)blk * a.out_stride + (size_t)n0 * stride + (size_t)c0 * 3;
            const T *row = &tile[w][lane * TS];
#pragma unroll
            for (int r = 0; r < RW; r++) {
                const T y = live ? row[r] : (T)0;
                int32_t q, cand;
                if (!limit && real_to_int_fast(y, thr, q, cand)) {
                    st.intlargest = max(st.intlargest, cand);
                } else {
                    sample_test<T>(y, a.safety_limit, of_max, st);
                    q = real_to_int<T>(y, rmin, rmax, imin, imax, st);
                }
                const uint32_t q24 = (uint32_t)q & 0xffffffu;
                const uint32_t qa = __shfl_sync(0xffffffffu, q24, ca & 31), qb = __shfl_sync(0xffffffffu, q24, (ca + 1) & 31);
                if (4 * lane < 3 * nc) {        // bytes [o, o + 4) of (qa's three bytes | qb's three bytes), o = sh / 8
                    reinterpret_cast<uint32_t *>(rowp + r * stride)[lane] = (qa >> sh) | (qb << (24 - sh));
                }
            }
        } else if (c < a.n_out && !(a.chans[c].shared & 6)) {      // dithered outputs are k_dither's; bit 2: mixed into another channel
            const SampleFormat f = a.fmt[c];
            const size_t stride = (size_t)f.sample_spacing * f.bytes;
            uint8_t *p = a.raw_out + (size_t)blk * a.out_stride + f.byte_offset + (size_t)n0 * stride;
            const double of_max = a.overflow[c].max;
            const T *row = &tile[w][lane * TS];
            if (a.fast_fmt == 1) {
                const int bits_n = f.sbytes << 3;
                const int32_t imin = (int32_t)(-((uint64_t)1 << (bits_n - 1)));
                const int32_t imax = (int32_t)(((uint64_t)1 << (bits_n - 1)) - 1);
                const double rmin = (double)(T)imin, rmax = (double)(T)imax;
                // float samples inside +-(2^22 - 1) take the single-precision form of the quantiser (same results, a
                // quarter of the instructions: k_pack was instruction bound, 58 per sample); the safety limit, when
                // set, is never below full scale times a factor >= 1, but it is checked all the same
                const float thr = fminf(4194303.0f, (float)(imax - 1));
                const bool limit = a.safety_limit != 0.0;
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    const T y = row[r];
                    int32_t q, cand;
                    if (!limit && real_to_int_fast(y, thr, q, cand)) {
                        st.intlargest = max(st.intlargest, cand);
                    } else {
                        sample_test<T>(y, a.safety_limit, of_max, st);
                        q = real_to_int<T>(y, rmin, rmax, imin, imax, st);
                    }
                    *reinterpret_cast<int32_t *>(p + r * stride) = q;
                }
            } else if (a.fast_fmt == 2) {
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    const T y = row[r];
                    sample_test<T>(y, a.safety_limit, of_max, st);
                    float_overflow_update<T>(y, (T)-of_max, (T)of_max, st);
                    *reinterpret_cast<float *>(p + r * stride) = (float)y;
                }
            } else {
                for (int r = 0; r < RW; r++) {
                    const T y = row[r];
                    store_raw_le(p + r * stride,
                                 encode_sample<T>(y, f.bytes, f.sbytes, f.isfloat, f.swap, a.safety_limit, of_max, st),
                                 f.bytes);
                }
            }
        }
    }
    // the eight warps of a block hold the same 32 channels: combine, then one set of atomics per channel -- and
    // only where the block actually raises a maximum / counted something (sum and max commute, so the totals equal
    // the reference's sequential running count and maxima, real2raw.h:44-59, dither_funs.h:70-114)
    wstats[w][lane] = st;
    __syncthreads();
    if (w == 0 && c0 + lane < a.n_out) {
#pragma unroll
        for (int i = 1; i < WPB; i++) {
            const QuantStats o = wstats[i][lane];
            st.n_overflows += o.n_overflows;
            st.intlargest = max(st.intlargest, o.intlargest);
            st.largest = fmax(st.largest, o.largest);
            st.status |= o.status;
        }
        Overflow *of = &a.overflow[c0 + lane];
        if (st.n_overflows != 0) {
            atomicAdd(&of->n_overflows, st.n_overflows);
        }
        if (st.intlargest > of->intlargest) {
            atomicMax(&of->intlargest, st.intlargest);
        }
        if (st.largest > of->largest) {
            // non-negative doubles order like their bit patterns
            atomicMax(reinterpret_cast<unsigned long long *>(&of->largest),
                      (unsigned long long)__double_as_longlong(st.largest));
        }
        if (st.status != 0) {
            atomicOr(a.status, st.status);
        }
    }
}

// ======================================================================================================
// k_forward2
// ======================================================================================================

// frame = [previous block | this block] (fftw_convolver.c:180-193) packed as z_i = x_2i + i x_2i+1; a thread's
// first-pass butterfly takes z[tid + q NT]: q < 8 lies in the previous block, q >= 8 in this one.
template <typename T, int LOG2M>
__device__ __forceinline__ void load_frame(const ForwardArgs &a, int item, int tid, cpx<T> *v)
{
    typedef Fft2<LOG2M> F;
    constexpr int L = F::M, NT = F::NT;
    const int c = item % a.n_in, blk = item / a.n_in;
    const T *cur = reinterpret_cast<const T *>(a.xt_cur) + ((size_t)blk * a.n_in + c) * L;
    const T *old = blk == 0 ? reinterpret_cast<const T *>(a.xt_prev) + (size_t)c * L
                            : cur - (size_t)a.n_in * L;
    const cpx<T> *po = reinterpret_cast<const cpx<T> *>(old), *pc = reinterpret_cast<const cpx<T> *>(cur);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        v[q] = po[tid + q * NT];
        v[8 + q] = pc[tid + q * NT];
    }
}

// MINB = 2 ("slim"): at most 64 registers per thread and no register prefetch of the next frame, so that a block fits
// beside one resident block of the batched MAC (half an SM's registers) instead of waiting for an empty SM
template <typename T, int LOG2M, bool SINGLE, bool TWS, int MINB = 1>
__global__ void __launch_bounds__(Fft2<LOG2M>::CTA, MINB) k_forward2(ForwardArgs a, const cpx<T> *__restrict__ tw_global)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M, N = 2 * F::M;
    // holding the next frame in registers while the current one is stored: not with 1024 threads (64 registers each)
    // and not in double precision (the 16 points alone are 64 registers)
    constexpr bool PREFETCH = F::NT < 1024 && sizeof(T) == 4 && MINB == 1;
    extern __shared__ __align__(128) unsigned char smem2[];
    const int tid = threadIdx.x % F::NT, sub = threadIdx.x / F::NT;     // thread of its transform, transform of the block
    cpx<T> *s = reinterpret_cast<cpx<T> *>(smem2) + sub * F::SUB_STRIDE;
    const cpx<T> *tw = stage_twiddles<T, LOG2M, TWS>(smem2, tw_global, threadIdx.x);
    const int total = a.n_in * a.batch;
    const int stride = (int)gridDim.x * F::SUBS;
    T *fdl = reinterpret_cast<T *>(a.fdl);
    const int ring = a.ring;

    cpx<T> v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        v[q].x = (T)0;
        v[q].y = (T)0;
    }
    int item = (int)blockIdx.x * F::SUBS + sub;
    if (item < total) {
        load_frame<T, LOG2M>(a, item, tid, v);
    }
    wait_twiddles<T, LOG2M, TWS>(smem2);
    // every transform of the block takes the same number of turns (the barriers are block wide); one without an item
    // left runs on whatever its registers hold and stores nothing
    for (int base = (int)blockIdx.x * F::SUBS; base < total; base += stride, item += stride) {
        const bool active = item < total;
        const int c = active ? item % a.n_in : 0, blk = active ? item / a.n_in : 0;
        if (!PREFETCH && active && base != (int)blockIdx.x * F::SUBS) {
            load_frame<T, LOG2M>(a, item, tid, v);
        }
        fft2_complex<T, LOG2M, false, false>(s, tw, tid, v, BlockSync2());
        // the next transform's samples travel while this one's spectrum is split and stored
        if (PREFETCH && item + stride < total) {
            load_frame<T, LOG2M>(a, item + stride, tid, v);
        }
        if (active) {
            const int d0 = a.dest_first[c], d1 = a.dest_first[c + 1];
            const int t = a.t + blk;
            // powersave: a silent frame leaves zeros (the reference memsets, bfrun.c:1548-1552) and flags its slots
            const bool silent = frame_silent(a, c, blk);
            if (a.slot_zero != nullptr && tid == 0) {
                for (int d = d0; d < d1; d++) {
                    const FwdDest ds = a.dests[d];
                    a.slot_zero[(size_t)ds.stream * ring + (t + ds.delay) % ring] = silent ? 1 : 0;
                }
            }
            if (SINGLE) {
                // one filter per input: one destination, hoisted out of the bin loop
                const FwdDest ds = a.dests[d0];
                T *dst = fdl + ((size_t)ds.stream * ring + (t + ds.delay) % ring) * N;
                const T sc = silent ? (T)0 : (T)ds.scale;
                fft2_split_emit<T, LOG2M>(s, tw, tid, [&](int k, T re, T im) {
                    dst[k] = silent ? (T)0 : mul_rn(re, sc);
                    dst[M + k] = silent ? (T)0 : mul_rn(im, sc);
                });
            } else {
                T *xin = (a.xin != nullptr && a.need_xin[c])
                                 ? reinterpret_cast<T *>(a.xin) + ((size_t)blk * a.n_vin + c) * N : nullptr;
                const FwdDest *dests = a.dests;
                fft2_split_emit<T, LOG2M>(s, tw, tid, [&](int k, T re, T im) {
                    if (silent) {
                        re = (T)0;
                        im = (T)0;
                    }
                    if (xin != nullptr) {
                        xin[k] = re;
                        xin[M + k] = im;
                    }
                    for (int d = d0; d < d1; d++) {
                        const FwdDest ds = dests[d];
                        T *dst = fdl + ((size_t)ds.stream * ring + (t + ds.delay) % ring) * N;
                        const T sc = (T)ds.scale;
                        dst[k] = silent ? (T)0 : mul_rn(re, sc);
                        dst[M + k] = silent ? (T)0 : mul_rn(im, sc);
                    }
                });
            }
        }
        __syncthreads();        // the split phase has finished reading shared memory
    }
}

// ======================================================================================================
// k_inverse2
// ======================================================================================================

// SIMPLE: every output is fed by exactly one filter, the partition sum is not split and no crossfade is pending
// (the usual block): one scaled spectrum per transform, and the next transform's spectrum is fetched while this
// one's samples are stored.
template <typename T, int LOG2M, bool SIMPLE, bool TWS, int MINB = 1>
__global__ void __launch_bounds__(Fft2<LOG2M>::CTA, MINB) k_inverse2(InverseArgs a, const cpx<T> *__restrict__ tw_global)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M, L = F::M, N = 2 * F::M, NT = F::NT;
    constexpr int RL = F::radix(F::NP - 1);         // radix of the last pass
    constexpr int BPT = 16 / RL, HALF = RL / 2;      // butterflies per thread, valid outputs per butterfly
    constexpr bool PREFETCH = F::NT < 1024 && sizeof(T) == 4 && MINB == 1;
    extern __shared__ __align__(128) unsigned char smem2[];
    const int tid = threadIdx.x % NT, sub = threadIdx.x / NT;
    cpx<T> *s = reinterpret_cast<cpx<T> *>(smem2) + sub * F::SUB_STRIDE;
    const cpx<T> *tw = stage_twiddles<T, LOG2M, TWS>(smem2, tw_global, threadIdx.x);
    const int total = a.n_out * a.batch;
    const int stride = (int)gridDim.x * F::SUBS;
    const int first_item = (int)blockIdx.x * F::SUBS;
    const int zstride = a.batch * a.n_slots;        // Y slots between two partial sums of the split
    cpx<T> v[16];

    // overlap-save: only the first L samples are output (fftw_convolver.c:498-501); they are elements
    // i = (tid + b NT) + q M/RL, q < RL/2, of the complex result, sample 2i in .x and 2i + 1 in .y
    auto store_time = [&](int o, int blk, const cpx<T> *keep) {
        T *tdst = reinterpret_cast<T *>(a.out_time) + ((size_t)blk * a.n_out + o) * L;
#pragma unroll
        for (int b = 0; b < BPT; b++) {
#pragma unroll
            for (int q = 0; q < HALF; q++) {
                const int i = (tid + b * NT) + q * (M / RL);
                cpx<T> y = v[b * RL + q];
                if (keep != nullptr) {
                    y.x = xfade<T>(keep[b * HALF + q].x, y.x, 2 * i, L);
                    y.y = xfade<T>(keep[b * HALF + q].y, y.y, 2 * i + 1, L);
                }
                *reinterpret_cast<cpx<T> *>(tdst + 2 * i) = y;
            }
        }
    };

    // Every transform of the block takes the same number of turns (block-wide barriers); one without an item left
    // transforms whatever its registers hold and stores nothing.
    if (SIMPLE) {
        T x[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
            x[i] = (T)0;
        }
        T sc = (T)0;
        auto fetch = [&](int item) {
            const int o = item % a.n_out, blk = item / a.n_out;
            const MixTerm tm = a.terms[a.chans[o].first];
            const T *y = reinterpret_cast<const T *>(a.Y) + ((size_t)blk * a.n_slots + tm.index) * N;
            sc = (T)tm.scale;
            fft2_merge_fetch<T, LOG2M>(tid, x, [&](int i) { return __ldg(y + i); });
        };
        int item = first_item + sub;
        if (item < total) {
            fetch(item);
        }
        wait_twiddles<T, LOG2M, TWS>(smem2);
        for (int base = first_item; base < total; base += stride, item += stride) {
            const bool active = item < total;
            if (!PREFETCH && active && base != first_item) {
                fetch(item);
            }
#pragma unroll
            for (int i = 0; i < 32; i++) {
                x[i] = mul_rn(x[i], sc);        // mixnscale(OUTPUT) with one term (fftw_convfuns.h:268-494)
            }
            fft2_merge_store<T, LOG2M>(s, tw, tid, x);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; q++) {
                v[q] = s[tid + q * NT];
            }
            __syncthreads();        // everybody holds its inputs: pass 0 may overwrite
            fft2_complex<T, LOG2M, true, true>(s, tw, tid, v, BlockSync2());
            const int o = active ? item % a.n_out : 0, blk = active ? item / a.n_out : 0;
            if (PREFETCH && item + stride < total) {
                fetch(item + stride);
            }
            if (active) {
                store_time(o, blk, nullptr);
            }
            __syncthreads();        // the last pass has finished reading shared memory
        }
    } else {
        cpx<T> keep[BPT * HALF];
        wait_twiddles<T, LOG2M, TWS>(smem2);
        // a launch that contains a crossfading output runs two transforms for EVERY item (same turn count for all
        // transforms of a block); items without a crossfade transform their mix twice and use the second
        const int npass = a.any_xfade ? 2 : 1;
        int item = first_item + sub;
        for (int base = first_item; base < total; base += stride, item += stride) {
            const bool active = item < total;
            const int o = active ? item % a.n_out : 0, blk = active ? item / a.n_out : 0;
            const OutChan ch = a.chans[o];
            const T *Y = reinterpret_cast<const T *>(a.Y) + (size_t)blk * a.n_slots * N;
            for (int pass = 0; pass < npass; pass++) {
                const int term0 = (ch.xf_first >= 0 && pass + 1 < npass) ? ch.xf_first : ch.first;
                if (active) {
                    fft2_merge_load<T, LOG2M>(s, tw, tid, [&](int i) {
                        return mix_terms<T>(Y, a.terms, term0, ch.n, zstride, a.split, N, i);
                    });
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    v[q] = s[tid + q * NT];
                }
                __syncthreads();    // everybody holds its inputs: pass 0 may overwrite
                fft2_complex<T, LOG2M, true, true>(s, tw, tid, v, BlockSync2());
                if (pass + 1 < npass) {
#pragma unroll
                    for (int b = 0; b < BPT; b++) {
#pragma unroll
                        for (int q = 0; q < HALF; q++) {
                            keep[b * HALF + q] = v[b * RL + q];
                        }
                    }
                }
                __syncthreads();    // the last pass has finished reading shared memory
            }
            if (active) {
                store_time(o, blk, (npass == 2 && ch.xf_first >= 0) ? keep : nullptr);
            }
        }
    }
}

// ======================================================================================================
// plan tables and launchers
// ======================================================================================================

bool fft2_supported(int N, int realsize)
{
    const char *e = getenv("BFCUDA_FFT_V1");
    if (e != nullptr && atoi(e) != 0) {
        return false;
    }
    if ((N & (N - 1)) != 0 || N < 128) {
        return false;
    }
    // float: M = N/2 up to 16384 complex points in shared memory; double: up to 8192
    return realsize == 4 ? N <= 32768 : N <= 16384;
}

template <typename T, int LOG2M>
static cudaError_t make_table(void **out)
{
    typedef Fft2<LOG2M> F;
    std::vector<cpx<T>> h((size_t)F::TW_TOTAL);
    fft2_fill_table<T, LOG2M>(h.data());
    cudaError_t err = cudaMalloc(out, h.size() * sizeof(cpx<T>));
    if (err != cudaSuccess) {
        return err;
    }
    return cudaMemcpy(*out, h.data(), h.size() * sizeof(cpx<T>), cudaMemcpyHostToDevice);
}

template <typename T>
static cudaError_t make_table_for(void **out, int N)
{
    switch (N) {
    case 128: return make_table<T, 6>(out);
    case 256: return make_table<T, 7>(out);
    case 512: return make_table<T, 8>(out);
    case 1024: return make_table<T, 9>(out);
    case 2048: return make_table<T, 10>(out);
    case 4096: return make_table<T, 11>(out);
    case 8192: return make_table<T, 12>(out);
    case 16384: return make_table<T, 13>(out);
    default: return make_table<T, 14>(out);
    }
}

cudaError_t fft2_plan_create(FftPlan *plan)
{
    plan->tw2 = nullptr;
    if (!fft2_supported(plan->N, plan->realsize)) {
        return cudaSuccess;
    }
    return plan->realsize == 4 ? make_table_for<float>(&plan->tw2, plan->N) : make_table_for<double>(&plan->tw2, plan->N);
}

static int sm_count_of_current_device()
{
    static int cached[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) {
        return 148;
    }
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            n = 148;
        }
        cached[dev] = n;
    }
    return cached[dev];
}

template <typename K>
static cudaError_t persistent_grid(K kernel, int threads, size_t smem, int total, int *grid)
{
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
        return err;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (err != cudaSuccess) {
        return err;
    }
    if (per_sm < 1) {
        return cudaErrorLaunchOutOfResources;
    }
    const int resident = per_sm * sm_count_of_current_device();
    *grid = total < resident ? total : resident;
    return cudaSuccess;
}

template <typename T, int LOG2M, bool SINGLE, bool TWS, int MINB = 1>
static cudaError_t launch_forward2_t(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    typedef Fft2<LOG2M> F;
    static int resident[64];
    const size_t smem = Smem2<T, LOG2M, TWS>::total;
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (resident[dev] == 0) {
        cudaError_t err = persistent_grid(k_forward2<T, LOG2M, SINGLE, TWS, MINB>, F::CTA, smem, 1 << 30, &resident[dev]);
        if (err != cudaSuccess) return err;
    }
    const int total = (a.n_in * a.batch + F::SUBS - 1) / F::SUBS;      // blocks needed: SUBS transforms each
    const int grid = total < resident[dev] ? total : resident[dev];
    g_last_func = (const void *)k_forward2<T, LOG2M, SINGLE, TWS, MINB>;
    k_forward2<T, LOG2M, SINGLE, TWS, MINB><<<grid, F::CTA, smem, s>>>(a, reinterpret_cast<const cpx<T> *>(plan.tw2));
    return cudaGetLastError();
}

template <typename T, int LOG2M, bool SIMPLE, bool TWS, int MINB = 1>
static cudaError_t launch_inverse2_t(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    typedef Fft2<LOG2M> F;
    static int resident[64];
    const size_t smem = Smem2<T, LOG2M, TWS>::total;
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (resident[dev] == 0) {
        cudaError_t err = persistent_grid(k_inverse2<T, LOG2M, SIMPLE, TWS, MINB>, F::CTA, smem, 1 << 30, &resident[dev]);
        if (err != cudaSuccess) return err;
    }
    const int total = (a.n_out * a.batch + F::SUBS - 1) / F::SUBS;
    const int grid = total < resident[dev] ? total : resident[dev];
    k_inverse2<T, LOG2M, SIMPLE, TWS, MINB><<<grid, F::CTA, smem, s>>>(a, reinterpret_cast<const cpx<T> *>(plan.tw2));
    return cudaGetLastError();
}

// float: twiddle tables resident in shared memory up to M = 8192 (98 KB of tables beside 64 KB of data); double: up to
// M = 4096 (the same byte counts), M = 8192 reads them through L1 (128 KB of data fills the block's shared memory)
#define BF_FFT2_SIZES(FN, FLAG, ...)                                                          \
    if (plan.realsize == 4) {                                                                 \
        switch (plan.N) {                                                                     \
        case 128: return FN<float, 6, FLAG, true>(__VA_ARGS__);                               \
        case 256: return FN<float, 7, FLAG, true>(__VA_ARGS__);                               \
        case 512: return FN<float, 8, FLAG, true>(__VA_ARGS__);                               \
        case 1024: return FN<float, 9, FLAG, true>(__VA_ARGS__);                              \
        case 2048: return FN<float, 10, FLAG, true>(__VA_ARGS__);                             \
        case 4096: return FN<float, 11, FLAG, true>(__VA_ARGS__);                             \
        case 8192: return FN<float, 12, FLAG, true>(__VA_ARGS__);                             \
        case 16384: return FN<float, 13, FLAG, true>(__VA_ARGS__);                            \
        case 32768: return FN<float, 14, FLAG, false>(__VA_ARGS__);                           \
        default: return cudaErrorInvalidValue;                                                \
        }                                                                                     \
    }                                                                                         \
    switch (plan.N) {                                                                         \
    case 128: return FN<double, 6, FLAG, true>(__VA_ARGS__);                                  \
    case 256: return FN<double, 7, FLAG, true>(__VA_ARGS__);                                  \
    case 512: return FN<double, 8, FLAG, true>(__VA_ARGS__);                                  \
    case 1024: return FN<double, 9, FLAG, true>(__VA_ARGS__);                                 \
    case 2048: return FN<double, 10, FLAG, true>(__VA_ARGS__);                                \
    case 4096: return FN<double, 11, FLAG, true>(__VA_ARGS__);                                \
    case 8192: return FN<double, 12, FLAG, true>(__VA_ARGS__);                                \
    case 16384: return FN<double, 13, FLAG, false>(__VA_ARGS__);                              \
    default: return cudaErrorInvalidValue;                                                    \
    }

static bool fft2_slim()
{
    static const int v = [] { const char *e = getenv("BFCUDA_FFT_SLIM"); return e != nullptr ? atoi(e) : 0; }();
    return v != 0;
}

cudaError_t launch_unpack(const FftPlan &plan, const UnpackArgs &a, cudaStream_t s)
{
    if (a.n_in == 0) return cudaSuccess;
    dim3 grid((a.L + RW * WPB - 1) / (RW * WPB), (a.n_in + 31) / 32, a.batch);
    if (plan.realsize == 4) {
        k_unpack<float><<<grid, 32 * WPB, 0, s>>>(a);
    } else {
        k_unpack<double><<<grid, 32 * WPB, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_pack(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    const int L = plan.N / 2;
    dim3 grid((L + RW * WPB - 1) / (RW * WPB), (a.n_out + 31) / 32, a.batch);
    if (plan.realsize == 4) {
        k_pack<float><<<grid, 32 * WPB, 0, s>>>(a, L);
    } else {
        k_pack<double><<<grid, 32 * WPB, 0, s>>>(a, L);
    }
    return cudaGetLastError();
}

cudaError_t launch_forward2(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    if (a.n_in == 0) return cudaSuccess;
    if (fft2_slim() && plan.realsize == 4 && plan.N == 16384) {
        return a.single_dest ? launch_forward2_t<float, 13, true, false, 2>(plan, a, s)
                             : launch_forward2_t<float, 13, false, false, 2>(plan, a, s);
    }
    if (a.single_dest) {
        BF_FFT2_SIZES(launch_forward2_t, true, plan, a, s)
    }
    BF_FFT2_SIZES(launch_forward2_t, false, plan, a, s)
}

cudaError_t launch_inverse2(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    if (fft2_slim() && plan.realsize == 4 && plan.N == 16384) {
        return a.simple_mix ? launch_inverse2_t<float, 13, true, false, 2>(plan, a, s)
                            : launch_inverse2_t<float, 13, false, false, 2>(plan, a, s);
    }
    if (a.simple_mix) {
        BF_FFT2_SIZES(launch_inverse2_t, true, plan, a, s)
    }
    BF_FFT2_SIZES(launch_inverse2_t, false, plan, a, s)
}

}  // namespace bf
