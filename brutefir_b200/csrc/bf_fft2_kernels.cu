// bf_fft2_kernels.cu -- the forward and inverse stages on the size-specialised FFT core (bf_fft2.cuh).
//
//   k_forward2   raw2real + frame assembly + R2HC + input scale into the delay line
//                (K1..K3; /root/reference/bfrun.c:1494-1560, 1671; fftw_convolver.c:170-214; raw2real.h)
//   k_inverse2   output mix + HC2R + overlap-save discard + crossfade + real2raw
//                (K6..K9; bfrun.c:1847-1936; fftw_convolver.c:330-368, 391-409, 482-518; real2raw.h)
//
// Same results contract as k_forward / k_inverse in bf_kernels.cu (which stay as the path for float_bits 64 and
// for partitions shorter than 1024 samples); what changes is the cost:
//   * persistent blocks: grid = min(transforms, resident blocks), each block walks its transforms, the twiddle
//     table is bulk-copied (cp.async.bulk + mbarrier) into shared memory once per block;
//   * forward: the first radix-16 pass reads its samples straight from the raw block / the previous-block buffer;
//   * inverse: the last pass leaves its results in registers -- only the L valid samples of the overlap-save
//     frame are produced -- and the crossfade / quantise / pack epilogue consumes them there;
//   * sample formats: the two layouts every shipped config uses (aligned 4-byte little-endian integers --
//     S32_LE and S24_4LE -- and FLOAT_LE) are compile-time fast paths; everything else takes the generic
//     per-sample routines of bf_sample.cuh.
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

enum { FMT_GENERIC = 0, FMT_INT32LE = 1, FMT_FLOAT32LE = 2 };

struct BlockSync2 {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// shared memory: [M complex data][TW_TOTAL complex twiddles (TWS only)][32 QuantStats][mbarrier]
template <int LOG2M, bool TWS>
struct Smem2 {
    typedef Fft2<LOG2M> F;
    static constexpr size_t data_bytes = (size_t)F::M * sizeof(cpx<float>);
    static constexpr size_t tw_bytes = TWS ? (size_t)F::TW_TOTAL * sizeof(cpx<float>) : 0;
    static constexpr size_t stats_off = data_bytes + tw_bytes;
    static constexpr size_t bar_off = stats_off + 32 * sizeof(QuantStats);
    static constexpr size_t total = bar_off + 16;
};

template <int LOG2M, bool TWS>
__device__ __forceinline__ const cpx<float> *stage_twiddles(unsigned char *smem, const cpx<float> *tw_global, int tid)
{
    typedef Smem2<LOG2M, TWS> S;
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
    return reinterpret_cast<const cpx<float> *>(dst);
}

template <int LOG2M, bool TWS>
__device__ __forceinline__ void wait_twiddles(unsigned char *smem)
{
    if (TWS) {
        mbar_wait(reinterpret_cast<uint64_t *>(smem + Smem2<LOG2M, TWS>::bar_off), 0);
    }
}

// two consecutive samples n, n + 1 of one channel of a raw block -> complex (raw2real.h:7-160; integers unscaled)
template <int FMT>
__device__ __forceinline__ cpx<float> load_pair(const uint8_t *chan_base, size_t stride, int n, const SampleFormat &f)
{
    cpx<float> z;
    const uint8_t *p = chan_base + (size_t)n * stride;
    if (FMT == FMT_INT32LE) {
        z.x = (float)*reinterpret_cast<const int32_t *>(p);
        z.y = (float)*reinterpret_cast<const int32_t *>(p + stride);
    } else if (FMT == FMT_FLOAT32LE) {
        z.x = *reinterpret_cast<const float *>(p);
        z.y = *reinterpret_cast<const float *>(p + stride);
    } else {
        z.x = decode_sample<float>(load_raw_le(p, f.bytes), f.bytes, f.isfloat, f.swap);
        z.y = decode_sample<float>(load_raw_le(p + stride, f.bytes), f.bytes, f.isfloat, f.swap);
    }
    return z;
}

// ======================================================================================================
// k_forward2
// ======================================================================================================

template <int LOG2M, int FMT, bool TWS>
__global__ void __launch_bounds__(Fft2<LOG2M>::NT, 1) k_forward2(ForwardArgs a, const cpx<float> *__restrict__ tw_global)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M, L = F::M, N = 2 * F::M, NT = F::NT;
    extern __shared__ __align__(128) unsigned char smem2[];
    cpx<float> *s = reinterpret_cast<cpx<float> *>(smem2);
    const int tid = threadIdx.x;
    const cpx<float> *tw = stage_twiddles<LOG2M, TWS>(smem2, tw_global, tid);
    const int total = a.n_in * a.batch;
    bool first = true;

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int c = item % a.n_in, blk = item / a.n_in;
        const SampleFormat f = a.fmt[c];
        const size_t stride = (size_t)f.sample_spacing * f.bytes;
        const uint8_t *raw = a.raw_in + (size_t)blk * a.in_stride + f.byte_offset;
        const bool last = blk == a.batch - 1;

        // frame = [previous block | this block] (fftw_convolver.c:180-193) packed as z_i = x_2i + i x_2i+1; this
        // thread's first-pass butterfly takes z[tid + q NT]: q < 8 lies in the previous block, q >= 8 in this one.
        cpx<float> v[16];
        if (blk == 0) {
            const cpx<float> *pv = reinterpret_cast<const cpx<float> *>(reinterpret_cast<const float *>(a.prev_in) +
                                                                         (size_t)c * L);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v[q] = pv[tid + q * NT];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v[q] = load_pair<FMT>(raw - a.in_stride, stride, 2 * (tid + q * NT), f);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            v[8 + q] = load_pair<FMT>(raw, stride, 2 * (tid + q * NT), f);
        }
        if (last) {
            cpx<float> *po = reinterpret_cast<cpx<float> *>(reinterpret_cast<float *>(a.prev_out) + (size_t)c * L);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                po[tid + q * NT] = v[8 + q];
            }
        }
        if (first) {
            wait_twiddles<LOG2M, TWS>(smem2);
            first = false;
        } else {
            __syncthreads();        // the previous transform's split phase has finished reading shared memory
        }
        fft2_complex<float, LOG2M, false, false>(s, tw, tid, v, BlockSync2());

        const int d0 = a.dest_first[c], d1 = a.dest_first[c + 1];
        float *xin = (a.xin != nullptr && a.need_xin[c])
                         ? reinterpret_cast<float *>(a.xin) + ((size_t)blk * a.n_in + c) * N : nullptr;
        float *fdl = reinterpret_cast<float *>(a.fdl);
        const FwdDest *dests = a.dests;
        const int ring = a.ring;
        const int t = a.t + blk;
        if (d1 - d0 == 1 && xin == nullptr) {
            // the usual case, one filter per input: one destination, hoisted out of the bin loop
            const FwdDest ds = dests[d0];
            float *dst = fdl + ((size_t)ds.stream * ring + (t + ds.delay) % ring) * N;
            const float sc = (float)ds.scale;
            fft2_split_emit<float, LOG2M>(s, tw, tid, [&](int k, float re, float im) {
                dst[k] = mul_rn(re, sc);
                dst[M + k] = mul_rn(im, sc);
            });
        } else {
            fft2_split_emit<float, LOG2M>(s, tw, tid, [&](int k, float re, float im) {
                if (xin != nullptr) {
                    xin[k] = re;
                    xin[M + k] = im;
                }
                for (int d = d0; d < d1; d++) {
                    const FwdDest ds = dests[d];
                    float *dst = fdl + ((size_t)ds.stream * ring + (t + ds.delay) % ring) * N;
                    const float sc = (float)ds.scale;
                    dst[k] = mul_rn(re, sc);
                    dst[M + k] = mul_rn(im, sc);
                }
            });
        }
    }
}

// ======================================================================================================
// k_inverse2
// ======================================================================================================

// one output sample: test, quantise or copy, account, store (real2raw.h:61-251, dither_funs.h:70-114)
template <int FMT>
__device__ __forceinline__ void store_sample(float y, uint8_t *p, const SampleFormat &f, double safety_limit,
                                             double of_max, double rmin, double rmax, int32_t imin, int32_t imax,
                                             QuantStats &st)
{
    if (FMT == FMT_INT32LE) {
        sample_test<float>(y, safety_limit, of_max, st);
        *reinterpret_cast<int32_t *>(p) = real_to_int<float>(y, rmin, rmax, imin, imax, st);
    } else if (FMT == FMT_FLOAT32LE) {
        sample_test<float>(y, safety_limit, of_max, st);
        float_overflow_update<float>(y, (float)-of_max, (float)of_max, st);
        *reinterpret_cast<float *>(p) = y;
    } else {
        store_raw_le(p, encode_sample<float>(y, f.bytes, f.sbytes, f.isfloat, f.swap, safety_limit, of_max, st), f.bytes);
    }
}

template <int LOG2M, int FMT, bool TWS>
__global__ void __launch_bounds__(Fft2<LOG2M>::NT, 1) k_inverse2(InverseArgs a, const cpx<float> *__restrict__ tw_global)
{
    typedef Fft2<LOG2M> F;
    constexpr int M = F::M, L = F::M, N = 2 * F::M, NT = F::NT;
    constexpr int RL = F::radix(F::NP - 1);         // radix of the last pass
    constexpr int BPT = 16 / RL, HALF = RL / 2;      // butterflies per thread, valid outputs per butterfly
    extern __shared__ __align__(128) unsigned char smem2[];
    cpx<float> *s = reinterpret_cast<cpx<float> *>(smem2);
    void *stats_scratch = smem2 + Smem2<LOG2M, TWS>::stats_off;
    const int tid = threadIdx.x;
    const cpx<float> *tw = stage_twiddles<LOG2M, TWS>(smem2, tw_global, tid);
    const int total = a.n_out * a.batch;
    const int zstride = a.batch * a.n_slots;        // Y slots between two partial sums of the split
    bool first = true;

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int o = item % a.n_out, blk = item / a.n_out;
        const OutChan ch = a.chans[o];
        const float *Y = reinterpret_cast<const float *>(a.Y) + (size_t)blk * a.n_slots * N;
        const int npass = ch.xf_first >= 0 ? 2 : 1;
        cpx<float> v[16], keep[BPT * HALF];

        for (int pass = 0; pass < npass; pass++) {
            const int term0 = (npass == 2 && pass == 0) ? ch.xf_first : ch.first;
            if (first) {
                wait_twiddles<LOG2M, TWS>(smem2);
                first = false;
            } else {
                __syncthreads();    // the previous transform's last pass has finished reading shared memory
            }
            if (ch.n == 1 && a.split == 1) {
                // the usual case, one filter per output: one scaled spectrum, hoisted out of the bin loop
                const MixTerm tm = a.terms[term0];
                const float *y = Y + (size_t)tm.index * N;
                const float sc = (float)tm.scale;
                fft2_merge_load<float, LOG2M>(s, tw, tid, [&](int i) { return mul_rn(__ldg(y + i), sc); });
            } else {
                fft2_merge_load<float, LOG2M>(s, tw, tid, [&](int i) {
                    return mix_terms<float>(Y, a.terms, term0, ch.n, zstride, a.split, N, i);
                });
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; q++) {
                v[q] = s[tid + q * NT];
            }
            __syncthreads();        // everybody holds its inputs: pass 0 may overwrite
            fft2_complex<float, LOG2M, true, true>(s, tw, tid, v, BlockSync2());
            if (pass + 1 < npass) {
#pragma unroll
                for (int b = 0; b < BPT; b++) {
#pragma unroll
                    for (int q = 0; q < HALF; q++) {
                        keep[b * HALF + q] = v[b * RL + q];
                    }
                }
            }
        }

        // overlap-save: only the first L samples are output (fftw_convolver.c:498-501); they are elements
        // i = (tid + b NT) + q M/RL, q < RL/2, of the complex result, sample 2i in .x and 2i + 1 in .y
        const SampleFormat f = a.fmt[o];
        float *tdst = reinterpret_cast<float *>(a.out_time) + ((size_t)blk * a.n_out + o) * L;
        uint8_t *raw = a.raw_out + (size_t)blk * a.out_stride + f.byte_offset;
        const size_t stride = (size_t)f.sample_spacing * f.bytes;
        const double of_max = a.overflow[o].max;
        const int bits_n = f.sbytes << 3;
        const int32_t imin = (int32_t)(-((uint64_t)1 << (bits_n - 1)));
        const int32_t imax = (int32_t)(((uint64_t)1 << (bits_n - 1)) - 1);
        const double rmin = (double)(float)imin, rmax = (double)(float)imax;
        QuantStats st;
        quant_stats_init(st);
#pragma unroll
        for (int b = 0; b < BPT; b++) {
#pragma unroll
            for (int q = 0; q < HALF; q++) {
                const int i = (tid + b * NT) + q * (M / RL);
                cpx<float> y = v[b * RL + q];
                if (npass == 2) {
                    y.x = xfade<float>(keep[b * HALF + q].x, y.x, 2 * i, L);
                    y.y = xfade<float>(keep[b * HALF + q].y, y.y, 2 * i + 1, L);
                }
                *reinterpret_cast<cpx<float> *>(tdst + 2 * i) = y;
                if (!ch.shared) {
                    uint8_t *p = raw + (size_t)(2 * i) * stride;
                    store_sample<FMT>(y.x, p, f, a.safety_limit, of_max, rmin, rmax, imin, imax, st);
                    store_sample<FMT>(y.y, p + stride, f, a.safety_limit, of_max, rmin, rmax, imin, imax, st);
                }
            }
        }
        if (!ch.shared) {
            reduce_stats(st, &a.overflow[o], a.status, stats_scratch, tid, NT);
        }
    }
}

// ======================================================================================================
// plan tables and launchers
// ======================================================================================================

bool fft2_supported(int N, int realsize)
{
    if (realsize != 4) {
        return false;
    }
    const char *e = getenv("BFCUDA_FFT_V1");
    if (e != nullptr && atoi(e) != 0) {
        return false;
    }
    return N == 2048 || N == 4096 || N == 8192 || N == 16384 || N == 32768;
}

template <int LOG2M>
static cudaError_t make_table(void **out)
{
    typedef Fft2<LOG2M> F;
    std::vector<cpx<float>> h((size_t)F::TW_TOTAL);
    fft2_fill_table<float, LOG2M>(h.data());
    cudaError_t err = cudaMalloc(out, h.size() * sizeof(cpx<float>));
    if (err != cudaSuccess) {
        return err;
    }
    return cudaMemcpy(*out, h.data(), h.size() * sizeof(cpx<float>), cudaMemcpyHostToDevice);
}

cudaError_t fft2_plan_create(FftPlan *plan)
{
    plan->tw2 = nullptr;
    if (!fft2_supported(plan->N, plan->realsize)) {
        return cudaSuccess;
    }
    switch (plan->N) {
    case 2048: return make_table<10>(&plan->tw2);
    case 4096: return make_table<11>(&plan->tw2);
    case 8192: return make_table<12>(&plan->tw2);
    case 16384: return make_table<13>(&plan->tw2);
    default: return make_table<14>(&plan->tw2);
    }
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

template <int LOG2M, int FMT, bool TWS>
static cudaError_t launch_forward2_t(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    typedef Fft2<LOG2M> F;
    static int grid_cache[64][2];       // [device][0 = resident blocks]
    const size_t smem = Smem2<LOG2M, TWS>::total;
    int dev = 0;
    cudaGetDevice(&dev);
    const int total = a.n_in * a.batch;
    int grid = 0;
    if (dev >= 0 && dev < 64 && grid_cache[dev][0] > 0) {
        grid = total < grid_cache[dev][0] ? total : grid_cache[dev][0];
    } else {
        cudaError_t err = persistent_grid(k_forward2<LOG2M, FMT, TWS>, F::NT, smem, 1 << 30, &grid);
        if (err != cudaSuccess) return err;
        if (dev >= 0 && dev < 64) grid_cache[dev][0] = grid;
        grid = total < grid ? total : grid;
    }
    k_forward2<LOG2M, FMT, TWS><<<grid, F::NT, smem, s>>>(a, reinterpret_cast<const cpx<float> *>(plan.tw2));
    return cudaGetLastError();
}

template <int LOG2M, int FMT, bool TWS>
static cudaError_t launch_inverse2_t(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    typedef Fft2<LOG2M> F;
    static int grid_cache[64][2];
    const size_t smem = Smem2<LOG2M, TWS>::total;
    int dev = 0;
    cudaGetDevice(&dev);
    const int total = a.n_out * a.batch;
    int grid = 0;
    if (dev >= 0 && dev < 64 && grid_cache[dev][0] > 0) {
        grid = total < grid_cache[dev][0] ? total : grid_cache[dev][0];
    } else {
        cudaError_t err = persistent_grid(k_inverse2<LOG2M, FMT, TWS>, F::NT, smem, 1 << 30, &grid);
        if (err != cudaSuccess) return err;
        if (dev >= 0 && dev < 64) grid_cache[dev][0] = grid;
        grid = total < grid ? total : grid;
    }
    k_inverse2<LOG2M, FMT, TWS><<<grid, F::NT, smem, s>>>(a, reinterpret_cast<const cpx<float> *>(plan.tw2));
    return cudaGetLastError();
}

#define BF_FFT2_SIZES(FN, FMT, ...)                                                    \
    switch (plan.N) {                                                                  \
    case 2048: return FN<10, FMT, true>(__VA_ARGS__);                                  \
    case 4096: return FN<11, FMT, true>(__VA_ARGS__);                                  \
    case 8192: return FN<12, FMT, true>(__VA_ARGS__);                                  \
    case 16384: return FN<13, FMT, true>(__VA_ARGS__);                                 \
    case 32768: return FN<14, FMT, false>(__VA_ARGS__);                                \
    default: return cudaErrorInvalidValue;                                             \
    }

cudaError_t launch_forward2(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    if (a.n_in == 0) return cudaSuccess;
    if (a.fast_fmt == FMT_INT32LE) {
        BF_FFT2_SIZES(launch_forward2_t, FMT_INT32LE, plan, a, s)
    } else if (a.fast_fmt == FMT_FLOAT32LE) {
        BF_FFT2_SIZES(launch_forward2_t, FMT_FLOAT32LE, plan, a, s)
    }
    BF_FFT2_SIZES(launch_forward2_t, FMT_GENERIC, plan, a, s)
}

cudaError_t launch_inverse2(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    if (a.fast_fmt == FMT_INT32LE) {
        BF_FFT2_SIZES(launch_inverse2_t, FMT_INT32LE, plan, a, s)
    } else if (a.fast_fmt == FMT_FLOAT32LE) {
        BF_FFT2_SIZES(launch_inverse2_t, FMT_FLOAT32LE, plan, a, s)
    }
    BF_FFT2_SIZES(launch_inverse2_t, FMT_GENERIC, plan, a, s)
}

}  // namespace bf
