// bf_kernels.cu -- hand-written sm_100a kernels of the partitioned-convolution engine.
//
//   k_forward      raw2real + frame assembly + N-point real FFT + input scale, written straight into
//                  the frequency-domain delay line (K1+K2+K3 of SURVEY.md 2.2; bfrun.c:1494-1560, 1671)
//   k_stream_mix   delay-line slots that mix several inputs (mixnscale INPUT, n_bufs > 1)
//   k_mac          the delay-line complex multiply-accumulate over partitions (K4/K5; bfrun.c:1737-1754)
//   k_inverse      output mix + inverse real FFT + overlap-save discard + crossfade + quantise/pack
//                  (K6..K9; bfrun.c:1847-1936, fftw_convolver.c:330-368, real2raw.h)
//   k_coeff_fft    coefficient preprocessing (K10; fftw_convolver.c:526-573)
//   k_cv_*         the per-call convolver.h surface on the reference's own layouts
//
// Roofline notes (DESIGN.md has the full table): k_mac is a pure HBM stream, 16 B of operands per
// complex bin-partition and 8 "flop"; the FFT kernels move ~1 % of its bytes.  Tensor cores are not
// used: there is no contraction here, only element-wise complex products.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "bf_kernels.h"
#include "bf_fft.cuh"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"

namespace bf {

thread_local const void *g_last_func = nullptr;

// ======================================================================================================
// FFT plan (twiddle table)
// ======================================================================================================

bool fft_single_block_supported(int N, int realsize)
{
    if (N < 8 || (N & (N - 1)) != 0) {
        return false;
    }
    // one block holds the packed M = N/2 complex points in shared memory (227 KB max per block)
    return realsize == 4 ? N <= 32768 : N <= 16384;
}

bool fft_size_supported(int N, int realsize)
{
    return fft_single_block_supported(N, realsize) || fft_big_supported(N, realsize);     // bf_fft4.cu beyond one block
}

cudaError_t fft_plan_create(FftPlan *plan, int N, int realsize, int big_items_hint)
{
    plan->N = N;
    plan->realsize = realsize;
    plan->tw = nullptr;
    plan->tw2 = nullptr;
    plan->big_m1 = plan->big_m2 = plan->big_items = 0;
    plan->big_tw1 = plan->big_tw2 = plan->big_scr0 = plan->big_scr1 = plan->big_old = nullptr;
    const int half = N / 2;
    cudaError_t err = cudaMalloc(&plan->tw, (size_t)N * realsize);
    if (err != cudaSuccess) {
        return err;
    }
    std::vector<double> td((size_t)N);
    for (int j = 0; j < half; j++) {
        // roots in long double, rounded once to the real type
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)N;
        td[2 * (size_t)j] = (double)cosl(a);
        td[2 * (size_t)j + 1] = (double)sinl(a);
    }
    if (N >= 8) {
        td[2 * (size_t)(N / 4)] = 0.0;          // W^{N/4} = -i exactly
        td[2 * (size_t)(N / 4) + 1] = -1.0;
    }
    if (realsize == 4) {
        std::vector<float> tf((size_t)N);
        for (size_t i = 0; i < (size_t)N; i++) {
            tf[i] = (float)td[i];
        }
        err = cudaMemcpy(plan->tw, tf.data(), (size_t)N * 4, cudaMemcpyHostToDevice);
    } else {
        err = cudaMemcpy(plan->tw, td.data(), (size_t)N * 8, cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess) {
        return err;
    }
    if (!fft_single_block_supported(N, realsize)) {
        return fft_big_plan_create(plan, big_items_hint);
    }
    return fft2_plan_create(plan);
}

void fft_plan_destroy(FftPlan *plan)
{
    if (plan->tw != nullptr) {
        cudaFree(plan->tw);
        plan->tw = nullptr;
    }
    if (plan->tw2 != nullptr) {
        cudaFree(plan->tw2);
        plan->tw2 = nullptr;
    }
    fft_big_plan_destroy(plan);
}

// ======================================================================================================
// device building blocks
// ======================================================================================================

// forward real transform of the frame already packed in (sre, sim); calls emit(k, re, im) for every
// bin k in [0, M); the Nyquist value is passed as the imaginary part of bin 0 (planar convention).
template <typename T, int E, typename Emit>
__device__ __forceinline__ void forward_and_emit(T *sre, T *sim, const T *__restrict__ tw, int M, int tid, int nt,
                                                 Emit emit)
{
    fft_complex_inplace<T, E, false>(sre, sim, tw, M, tid, nt, BlockSync());
#pragma unroll 1
    for (int b = 0; b <= E / 2; b++) {
        const int k = tid + b * nt;
        if (k > M / 2) {
            break;
        }
        if (k == 0) {
            const T zr = sre[0], zi = sim[0];
            emit(0, zr + zi, zr - zi);
        } else {
            T wr, wi, xkr, xki, xmr, xmi;
            fft_twiddle<T>(tw, M, k, false, wr, wi);
            fft_split_pair<T>(sre[fft_pad(k)], sim[fft_pad(k)], sre[fft_pad(M - k)], sim[fft_pad(M - k)], wr, wi,
                              xkr, xki, xmr, xmi);
            emit(k, xkr, xki);
            if (k != M - k) {
                emit(M - k, xmr, xmi);
            }
        }
    }
}

// inverse real transform: load(i) returns planar element i of the spectrum; on return the time
// samples sit in shared memory as y[2j] = sre[pad(j)], y[2j+1] = sim[pad(j)], and a sync was issued.
template <typename T, int E, typename Load>
__device__ __forceinline__ void load_and_inverse(T *sre, T *sim, const T *__restrict__ tw, int M, int tid, int nt,
                                                 Load load)
{
#pragma unroll 1
    for (int b = 0; b <= E / 2; b++) {
        const int k = tid + b * nt;
        if (k > M / 2) {
            break;
        }
        if (k == 0) {
            const T x0 = load(0), xm = load(M);
            sre[0] = x0 + xm;
            sim[0] = x0 - xm;
        } else {
            T wr, wi, zkr, zki, zmr, zmi;
            const T xkr = load(k), xki = load(M + k);
            const T xmr = load(M - k), xmi = load(2 * M - k);
            fft_twiddle<T>(tw, M, k, false, wr, wi);
            fft_merge_pair<T>(xkr, xki, xmr, xmi, wr, wi, zkr, zki, zmr, zmi);
            sre[fft_pad(k)] = zkr;
            sim[fft_pad(k)] = zki;
            if (k != M - k) {
                sre[fft_pad(M - k)] = zmr;
                sim[fft_pad(M - k)] = zmi;
            }
        }
    }
    __syncthreads();
    fft_complex_inplace<T, E, true>(sre, sim, tw, M, tid, nt, BlockSync());
}

// ======================================================================================================
// k_forward
// ======================================================================================================

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_forward(ForwardArgs a, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int c = blockIdx.x, blk = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    const SampleFormat f = a.fmt[c];
    const T *prev_in = reinterpret_cast<const T *>(a.prev_in) + (size_t)c * L;
    T *prev_out = reinterpret_cast<T *>(a.prev_out) + (size_t)c * L;
    const uint8_t *raw = a.raw_in + (size_t)blk * a.in_stride + f.byte_offset;
    const size_t stride = (size_t)f.sample_spacing * f.bytes;
    const bool last = blk == a.batch - 1;

    // frame = [previous block | this block] (fftw_convolver.c:180-193), packed z_j = x_2j + i x_2j+1.
    // All loads of this thread's samples are issued before the first use: the interleaved layouts put every
    // sample of a channel in a different sector, so memory-level parallelism is what hides the latency.
    {
        uint64_t bits[E];
        T old[E];
#pragma unroll
        for (int b = 0; b < E; b++) {
            const int n = tid + b * nt;
            if (n < L) {
                bits[b] = load_raw_le(raw + (size_t)n * stride, f.bytes);
                // first half of the frame: the previous block -- kept in real form across calls for the first
                // block of a batch, re-read from the raw batch otherwise
                old[b] = blk == 0 ? prev_in[n]
                                  : decode_sample<T>(load_raw_le(raw - a.in_stride + (size_t)n * stride, f.bytes),
                                                     f.bytes, f.isfloat, f.swap);
            }
        }
#pragma unroll
        for (int b = 0; b < E; b++) {
            const int n = tid + b * nt;
            if (n < L) {
                const T cur = decode_sample<T>(bits[b], f.bytes, f.isfloat, f.swap);
                if (last) {
                    prev_out[n] = cur;
                }
                const int j0 = fft_pad(n >> 1), j1 = fft_pad((L + n) >> 1);
                if (n & 1) {
                    sim[j0] = old[b];
                    sim[j1] = cur;
                } else {
                    sre[j0] = old[b];
                    sre[j1] = cur;
                }
            }
        }
    }
    __syncthreads();

    const int d0 = a.dest_first[c], d1 = a.dest_first[c + 1];
    T *xin = (a.xin != nullptr && a.need_xin[c])
                 ? reinterpret_cast<T *>(a.xin) + ((size_t)blk * a.n_vin + c) * N : nullptr;
    T *fdl = reinterpret_cast<T *>(a.fdl);
    const FwdDest *dests = a.dests;
    const int ring = a.ring;
    const int t = a.t + blk;
    forward_and_emit<T, E>(sre, sim, tw, M, tid, nt, [&](int k, T re, T im) {
        if (xin != nullptr) {
            xin[k] = re;
            xin[M + k] = im;
        }
        for (int d = d0; d < d1; d++) {
            const FwdDest ds = dests[d];
            const int slot = (t + ds.delay) % ring;
            T *dst = fdl + ((size_t)ds.stream * ring + slot) * N;
            const T s = (T)ds.scale;
            dst[k] = mul_rn(re, s);
            dst[M + k] = mul_rn(im, s);
        }
    });
}

template <typename T>
__global__ void __launch_bounds__(256) k_stream_mix(StreamMixArgs a, int N)
{
    const MixStream ms = a.streams[blockIdx.y];
    const int blk = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) {
        return;
    }
    const T *xin = reinterpret_cast<const T *>(a.xin) + (size_t)blk * a.n_in * N;
    const int slot = (a.t + blk + ms.delay) % a.ring;
    if (a.slot_zero != nullptr && i == 0) {
        a.slot_zero[(size_t)ms.stream * a.ring + slot] = 0;     // powersave: a mixed slot is never skipped
    }
    T acc = (T)0;
    for (int j = 0; j < ms.n_inputs; j++) {
        const MixTerm tm = a.terms[ms.first + j];
        const T v = mul_rn(xin[(size_t)tm.index * N + i], (T)tm.scale);
        acc = j == 0 ? v : add_rn(acc, v);
    }
    reinterpret_cast<T *>(a.fdl)[((size_t)ms.stream * a.ring + slot) * N + i] = acc;
}

// ======================================================================================================
// k_mac -- variant 0: direct streaming loads
// ======================================================================================================

template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; };
template <> struct Vec16<double> { typedef double2 type; };

// streaming loads: read-only path, do not allocate in L1 (every operand byte is used once)
template <typename V>
__device__ __forceinline__ V ldg_stream(const V *p)
{
    if (sizeof(V) == 16) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                     : "l"(p));
        return *reinterpret_cast<V *>(&r);
    } else {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return *reinterpret_cast<V *>(&r);
    }
}

template <typename T, int W>
struct __align__(sizeof(T) * W) Lanes {
    T v[W];
};
template <typename T, int W, typename V>
__device__ __forceinline__ Lanes<T, W> as_lanes(const V &x)
{
    return *reinterpret_cast<const Lanes<T, W> *>(&x);
}

template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) k_mac(MacArgs a, int N)
{
    constexpr int W = 16 / (int)sizeof(T);
    typedef typename Vec16<T>::type V;
    const int M = N >> 1;
    const int vecs = M / W;     // 16-byte vectors per half spectrum
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long)a.n_jobs * vecs) {
        return;
    }
    const int job = (int)(g / vecs), v = (int)(g - (long)job * vecs);
    const MacJob jb = a.jobs[job];
    const int z = (int)blockIdx.y + a.z_first;
    const int P = a.ring;       // ring slots per stream
    T *out = reinterpret_cast<T *>(a.Y) + ((size_t)z * a.n_slots + jb.out) * N + (size_t)v * W;
    const T *X = reinterpret_cast<const T *>(a.fdl) + (size_t)jb.stream * P * N + (size_t)v * W;
    const int slot0 = a.t;
    // powersave (bfrun.c:1737-1754: `if (!cbuf_zero[n][j] || !powersave)`): partitions whose delay-line slot holds zeros
    // are skipped -- neither the slot nor its coefficient block is read
    const uint8_t *sz = a.slot_zero != nullptr ? a.slot_zero + (size_t)jb.stream * P : nullptr;

    Lanes<T, W> are, aim;
#pragma unroll
    for (int l = 0; l < W; l++) {
        are.v[l] = (T)0;
        aim.v[l] = (T)0;
    }

    if (jb.hbase < 0) {
        // coeff = -1: unit pulse in the shifted-coefficient convention = (+1/N, -1/N, ...) per bin
        if (z == 0) {
            const T fr = (T)(1.0 / (T)N);
            const Lanes<T, W> xr = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(X + (size_t)slot0 * N)));
            const Lanes<T, W> xi = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(X + (size_t)slot0 * N + M)));
#pragma unroll
            for (int l = 0; l < W; l++) {
                const T s = (l & 1) ? -fr : fr;
                are.v[l] = mul_rn(xr.v[l], s);
                aim.v[l] = mul_rn(xi.v[l], s);
            }
        }
    } else {
        const int chunk = (jb.n_parts + a.split - 1) / a.split;
        int i = z * chunk;
        int i1 = min(jb.n_parts, i + chunk);
        if (a.head > 0) {
            // uneven two-way split: partial 0 = the first `head` partitions, partial 1 = the others
            const int h = min(a.head, jb.n_parts);
            i = z == 0 ? 0 : h;
            i1 = z == 0 ? h : jb.n_parts;
        }
        const T *H = reinterpret_cast<const T *>(a.H) + (size_t)jb.hbase * N + (size_t)v * W;
        T dc = (T)0, ny = (T)0;
        if (i < i1 && sz != nullptr && sz[(slot0 - i) + ((slot0 - i) < 0 ? P : 0)]) {
            i++;        // a zero first partition: the sum starts from zero (memset of ocbuf, bfrun.c:1741-1744)
        } else if (i < i1) {
            // convolver_convolve: plain product for the first partition of the range
            int slot = slot0 - i;
            slot += (slot < 0) ? P : 0;
            const T *xp = X + (size_t)slot * N;
            const T *hp = H + (size_t)i * N;
            const Lanes<T, W> br = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(xp)));
            const Lanes<T, W> bi = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(xp + M)));
            const Lanes<T, W> cr = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(hp)));
            const Lanes<T, W> ci = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(hp + M)));
#pragma unroll
            for (int l = 0; l < W; l++) {
                cprod<T>(br.v[l], bi.v[l], cr.v[l], ci.v[l], are.v[l], aim.v[l]);
            }
            dc = mul_rn(br.v[0], cr.v[0]);
            ny = mul_rn(bi.v[0], ci.v[0]);
            i++;
        }
        // convolver_convolve_add: partitions ascending, accumulated in the real type -- the
        // reference's summation order; UNROLL partitions of loads are in flight per thread
        for (; i < i1; i += UNROLL) {
            V xr[UNROLL], xi[UNROLL], hr[UNROLL], hi[UNROLL];
            bool live[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                int slot = slot0 - (i + u);
                slot += (slot < 0) ? P : 0;
                live[u] = i + u < i1 && !(sz != nullptr && sz[slot]);
                if (live[u]) {
                    const T *xp = X + (size_t)slot * N;
                    const T *hp = H + (size_t)(i + u) * N;
                    xr[u] = ldg_stream(reinterpret_cast<const V *>(xp));
                    xi[u] = ldg_stream(reinterpret_cast<const V *>(xp + M));
                    hr[u] = ldg_stream(reinterpret_cast<const V *>(hp));
                    hi[u] = ldg_stream(reinterpret_cast<const V *>(hp + M));
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (live[u]) {
                    const Lanes<T, W> br = as_lanes<T, W>(xr[u]), bi = as_lanes<T, W>(xi[u]);
                    const Lanes<T, W> cr = as_lanes<T, W>(hr[u]), ci = as_lanes<T, W>(hi[u]);
#pragma unroll
                    for (int l = 0; l < W; l++) {
                        T re, im;
                        cprod<T>(br.v[l], bi.v[l], cr.v[l], ci.v[l], re, im);
                        are.v[l] = add_rn(are.v[l], re);
                        aim.v[l] = add_rn(aim.v[l], im);
                    }
                    dc = add_rn(dc, mul_rn(br.v[0], cr.v[0]));
                    ny = add_rn(ny, mul_rn(bi.v[0], ci.v[0]));
                }
            }
        }
        if (v == 0) {
            // slots [0] and [4] of the blocked layout: DC and Nyquist are real products
            are.v[0] = dc;
            aim.v[0] = ny;
        }
    }
    *reinterpret_cast<V *>(out) = *reinterpret_cast<V *>(&are);
    *reinterpret_cast<V *>(out + M) = *reinterpret_cast<V *>(&aim);
}

// ======================================================================================================
// k_split_reduce -- partial sums of a split partition range -> one spectrum per output, summed in range order
// ======================================================================================================

// Short spectra make this kernel latency: BASELINE config 4 (N = 512, 37 partials) is 128 vectors per output.  Blocks of
// 64 threads spread them over the machine and twelve partials are in flight per thread (was: 256-thread blocks = one
// block per output on 32 SMs, four in flight: 12.2 us for 2.4 MB, ncu profiles/r2_ncu_full_summary.json capture c4_mac).
template <typename T>
__global__ void __launch_bounds__(64) k_split_reduce(MacArgs a, int N)
{
    constexpr int W = 16 / (int)sizeof(T);
    constexpr int FL = 12;      // partials in flight
    typedef typename Vec16<T>::type V;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N / W) {
        return;
    }
    const MacJob jb = a.jobs[blockIdx.y];
    const int b = blockIdx.z;
    const size_t zstride = (size_t)a.batch * a.n_slots * N;
    T *y0 = reinterpret_cast<T *>(a.Y) + ((size_t)b * a.n_slots + jb.out) * N + (size_t)v * W;
    Lanes<T, W> acc = as_lanes<T, W>(*reinterpret_cast<const V *>(y0));
    int z = 1;
    for (; z + FL <= a.split; z += FL) {    // added in ascending order
        V p[FL];
#pragma unroll
        for (int u = 0; u < FL; u++) {
            p[u] = ldg_stream(reinterpret_cast<const V *>(y0 + (size_t)(z + u) * zstride));
        }
#pragma unroll
        for (int u = 0; u < FL; u++) {
            const Lanes<T, W> q = as_lanes<T, W>(p[u]);
#pragma unroll
            for (int l = 0; l < W; l++) {
                acc.v[l] = add_rn(acc.v[l], q.v[l]);
            }
        }
    }
    {
        V p[FL];
#pragma unroll
        for (int u = 0; u < FL; u++) {
            if (z + u < a.split) {
                p[u] = ldg_stream(reinterpret_cast<const V *>(y0 + (size_t)(z + u) * zstride));
            }
        }
#pragma unroll
        for (int u = 0; u < FL; u++) {
            if (z + u < a.split) {
                const Lanes<T, W> q = as_lanes<T, W>(p[u]);
#pragma unroll
                for (int l = 0; l < W; l++) {
                    acc.v[l] = add_rn(acc.v[l], q.v[l]);
                }
            }
        }
    }
    *reinterpret_cast<V *>(y0) = *reinterpret_cast<V *>(&acc);
}

// ======================================================================================================
// k_out_mix -- mixnscale(OUTPUT) for outputs fed by several filters
// ======================================================================================================

template <typename T>
__global__ void __launch_bounds__(256) k_out_mix(OutMixArgs a, int N)
{
    constexpr int W = 16 / (int)sizeof(T);
    typedef typename Vec16<T>::type V;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N / W) {
        return;
    }
    const int o = blockIdx.y, blk = blockIdx.z;
    const OutChan ch = a.mixes[o];
    if (ch.n <= 1) {
        return;         // a single filter: the inverse stage scales it itself
    }
    T *Y = reinterpret_cast<T *>(a.Y) + (size_t)blk * a.n_slots * N + (size_t)v * W;
    for (int pass = 0; pass < (ch.xf_first >= 0 ? 2 : 1); pass++) {
        const int first = pass == 0 ? ch.first : ch.xf_first;
        Lanes<T, W> acc;
        for (int j = 0; j < ch.n; j++) {
            const MixTerm tm = a.terms[first + j];
            const Lanes<T, W> y = as_lanes<T, W>(ldg_stream(reinterpret_cast<const V *>(Y + (size_t)tm.index * N)));
            const T sc = (T)tm.scale;
#pragma unroll
            for (int l = 0; l < W; l++) {
                const T p = mul_rn(y.v[l], sc);
                acc.v[l] = j == 0 ? p : add_rn(acc.v[l], p);
            }
        }
        *reinterpret_cast<V *>(Y + (size_t)(a.z_first + pass * a.n_out + o) * N) = *reinterpret_cast<V *>(&acc);
    }
}

// ======================================================================================================
// k_inverse
// ======================================================================================================

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_inverse(InverseArgs a, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int o = blockIdx.x, blk = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    const OutChan ch = a.chans[o];
    const T *Y = reinterpret_cast<const T *>(a.Y) + (size_t)blk * a.n_slots * N;
    const int zstride = a.batch * a.n_slots;    // Y slots between two partial sums of the split
    const int npass = ch.xf_first >= 0 ? 2 : 1;
    T keep[E];

    for (int pass = 0; pass < npass; pass++) {
        const int first = (npass == 2 && pass == 0) ? ch.xf_first : ch.first;
        load_and_inverse<T, E>(sre, sim, tw, M, tid, nt, [&](int i) {
            return mix_terms<T>(Y, a.terms, first, ch.n, zstride, a.split, N, i);
        });
        if (pass + 1 < npass) {
            // stash the "old" signal's valid half in registers, then reuse shared memory
#pragma unroll
            for (int b = 0; b < E / 2; b++) {
                const int j = tid + b * nt;
                if (j < M / 2) {
                    keep[2 * b] = sre[fft_pad(j)];
                    keep[2 * b + 1] = sim[fft_pad(j)];
                }
            }
            __syncthreads();
        }
    }

    // overlap-save: the first L samples are the valid output (fftw_convolver.c:498-501)
    const SampleFormat f = a.fmt[o];
    T *tdst = reinterpret_cast<T *>(a.out_time) + ((size_t)blk * a.n_out + o) * L;
    uint8_t *raw = a.raw_out + (size_t)blk * a.out_stride + f.byte_offset;
    const size_t stride = (size_t)f.sample_spacing * f.bytes;
    const double of_max = a.overflow[o].max;
    QuantStats st;
    quant_stats_init(st);
    uint64_t enc[E];
#pragma unroll
    for (int b = 0; b < E / 2; b++) {
        const int j = tid + b * nt;
        if (j < M / 2) {
            T y0 = sre[fft_pad(j)], y1 = sim[fft_pad(j)];
            if (npass == 2) {
                y0 = xfade<T>(keep[2 * b], y0, 2 * j, L);
                y1 = xfade<T>(keep[2 * b + 1], y1, 2 * j + 1, L);
            }
            *reinterpret_cast<typename Vec2<T>::type *>(tdst + 2 * j) = Vec2<T>::make(y0, y1);
            if (!ch.shared) {
                enc[2 * b] = encode_sample<T>(y0, f.bytes, f.sbytes, f.isfloat, f.swap, a.safety_limit, of_max, st);
                enc[2 * b + 1] = encode_sample<T>(y1, f.bytes, f.sbytes, f.isfloat, f.swap, a.safety_limit, of_max, st);
            }
        }
    }
    if (!ch.shared) {
#pragma unroll
        for (int b = 0; b < E / 2; b++) {
            const int j = tid + b * nt;
            if (j < M / 2) {
                store_raw_le(raw + (size_t)(2 * j) * stride, enc[2 * b], f.bytes);
                store_raw_le(raw + (size_t)(2 * j + 1) * stride, enc[2 * b + 1], f.bytes);
            }
        }
    }
    if (!ch.shared) {
        reduce_stats(st, &a.overflow[o], a.status, sre, tid, nt);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_quantise_shared(InverseArgs a, int L)
{
    const int o = blockIdx.x, blk = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    __shared__ QuantStats wstats[32];
    if (a.chans[o].shared != 1) {      // not shared, or dithered (k_dither's job)
        return;
    }
    const SampleFormat f = a.fmt[o];
    const T *src = reinterpret_cast<const T *>(a.out_time) + ((size_t)blk * a.n_out + o) * L;
    uint8_t *raw = a.raw_out + (size_t)blk * a.out_stride + f.byte_offset;
    const size_t stride = (size_t)f.sample_spacing * f.bytes;
    const double of_max = a.overflow[o].max;
    QuantStats st;
    quant_stats_init(st);
    for (int n = tid; n < L; n += nt) {
        real_to_raw<T>(src[n], raw + (size_t)n * stride, f.bytes, f.sbytes, f.isfloat, f.swap, a.safety_limit,
                       of_max, st);
    }
    reduce_stats(st, &a.overflow[o], a.status, wstats, tid, nt);
}

// ======================================================================================================
// k_eval -- filter -> filter chaining
// ======================================================================================================

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_eval(EvalArgs a, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    const EvalEntry en = a.entries[blockIdx.x];
    T *keep = reinterpret_cast<T *>(a.keep) + (size_t)(en.vin - a.n_in) * L;
    const int zstride = a.batch * a.n_slots;
    const int npass = en.xf_first >= 0 ? 2 : 1;
    T old[E];

    for (int blk = 0; blk < a.batch; blk++) {       // the blocks of a batch depend on each other through `keep`
        const T *Y = reinterpret_cast<const T *>(a.Y) + (size_t)blk * a.n_slots * N;
        for (int pass = 0; pass < npass; pass++) {
            const int first = (npass == 2 && pass == 0) ? en.xf_first : en.first;
            // mixnscale(OUTPUT) of the source filters' outputs (bfrun.c:1611-1615), then HC2R
            load_and_inverse<T, E>(sre, sim, tw, M, tid, nt, [&](int i) {
                return mix_terms<T>(Y, a.terms, first, en.n, zstride, a.split, N, i);
            });
            if (pass + 1 < npass) {
#pragma unroll
                for (int b = 0; b < E / 2; b++) {
                    const int j = tid + b * nt;
                    if (j < M / 2) {
                        old[2 * b] = sre[fft_pad(j)];
                        old[2 * b + 1] = sim[fft_pad(j)];
                    }
                }
                __syncthreads();
            }
        }
        // frame = [previous valid block | this valid block] (fftw_convolver.c:416-432); the valid block is the first
        // L samples, packed complex elements j < M/2
        T cur[E], prv[E];
#pragma unroll
        for (int b = 0; b < E / 2; b++) {
            const int j = tid + b * nt;
            if (j < M / 2) {
                cur[2 * b] = sre[fft_pad(j)];
                cur[2 * b + 1] = sim[fft_pad(j)];
                if (npass == 2) {
                    // the source crossfaded this block: its output is the blend (fftw_convolver.c:349-355)
                    cur[2 * b] = xfade<T>(old[2 * b], cur[2 * b], 2 * j, L);
                    cur[2 * b + 1] = xfade<T>(old[2 * b + 1], cur[2 * b + 1], 2 * j + 1, L);
                }
                prv[2 * b] = keep[2 * j];
                prv[2 * b + 1] = keep[2 * j + 1];
            }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < E / 2; b++) {
            const int j = tid + b * nt;
            if (j < M / 2) {
                sre[fft_pad(j)] = prv[2 * b];
                sim[fft_pad(j)] = prv[2 * b + 1];
                sre[fft_pad(j + M / 2)] = cur[2 * b];
                sim[fft_pad(j + M / 2)] = cur[2 * b + 1];
                keep[2 * j] = cur[2 * b];
                keep[2 * j + 1] = cur[2 * b + 1];
            }
        }
        __syncthreads();
        T *dst = reinterpret_cast<T *>(a.xin) + ((size_t)blk * a.n_vin + en.vin) * N;
        forward_and_emit<T, E>(sre, sim, tw, M, tid, nt, [&](int k, T re, T im) {
            dst[k] = re;
            dst[M + k] = im;
        });
        __syncthreads();
    }
}

// ======================================================================================================
// k_dither -- real2raw with HP-TPDF dither and error feedback
// ======================================================================================================

template <typename T>
__global__ void __launch_bounds__(32) k_dither(InverseArgs a, DitherArgs d, int L)
{
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= d.n_dither) {
        return;
    }
    DitherChan st = d.chans[j];
    const int o = st.out;
    const SampleFormat f = a.fmt[o];
    const T *map = reinterpret_cast<const T *>(d.randmap) + 256;
    const size_t stride = (size_t)f.sample_spacing * f.bytes;
    const int bits_n = f.sbytes << 3;
    const int32_t imin = (int32_t)(-((uint64_t)1 << (bits_n - 1)));
    const int32_t imax = (int32_t)(((uint64_t)1 << (bits_n - 1)) - 1);
    const T rmin = (T)imin, rmax = (T)imax;
    // this lane is the only writer of its channel's counters: the reference's sequential bookkeeping, literally
    Overflow of = a.overflow[o];
    unsigned int status = 0;
    T e0 = (T)st.e0, e1 = (T)st.e1;

    for (int blk = 0; blk < a.batch; blk++) {
        // dither_preloop_real2int_hp_tpdf, dither.h:28-38.  On a wrap the reference copies the last used table byte
        // into slot 0 so that the differenced sequence continues; here the table is read-only and the byte travels
        // in a register.
        int8_t prev;
        if (st.randtab_ptr + L >= d.randtab_size) {
            prev = d.randtab[st.randtab_ptr - 1];
            st.randtab_ptr = 1;
        } else {
            prev = d.randtab[st.randtab_ptr - 1];
        }
        const int8_t *tab = d.randtab + st.randtab_ptr;
        st.randtab_ptr += L;
        const T *src = reinterpret_cast<const T *>(a.out_time) + ((size_t)blk * a.n_out + o) * L;
        uint8_t *raw = a.raw_out + (size_t)blk * a.out_stride + f.byte_offset;
        for (int n = 0; n < L; n++) {
            T real_sample = src[n];
            // real2raw.h:24-42
            if (!isfinite((double)real_sample)) {
                status |= BF_STATUS_NONFINITE;
            }
            if (a.safety_limit != 0.0 && ((double)real_sample < -a.safety_limit * of.max ||
                                          (double)real_sample > a.safety_limit * of.max)) {
                status |= BF_STATUS_SAFETY;
            }
            // dither_funs.h:20-33: error feedback {1, -1}, then dither + the 0.5 offset from the map
            real_sample = add_rn(real_sample, sub_rn(e0, e1));
            e1 = e0;
            const int8_t cur = tab[n];
            const T dithered = add_rn(real_sample, map[(int)cur - (int)prev]);
            prev = cur;
            int32_t sample;
            if (dithered < (T)0) {
                if (dithered <= rmin) {
                    sample = imin;
                    of.n_overflows++;
                    if ((double)real_sample < -of.largest) {
                        of.largest = (double)-dithered;
                    }
                } else {
                    sample = (int32_t)dithered;
                    sample--;
                    if (sample < -of.intlargest) {
                        of.intlargest = -sample;
                    }
                }
            } else {
                if (dithered > rmax) {
                    sample = imax;
                    of.n_overflows++;
                    if ((double)real_sample > of.largest) {
                        of.largest = (double)dithered;
                    }
                } else {
                    sample = (int32_t)dithered;
                    if (sample > of.intlargest) {
                        of.intlargest = sample;
                    }
                }
            }
            e0 = sub_rn(real_sample, (T)sample);
            uint64_t bits = (uint64_t)(uint32_t)sample;
            if (f.bytes < 4) {
                bits &= ((uint64_t)1 << (8 * f.bytes)) - 1;
            }
            if (f.swap && f.bytes > 1) {
                bits = swap_bytes(bits, f.bytes);
            }
            store_raw_le(raw + (size_t)n * stride, bits, f.bytes);
        }
    }
    st.e0 = (double)e0;
    st.e1 = (double)e1;
    d.chans[j] = st;
    a.overflow[o].n_overflows = of.n_overflows;
    a.overflow[o].intlargest = of.intlargest;
    a.overflow[o].largest = of.largest;
    if (status != 0) {
        atomicOr(a.status, status);
    }
}

// ======================================================================================================
// k_subdelay, k_virt_mix -- sub-sample delay and virtual -> physical output mixing on the engine path
// ======================================================================================================

template <typename T>
__global__ void __launch_bounds__(256) k_subdelay(SubdelayArgs a)
{
    constexpr int CH = 1024;            // samples per chunk
    __shared__ T tile[BF_SUBDELAY_MAX_TAPS - 1 + CH];
    __shared__ T taps[BF_SUBDELAY_MAX_TAPS];
    const SubdelayChan sc = a.chans[blockIdx.x];
    const int H = sc.n_taps - 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    T *hist = reinterpret_cast<T *>(a.hist) + (size_t)sc.ch * (BF_SUBDELAY_MAX_TAPS - 1);     // kept per channel across filter changes
    for (int k = tid; k < sc.n_taps; k += nt) {
        taps[k] = reinterpret_cast<const T *>(a.taps)[sc.tap_first + k];
    }
    for (int k = tid; k < H; k += nt) {
        tile[k] = hist[k];              // the last H inputs before this launch, oldest first
    }
    __syncthreads();
    for (int blk = 0; blk < a.batch; blk++) {
        T *x = reinterpret_cast<T *>(a.data) + ((size_t)blk * a.n_ch + sc.ch) * a.L;
        for (int n0 = 0; n0 < a.L; n0 += CH) {
            const int cn = min(CH, a.L - n0);
            for (int n = tid; n < cn; n += nt) {
                tile[H + n] = x[n0 + n];
            }
            __syncthreads();
            for (int n = tid; n < cn; n += nt) {
                // y[n] = sum_k h[k] x[n - k]; tile[H + n - k] is input n - k of this chunk (or history)
                T acc = (T)0;
                for (int k = 0; k <= H; k++) {
                    acc += taps[k] * tile[H + n - k];
                }
                x[n0 + n] = acc;
            }
            __syncthreads();
            // the last H inputs of what the tile holds become the history of the next chunk (H may exceed cn)
            T keep[(BF_SUBDELAY_MAX_TAPS - 1 + 255) / 256];
            int q = 0;
            for (int k = tid; k < H; k += nt, q++) {
                keep[q] = tile[cn + k];
            }
            __syncthreads();
            q = 0;
            for (int k = tid; k < H; k += nt, q++) {
                tile[k] = keep[q];
            }
            __syncthreads();
        }
    }
    for (int k = tid; k < H; k += nt) {
        hist[k] = tile[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_virt_mix(VirtMixArgs a)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.L) {
        return;
    }
    const VirtGroup g = a.groups[blockIdx.y];
    T *base = reinterpret_cast<T *>(a.out_time) + (size_t)blockIdx.z * a.n_out * a.L + n;
    // mixbuf = first unmuted member; mixbuf += the others, in channel order (bfrun.c:1955-1981); all muted: zeros
    T acc = (T)0;
    bool filled = false;
    for (int j = 0; j < g.n; j++) {
        const int o = a.members[g.first + j];
        if (a.muted != nullptr && a.muted[o]) {
            continue;
        }
        const T v = base[(size_t)o * a.L];
        acc = filled ? add_rn(acc, v) : v;
        filled = true;
    }
    base[(size_t)a.members[g.first + g.n - 1] * a.L] = acc;
}

cudaError_t launch_subdelay(const FftPlan &plan, const SubdelayArgs &a, cudaStream_t s)
{
    if (a.n_chans == 0) return cudaSuccess;
    if (plan.realsize == 4) {
        k_subdelay<float><<<a.n_chans, 256, 0, s>>>(a);
    } else {
        k_subdelay<double><<<a.n_chans, 256, 0, s>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_virt_mix(const FftPlan &plan, const VirtMixArgs &a, cudaStream_t s)
{
    if (a.n_groups == 0) return cudaSuccess;
    dim3 grid((a.L + 255) / 256, a.n_groups, a.batch);
    if (plan.realsize == 4) {
        k_virt_mix<float><<<grid, 256, 0, s>>>(a);
    } else {
        k_virt_mix<double><<<grid, 256, 0, s>>>(a);
    }
    return cudaGetLastError();
}

// ======================================================================================================
// coefficient preprocessing and plain transforms
// ======================================================================================================

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_coeff_fft(const T *__restrict__ taps, T scale, T *__restrict__ H,
                                                       int hbase, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    // [0_L | scale * h] : the circular shift by L makes the FIRST half of the inverse transform valid
    for (int n = tid; n < L; n += nt) {
        const T cur = mul_rn(taps[(size_t)b * L + n], scale);
        const int j0 = fft_pad(n >> 1), j1 = fft_pad((L + n) >> 1);
        if (n & 1) {
            sim[j0] = (T)0;
            sim[j1] = cur;
        } else {
            sre[j0] = (T)0;
            sre[j1] = cur;
        }
    }
    __syncthreads();
    T *dst = H + (size_t)(hbase + b) * N;
    const T inv_n = (T)(1.0 / (double)N);
    forward_and_emit<T, E>(sre, sim, tw, M, tid, nt, [&](int k, T re, T im) {
        dst[k] = mul_rn(re, inv_n);
        dst[M + k] = mul_rn(im, inv_n);
    });
}

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_r2hc(const T *in, T *out, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    const T *src = in + (size_t)blockIdx.x * N;
    T *dst = out + (size_t)blockIdx.x * N;
    for (int j = tid; j < M; j += nt) {
        sre[fft_pad(j)] = src[2 * j];
        sim[fft_pad(j)] = src[2 * j + 1];
    }
    __syncthreads();
    forward_and_emit<T, E>(sre, sim, tw, M, tid, nt, [&](int k, T re, T im) {
        dst[k] = re;                    // hc[k] = Re X_k
        if (k == 0) {
            dst[M] = im;                // Nyquist
        } else {
            dst[N - k] = im;            // hc[N-k] = Im X_k
        }
    });
}

template <typename T, int E>
__global__ void __launch_bounds__(1024, 1) k_hc2r(const T *in, T *out, const T *__restrict__ tw, int L)
{
    const int M = L, N = 2 * L;
    const int tid = threadIdx.x, nt = blockDim.x;
    T *sre = smem_re<T>();
    T *sim = sre + (M + (M >> 5) + 1);
    const T *src = in + (size_t)blockIdx.x * N;
    T *dst = out + (size_t)blockIdx.x * N;
    load_and_inverse<T, E>(sre, sim, tw, M, tid, nt, [&](int i) { return src[planar_to_hc(i, M)]; });
    for (int j = tid; j < M; j += nt) {
        dst[2 * j] = sre[fft_pad(j)];
        dst[2 * j + 1] = sim[fft_pad(j)];
    }
}

// ======================================================================================================
// layout permutation and per-call kernels on the reference layouts
// ======================================================================================================

template <typename T>
__global__ void __launch_bounds__(256) k_permute(const T *__restrict__ src, T *__restrict__ dst, int N, int mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;    // planar index
    if (i >= N) {
        return;
    }
    const int M = N >> 1;
    const size_t base = (size_t)blockIdx.y * N;
    switch (mode) {
    case BLOCKED_TO_PLANAR: dst[base + i] = src[base + planar_to_blocked(i, M)]; break;
    case PLANAR_TO_BLOCKED: dst[base + planar_to_blocked(i, M)] = src[base + i]; break;
    case HC_TO_PLANAR: dst[base + i] = src[base + planar_to_hc(i, M)]; break;
    default: dst[base + planar_to_hc(i, M)] = src[base + i]; break;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_mixnscale(const void *const *in_ptrs, const double *scales, int n_bufs,
                                                      T *out, int N, int mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;    // planar index
    if (i >= N) {
        return;
    }
    const int M = N >> 1;
    const int hc = planar_to_hc(i, M), bl = planar_to_blocked(i, M);
    const int si = mode == 1 ? hc : bl, di = mode == 1 ? bl : hc;
    T acc = (T)0;
    for (int j = 0; j < n_bufs; j++) {
        const T v = mul_rn(reinterpret_cast<const T *>(in_ptrs[j])[si], (T)scales[j]);
        acc = j == 0 ? v : add_rn(acc, v);
    }
    out[di] = acc;
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_convolve(const T *b, const T *c, T *d, int N, int op)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;    // 8-real chunk
    if (q >= N / 8) {
        return;
    }
    T br[4], bi[4], cr[4], ci[4], dr[4], di[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        br[j] = b[8 * q + j];
        bi[j] = b[8 * q + 4 + j];
        cr[j] = c[8 * q + j];
        ci[j] = c[8 * q + 4 + j];
        dr[j] = op ? d[8 * q + j] : (T)0;
        di[j] = op ? d[8 * q + 4 + j] : (T)0;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        T re, im;
        cprod<T>(br[j], bi[j], cr[j], ci[j], re, im);
        if (q == 0 && j == 0) {
            re = mul_rn(br[0], cr[0]);      // DC
            im = mul_rn(bi[0], ci[0]);      // Nyquist
        }
        d[8 * q + j] = op ? add_rn(dr[j], re) : re;
        d[8 * q + 4 + j] = op ? add_rn(di[j], im) : im;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_dirac(const T *b, T *d, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) {
        return;
    }
    const T fr = (T)(1.0 / (T)N);
    d[i] = mul_rn(b[i], (i & 1) ? -fr : fr);
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_xfade_blend(const T *old_time, T *new_time, int L)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < L) {
        new_time[n] = xfade<T>(old_time[n], new_time[n], n, L);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_raw2real(const uint8_t *raw, SampleFormat f, T *dst, int L)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < L) {
        dst[n] = raw_to_real<T>(raw + f.byte_offset + (size_t)n * f.sample_spacing * f.bytes, f.bytes, f.isfloat,
                                f.swap);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_cv_real2raw(const T *src, uint8_t *raw, SampleFormat f, Overflow *of,
                                                     unsigned int *status, double safety_limit, int L)
{
    __shared__ QuantStats wstats[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double of_max = of->max;
    QuantStats st;
    quant_stats_init(st);
    for (int n = tid; n < L; n += nt) {
        real_to_raw<T>(src[n], raw + f.byte_offset + (size_t)n * f.sample_spacing * f.bytes, f.bytes, f.sbytes,
                       f.isfloat, f.swap, safety_limit, of_max, st);
    }
    reduce_stats(st, of, status, wstats, tid, nt);
}

// ======================================================================================================
// launchers
// ======================================================================================================

static size_t fft_smem_bytes(int M, int realsize)
{
    return (size_t)fft_smem_reals(M) * realsize;
}

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    return cudaSuccess;
}

static bool fft_force16()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("BFCUDA_FFT_E16");
        v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    return v == 1;
}

// dispatch on (realsize, points per thread): float uses 16 points per thread above M = 8192
#define BF_FFT_DISPATCH(plan, KERNEL, grid, stream, ...)                                                   \
    do {                                                                                                   \
        const int M_ = (plan).N / 2;                                                                       \
        const size_t smem_ = fft_smem_bytes(M_, (plan).realsize);                                          \
        cudaError_t e_;                                                                                    \
        if ((plan).realsize == 4) {                                                                        \
            typedef float T;                                                                               \
            if (M_ > 8192 || (fft_force16() && M_ >= 512)) {                                               \
                if ((e_ = allow_smem(KERNEL<float, 16>, smem_)) != cudaSuccess) return e_;                 \
                KERNEL<float, 16><<<grid, M_ / 16, smem_, stream>>>(__VA_ARGS__);                          \
            } else {                                                                                       \
                if ((e_ = allow_smem(KERNEL<float, 8>, smem_)) != cudaSuccess) return e_;                  \
                KERNEL<float, 8><<<grid, fft_threads(M_), smem_, stream>>>(__VA_ARGS__);                   \
            }                                                                                              \
        } else {                                                                                           \
            typedef double T;                                                                              \
            if ((e_ = allow_smem(KERNEL<double, 8>, smem_)) != cudaSuccess) return e_;                     \
            KERNEL<double, 8><<<grid, fft_threads(M_), smem_, stream>>>(__VA_ARGS__);                      \
        }                                                                                                  \
        return cudaGetLastError();                                                                         \
    } while (0)

cudaError_t launch_forward(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    if (a.n_in == 0) return cudaSuccess;
    if (plan.big_m1 != 0) return launch_forward_big(plan, a, s);
    if (plan.tw2 != nullptr) return launch_forward2(plan, a, s);
    BF_FFT_DISPATCH(plan, k_forward, dim3(a.n_in, a.batch), s, a, (const T *)plan.tw, plan.N / 2);
}

cudaError_t launch_stream_mix(const FftPlan &plan, const StreamMixArgs &a, cudaStream_t s)
{
    if (a.n_streams == 0) return cudaSuccess;
    dim3 grid((plan.N + 255) / 256, a.n_streams, a.batch);
    if (plan.realsize == 4) {
        k_stream_mix<float><<<grid, 256, 0, s>>>(a, plan.N);
    } else {
        k_stream_mix<double><<<grid, 256, 0, s>>>(a, plan.N);
    }
    return cudaGetLastError();
}

cudaError_t launch_mac_tma(const FftPlan &plan, const MacArgs &a, cudaStream_t s);   // bf_mac_tma.cu
cudaError_t launch_mac_batch2(const FftPlan &plan, const MacArgs &a, cudaStream_t s); // bf_mac_batch.cu

cudaError_t launch_mac(const FftPlan &plan, const MacArgs &a, cudaStream_t s)
{
    if (a.n_jobs == 0) return cudaSuccess;
    if (a.variant == 1 && a.batch == 1 && a.head == 0 && a.z_count == 0) {
        return launch_mac_tma(plan, a, s);
    }
    const int W = 16 / plan.realsize;
    const long threads = (long)a.n_jobs * (plan.N / 2 / W);
    dim3 grid((unsigned int)((threads + 255) / 256), a.z_count > 0 ? a.z_count : a.split);
    if (a.batch > 1) {
        if (a.head > 0 || a.z_count > 0) return cudaErrorInvalidValue;     // block-by-block schedule only
        return launch_mac_batch2(plan, a, s);      // bf_mac_batch.cu
    }
    // (the cp.async ring kernel was tried for single blocks as well -- BASELINE config 4, 28 partitions per thread after
    // the split: 36.8 us against 32.0 us for this kernel, profiles/r2_macsweep_groups_b1ring.txt)
    if (plan.realsize == 4) {
        g_last_func = (const void *)k_mac<float, 4>;
        k_mac<float, 4><<<grid, 256, 0, s>>>(a, plan.N);
    } else {
        g_last_func = (const void *)k_mac<double, 4>;
        k_mac<double, 4><<<grid, 256, 0, s>>>(a, plan.N);
    }
    return cudaGetLastError();
}

cudaError_t launch_split_reduce(const FftPlan &plan, const MacArgs &a, cudaStream_t s)
{
    if (a.n_jobs == 0 || a.split <= 1) return cudaSuccess;
    const int W = 16 / plan.realsize;
    dim3 grid((plan.N / W + 63) / 64, a.n_jobs, a.batch);
    if (plan.realsize == 4) {
        k_split_reduce<float><<<grid, 64, 0, s>>>(a, plan.N);
    } else {
        k_split_reduce<double><<<grid, 64, 0, s>>>(a, plan.N);
    }
    return cudaGetLastError();
}

cudaError_t launch_out_mix(const FftPlan &plan, const OutMixArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    const int W = 16 / plan.realsize;
    dim3 grid((plan.N / W + 255) / 256, a.n_out, a.batch);
    if (plan.realsize == 4) {
        k_out_mix<float><<<grid, 256, 0, s>>>(a, plan.N);
    } else {
        k_out_mix<double><<<grid, 256, 0, s>>>(a, plan.N);
    }
    return cudaGetLastError();
}

cudaError_t launch_inverse(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    if (plan.big_m1 != 0) return launch_inverse_big(plan, a, s);
    if (plan.tw2 != nullptr) return launch_inverse2(plan, a, s);
    BF_FFT_DISPATCH(plan, k_inverse, dim3(a.n_out, a.batch), s, a, (const T *)plan.tw, plan.N / 2);
}

cudaError_t launch_eval(const FftPlan &plan, const EvalArgs &a, cudaStream_t s)
{
    if (a.n_entries == 0) return cudaSuccess;
    if (plan.big_m1 != 0) return cudaErrorNotSupported;     // chained filters stop at one block's transform size
    BF_FFT_DISPATCH(plan, k_eval, a.n_entries, s, a, (const T *)plan.tw, plan.N / 2);
}

cudaError_t launch_dither(const FftPlan &plan, const InverseArgs &a, const DitherArgs &d, cudaStream_t s)
{
    if (d.n_dither == 0) return cudaSuccess;
    const int grid = (d.n_dither + 31) / 32;
    if (plan.realsize == 4) {
        k_dither<float><<<grid, 32, 0, s>>>(a, d, plan.N / 2);
    } else {
        k_dither<double><<<grid, 32, 0, s>>>(a, d, plan.N / 2);
    }
    return cudaGetLastError();
}

cudaError_t launch_quantise_shared(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    if (plan.realsize == 4) {
        k_quantise_shared<float><<<dim3(a.n_out, a.batch), 256, 0, s>>>(a, plan.N / 2);
    } else {
        k_quantise_shared<double><<<dim3(a.n_out, a.batch), 256, 0, s>>>(a, plan.N / 2);
    }
    return cudaGetLastError();
}

cudaError_t launch_coeff_fft(const FftPlan &plan, const void *taps, int n_blocks, double scale, void *H, int hbase,
                             cudaStream_t s)
{
    if (n_blocks == 0) return cudaSuccess;
    if (plan.big_m1 != 0) return launch_coeff_fft_big(plan, taps, n_blocks, scale, H, hbase, s);
    BF_FFT_DISPATCH(plan, k_coeff_fft, n_blocks, s, (const T *)taps, (T)scale, (T *)H, hbase, (const T *)plan.tw,
                    plan.N / 2);
}

cudaError_t launch_r2hc(const FftPlan &plan, const void *in, void *out, int batch, cudaStream_t s)
{
    BF_FFT_DISPATCH(plan, k_r2hc, batch, s, (const T *)in, (T *)out, (const T *)plan.tw, plan.N / 2);
}

cudaError_t launch_hc2r(const FftPlan &plan, const void *in, void *out, int batch, cudaStream_t s)
{
    BF_FFT_DISPATCH(plan, k_hc2r, batch, s, (const T *)in, (T *)out, (const T *)plan.tw, plan.N / 2);
}

// ---- convolver_td_* pieces (the reference's small ordered-layout convolver, fftw_convolver.c:682-782) -----------
template <typename T>
__global__ void __launch_bounds__(256) k_td_scale(T *buf, int n, T scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        buf[i] = mul_rn(buf[i], scale);         // fftw_convolver.c:722-732
    }
}

// b (*)= c on FFTW's half-complex order, hc[k] = Re X_k, hc[n-k] = Im X_k; DC and Nyquist are real products
// (convolve_inplace_ordered, fftw_convolver.c:737-763), products and sums separately rounded as compiled there
template <typename T>
__global__ void __launch_bounds__(256) k_td_mul(T *b, const T *__restrict__ c, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = n >> 1;
    if (k > half) {
        return;
    }
    if (k == 0 || k == half) {
        b[k] = mul_rn(b[k], c[k]);
        return;
    }
    const T br = b[k], bi = b[n - k], cr = c[k], ci = c[n - k];
    b[k] = add_rn(mul_rn(br, cr), -mul_rn(bi, ci));
    b[n - k] = add_rn(mul_rn(br, ci), mul_rn(bi, cr));
}

// n = 2 or 4: X_k = sum_j x_j e^{-2 pi i jk/n} written out (all roots are +-1, +-i); unnormalised both ways like the
// FFTW_R2HC / FFTW_HC2R plans (fftw_convolver.c:98-126)
template <typename T>
__global__ void k_td_small(const T *in, T *out, int n, int dir)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) {
        return;
    }
    if (n == 2) {
        const T a = in[0], b = in[1];
        out[0] = add_rn(a, b);
        out[1] = add_rn(a, -b);
    } else if (dir == 0) {
        const T x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
        const T s02 = add_rn(x0, x2), s13 = add_rn(x1, x3);
        out[0] = add_rn(s02, s13);              // X_0
        out[1] = add_rn(x0, -x2);               // Re X_1
        out[2] = add_rn(s02, -s13);             // X_2 (Nyquist)
        out[3] = add_rn(x3, -x1);               // Im X_1
    } else {
        const T X0 = in[0], R1 = in[1], X2 = in[2], I1 = in[3];
        const T s = add_rn(X0, X2), d = add_rn(X0, -X2);
        const T r2 = add_rn(R1, R1), i2 = add_rn(I1, I1);
        out[0] = add_rn(s, r2);
        out[1] = add_rn(d, -i2);
        out[2] = add_rn(s, -r2);
        out[3] = add_rn(d, i2);
    }
}

cudaError_t launch_td_scale(int realsize, void *buf, int n, double scale, cudaStream_t s)
{
    const int grid = (n + 255) / 256;
    if (realsize == 4) {
        k_td_scale<float><<<grid, 256, 0, s>>>((float *)buf, n, (float)scale);
    } else {
        k_td_scale<double><<<grid, 256, 0, s>>>((double *)buf, n, scale);
    }
    return cudaGetLastError();
}

cudaError_t launch_td_mul(int realsize, void *buf, const void *coeffs, int n, cudaStream_t s)
{
    const int grid = (n / 2 + 1 + 255) / 256;
    if (realsize == 4) {
        k_td_mul<float><<<grid, 256, 0, s>>>((float *)buf, (const float *)coeffs, n);
    } else {
        k_td_mul<double><<<grid, 256, 0, s>>>((double *)buf, (const double *)coeffs, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_td_small(int realsize, const void *in, void *out, int n, int dir, cudaStream_t s)
{
    if (n != 2 && n != 4) {
        return cudaErrorInvalidValue;
    }
    if (realsize == 4) {
        k_td_small<float><<<1, 32, 0, s>>>((const float *)in, (float *)out, n, dir);
    } else {
        k_td_small<double><<<1, 32, 0, s>>>((const double *)in, (double *)out, n, dir);
    }
    return cudaGetLastError();
}

cudaError_t launch_permute(const FftPlan &plan, const void *src, void *dst, int n_spectra, int mode, cudaStream_t s)
{
    if (n_spectra == 0) return cudaSuccess;
    dim3 grid((plan.N + 255) / 256, n_spectra);
    if (plan.realsize == 4) {
        k_permute<float><<<grid, 256, 0, s>>>((const float *)src, (float *)dst, plan.N, mode);
    } else {
        k_permute<double><<<grid, 256, 0, s>>>((const double *)src, (double *)dst, plan.N, mode);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_mixnscale(const FftPlan &plan, const void *const *in_ptrs, const double *scales, int n_bufs,
                                void *out, int mode, cudaStream_t s)
{
    const int grid = (plan.N + 255) / 256;
    if (plan.realsize == 4) {
        k_cv_mixnscale<float><<<grid, 256, 0, s>>>(in_ptrs, scales, n_bufs, (float *)out, plan.N, mode);
    } else {
        k_cv_mixnscale<double><<<grid, 256, 0, s>>>(in_ptrs, scales, n_bufs, (double *)out, plan.N, mode);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_convolve(const FftPlan &plan, const void *b, const void *c, void *d, int op, cudaStream_t s)
{
    const int grid = (plan.N / 8 + 255) / 256;
    if (plan.realsize == 4) {
        k_cv_convolve<float><<<grid, 256, 0, s>>>((const float *)b, (const float *)c, (float *)d, plan.N, op);
    } else {
        k_cv_convolve<double><<<grid, 256, 0, s>>>((const double *)b, (const double *)c, (double *)d, plan.N, op);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_dirac(const FftPlan &plan, const void *b, void *d, cudaStream_t s)
{
    const int grid = (plan.N + 255) / 256;
    if (plan.realsize == 4) {
        k_cv_dirac<float><<<grid, 256, 0, s>>>((const float *)b, (float *)d, plan.N);
    } else {
        k_cv_dirac<double><<<grid, 256, 0, s>>>((const double *)b, (double *)d, plan.N);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_xfade_blend(const FftPlan &plan, const void *old_time, void *new_time, cudaStream_t s)
{
    const int L = plan.N / 2;
    const int grid = (L + 255) / 256;
    if (plan.realsize == 4) {
        k_cv_xfade_blend<float><<<grid, 256, 0, s>>>((const float *)old_time, (float *)new_time, L);
    } else {
        k_cv_xfade_blend<double><<<grid, 256, 0, s>>>((const double *)old_time, (double *)new_time, L);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_raw2real(const FftPlan &plan, const uint8_t *raw, SampleFormat fmt, void *dst, cudaStream_t s)
{
    const int L = plan.N / 2;
    const int grid = (L + 255) / 256;
    if (plan.realsize == 4) {
        k_cv_raw2real<float><<<grid, 256, 0, s>>>(raw, fmt, (float *)dst, L);
    } else {
        k_cv_raw2real<double><<<grid, 256, 0, s>>>(raw, fmt, (double *)dst, L);
    }
    return cudaGetLastError();
}

cudaError_t launch_cv_real2raw(const FftPlan &plan, const void *src, uint8_t *raw, SampleFormat fmt,
                               Overflow *overflow, unsigned int *status, double safety_limit, cudaStream_t s)
{
    const int L = plan.N / 2;
    if (plan.realsize == 4) {
        k_cv_real2raw<float><<<1, 256, 0, s>>>((const float *)src, raw, fmt, overflow, status, safety_limit, L);
    } else {
        k_cv_real2raw<double><<<1, 256, 0, s>>>((const double *)src, raw, fmt, overflow, status, safety_limit, L);
    }
    return cudaGetLastError();
}

}  // namespace bf
