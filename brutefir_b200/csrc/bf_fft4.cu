// bf_fft4.cu -- partitions longer than one thread block can hold: the four-step FFT through global memory.
//
// The reference takes any power-of-two filter_length (/root/reference/fftw_convolver.c:784-808, bfconf.c:1495-1520) and
// ships bench3_config with `filter_length: 65536` (/root/reference/bench3_config:2).  One thread block holds an
// M = N/2 point complex transform up to M = 16384 (float) / 8192 (double) in shared memory; beyond that the M-point
// transform of the packed sequence z_i = x_2i + i x_2i+1 is factored M = M1 x M2 (M1 <= M2, both powers of two):
//
//   forward   k_big_cols : for every column c2 the M1-point FFT over z[c2 + M2 r], times W_M^(c2 k1)  -> T[k1][c2]
//             k_big_rows : for every row k1 the M2-point FFT over T[k1][.]                              -> Z[k1 + M1 k2]
//             k_big_split: the O(N) real split of Z, scaled and written to every destination (delay line / H / xin)
//   inverse   k_big_merge: output mix (mixnscale OUTPUT) + the real merge                               -> Z
//             k_big_cols / k_big_rows with conjugated roots; the rows kernel stores only the first L samples
//             (overlap-save, fftw_convolver.c:498-501) as reals; a crossfade runs the chain twice and blends.
//
// The sub-transforms are the generic shared-memory FFT of bf_fft.cuh (several side by side in one block, in lock step);
// column loads and transposed stores are arranged so that a warp touches runs of consecutive elements.  Same definition
// as everywhere else (FFTW's unnormalised R2HC / HC2R), same planar spectrum layout, float and double.  Byte roofline:
// three passes over the M complex points per transform (scratch T, scratch Z, spectrum) instead of one -- at these
// lengths a block is >= 0.68 s of audio, the stage is ~1 % of the block's MAC traffic either way.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "bf_kernels.h"
#include "bf_fft.cuh"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"

namespace bf {

template <typename T>
struct alignas(2 * sizeof(T)) cpair {
    T x, y;
};

// ---- where the packed sequence z comes from (column pass) --------------------------------------------------------
// FRAME: z_i = (lo[2i], lo[2i+1]) for i < M/2, (hi[2(i - M/2)], ...) above, times scale; lo == NULL reads zeros
//        (forward: lo = previous block, hi = this block; coefficients: lo = NULL, hi = taps, fftw_convolver.c:535-560)
// CPX:   z_i = src[i] (inverse: the merged spectrum)
template <typename T>
struct BigColArgs {
    const T *lo, *hi;           // FRAME: item 0's halves
    size_t lo_stride, hi_stride;    // reals between two items
    const cpair<T> *src;        // CPX: item 0
    cpair<T> *dst;              // scratch T, item 0
    const T *tw_sub;            // W_{2 M1}^j, j < M1
    const T *tw_main;           // W_N^j, j < M  (N = 2M)
    T scale;
    int M, M1, M2;
    int frame;                  // 1 = FRAME source
};

template <typename T, bool INV>
__global__ void __launch_bounds__(1024) k_big_cols(BigColArgs<T> a)
{
    const int M1 = a.M1, M2 = a.M2, M = a.M;
    const int nt = M1 / 8 > 0 ? M1 / 8 : 1;         // threads per transform
    const int cpb = (int)blockDim.x / nt;           // columns per block
    const int item = blockIdx.y;
    const int c0 = blockIdx.x * cpb;
    T *base = smem_re<T>();
    const int region = fft_smem_reals(M1);
    const int half = M1 + (M1 >> 5) + 1;
    // load: consecutive threads take consecutive columns of one row (runs of cpb elements in memory)
    {
        const int col = threadIdx.x % cpb, r0 = threadIdx.x / cpb, rstep = (int)blockDim.x / cpb;
        T *sre = base + (size_t)col * region, *sim = sre + half;
        for (int r = r0; r < M1; r += rstep) {
            const int i = c0 + col + M2 * r;
            T zr, zi;
            if (a.frame) {
                const int hm = M >> 1;
                if (i < hm) {
                    if (a.lo != nullptr) {
                        const T *p = a.lo + (size_t)item * a.lo_stride + 2 * (size_t)i;
                        zr = p[0];
                        zi = p[1];
                    } else {
                        zr = (T)0;
                        zi = (T)0;
                    }
                } else {
                    const T *p = a.hi + (size_t)item * a.hi_stride + 2 * (size_t)(i - hm);
                    zr = mul_rn(p[0], a.scale);
                    zi = mul_rn(p[1], a.scale);
                }
            } else {
                const cpair<T> z = a.src[(size_t)item * M + i];
                zr = z.x;
                zi = z.y;
            }
            sre[fft_pad(r)] = zr;
            sim[fft_pad(r)] = zi;
        }
    }
    __syncthreads();
    {
        const int sub = threadIdx.x / nt, tid = threadIdx.x % nt;
        T *sre = base + (size_t)sub * region, *sim = sre + half;
        fft_complex_inplace<T, 8, INV>(sre, sim, a.tw_sub, M1, tid, nt, BlockSync());
    }
    // twiddle W_M^(c2 k1) (conjugated for the inverse) and store T[k1][c2]
    {
        const int col = threadIdx.x % cpb, r0 = threadIdx.x / cpb, rstep = (int)blockDim.x / cpb;
        const T *sre = base + (size_t)col * region, *sim = sre + half;
        const int c2 = c0 + col;
        cpair<T> *dst = a.dst + (size_t)item * M;
        for (int k1 = r0; k1 < M1; k1 += rstep) {
            const long p = ((long)c2 * k1) % M;         // exponent of W_M
            T wr, wi;
            fft_twiddle<T>(a.tw_main, M, (int)(2 * p), INV, wr, wi);     // W_M^p = W_N^(2p)
            const T xr = sre[fft_pad(k1)], xi = sim[fft_pad(k1)];
            cpair<T> o;
            o.x = xr * wr - xi * wi;
            o.y = xr * wi + xi * wr;
            dst[(size_t)k1 * M2 + c2] = o;
        }
    }
}

template <typename T>
struct BigRowArgs {
    const cpair<T> *src;        // scratch T, item 0
    cpair<T> *dst;              // item 0; element k = k1 + M1 k2 for k < kmax
    size_t dst_stride;          // complex elements between two items
    const T *tw_sub;            // W_{2 M2}^j, j < M2
    int M, M1, M2;
    int kmax;
};

template <typename T, bool INV>
__global__ void __launch_bounds__(1024) k_big_rows(BigRowArgs<T> a)
{
    const int M1 = a.M1, M2 = a.M2, M = a.M;
    const int nt = M2 / 8 > 0 ? M2 / 8 : 1;
    const int rpb = (int)blockDim.x / nt;           // rows per block
    const int item = blockIdx.y;
    const int k10 = blockIdx.x * rpb;
    T *base = smem_re<T>();
    const int region = fft_smem_reals(M2);
    const int half = M2 + (M2 >> 5) + 1;
    const int sub = threadIdx.x / nt, tid = threadIdx.x % nt;
    {
        T *sre = base + (size_t)sub * region, *sim = sre + half;
        const cpair<T> *src = a.src + (size_t)item * M + (size_t)(k10 + sub) * M2;
        for (int c = tid; c < M2; c += nt) {
            const cpair<T> z = src[c];
            sre[fft_pad(c)] = z.x;
            sim[fft_pad(c)] = z.y;
        }
        __syncthreads();
        fft_complex_inplace<T, 8, INV>(sre, sim, a.tw_sub, M2, tid, nt, BlockSync());
    }
    // transposed store: consecutive threads write consecutive k1 of one k2
    {
        const int row = threadIdx.x % rpb, q0 = threadIdx.x / rpb, qstep = (int)blockDim.x / rpb;
        const T *sre = base + (size_t)row * region, *sim = sre + half;
        cpair<T> *dst = a.dst + (size_t)item * a.dst_stride;
        for (int k2 = q0; k2 < M2; k2 += qstep) {
            const long k = (long)(k10 + row) + (long)M1 * k2;
            if (k < a.kmax) {
                cpair<T> o;
                o.x = sre[fft_pad(k2)];
                o.y = sim[fft_pad(k2)];
                dst[k] = o;
            }
        }
    }
}

// ---- forward: real split + emit ------------------------------------------------------------------------------------
template <typename T>
struct BigSplitArgs {
    const cpair<T> *Z;          // item 0
    const T *tw_main;
    int M;
    // engine mode (coeff_dst == NULL): item = blk * n_in + c, destinations from the forward tables
    ForwardArgs fa;
    int item0;                  // first item of this chunk (for c / blk)
    // coefficient mode: item b -> coeff_dst + (hbase + b) * N, scaled by coeff_scale (1/N)
    T *coeff_dst;
    int hbase;
    T coeff_scale;
};

template <typename T>
__global__ void __launch_bounds__(256) k_big_split(BigSplitArgs<T> a)
{
    const int M = a.M, N = 2 * M;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M / 2) {
        return;
    }
    const int item = blockIdx.y;
    const cpair<T> *Z = a.Z + (size_t)item * M;
    T x0r, x0i, x1r = (T)0, x1i = (T)0;
    bool two = false;
    if (k == 0) {
        const cpair<T> z = Z[0];
        x0r = z.x + z.y;        // X_0
        x0i = z.x - z.y;        // X_M rides as the imaginary part of bin 0
    } else {
        T wr, wi;
        fft_twiddle<T>(a.tw_main, M, k, false, wr, wi);
        const cpair<T> zk = Z[k], zm = Z[M - k];
        fft_split_pair<T>(zk.x, zk.y, zm.x, zm.y, wr, wi, x0r, x0i, x1r, x1i);
        two = k != M - k;
    }
    if (a.coeff_dst != nullptr) {
        T *dst = a.coeff_dst + (size_t)(a.hbase + item) * N;
        dst[k] = mul_rn(x0r, a.coeff_scale);
        dst[M + k] = mul_rn(x0i, a.coeff_scale);
        if (two) {
            dst[M - k] = mul_rn(x1r, a.coeff_scale);
            dst[2 * M - k] = mul_rn(x1i, a.coeff_scale);
        }
        return;
    }
    const ForwardArgs &f = a.fa;
    const int gi = a.item0 + item;
    const int c = gi % f.n_in, blk = gi / f.n_in;
    const bool silent = frame_silent(f, c, blk);        // powersave
    if (silent) {
        x0r = x0i = x1r = x1i = (T)0;
    }
    if (f.slot_zero != nullptr && k == 0) {
        for (int d = f.dest_first[c]; d < f.dest_first[c + 1]; d++) {
            const FwdDest ds = f.dests[d];
            f.slot_zero[(size_t)ds.stream * f.ring + (f.t + blk + ds.delay) % f.ring] = silent ? 1 : 0;
        }
    }
    if (f.xin != nullptr && f.need_xin[c]) {
        T *xin = reinterpret_cast<T *>(f.xin) + ((size_t)blk * f.n_vin + c) * N;
        xin[k] = x0r;
        xin[M + k] = x0i;
        if (two) {
            xin[M - k] = x1r;
            xin[2 * M - k] = x1i;
        }
    }
    T *fdl = reinterpret_cast<T *>(f.fdl);
    const int t = f.t + blk;
    for (int d = f.dest_first[c]; d < f.dest_first[c + 1]; d++) {
        const FwdDest ds = f.dests[d];
        T *dst = fdl + ((size_t)ds.stream * f.ring + (t + ds.delay) % f.ring) * N;
        const T sc = silent ? (T)0 : (T)ds.scale;      // exact zeros, never -0 or NaN * 0
        dst[k] = mul_rn(x0r, sc);
        dst[M + k] = mul_rn(x0i, sc);
        if (two) {
            dst[M - k] = mul_rn(x1r, sc);
            dst[2 * M - k] = mul_rn(x1i, sc);
        }
    }
}

// ---- inverse: output mix + real merge --------------------------------------------------------------------------------
template <typename T>
struct BigMergeArgs {
    cpair<T> *Z;                // item 0
    const T *tw_main;
    int M;
    InverseArgs ia;
    int item0;
    int old_pass;               // 1: take the "old coefficient" terms where an output crossfades
};

template <typename T>
__global__ void __launch_bounds__(256) k_big_merge(BigMergeArgs<T> a)
{
    const int M = a.M, N = 2 * M;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > M / 2) {
        return;
    }
    const int item = blockIdx.y;
    const InverseArgs &ia = a.ia;
    const int gi = a.item0 + item;
    const int o = gi % ia.n_out, blk = gi / ia.n_out;
    const OutChan ch = ia.chans[o];
    const int first = (a.old_pass && ch.xf_first >= 0) ? ch.xf_first : ch.first;
    const T *Y = reinterpret_cast<const T *>(ia.Y) + (size_t)blk * ia.n_slots * N;
    const int zstride = ia.batch * ia.n_slots;
    auto load = [&](int i) { return mix_terms<T>(Y, ia.terms, first, ch.n, zstride, ia.split, N, i); };
    cpair<T> *Z = a.Z + (size_t)item * M;
    if (k == 0) {
        const T x0 = load(0), xm = load(M);
        cpair<T> z;
        z.x = x0 + xm;
        z.y = x0 - xm;
        Z[0] = z;
        return;
    }
    T wr, wi;
    fft_twiddle<T>(a.tw_main, M, k, false, wr, wi);
    const T xkr = load(k), xki = load(M + k), xmr = load(M - k), xmi = load(2 * M - k);
    cpair<T> zk, zm;
    fft_merge_pair<T>(xkr, xki, xmr, xmi, wr, wi, zk.x, zk.y, zm.x, zm.y);
    Z[k] = zk;
    if (k != M - k) {
        Z[M - k] = zm;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_big_blend(InverseArgs ia, const T *old_time, int item0, int L)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= L) {
        return;
    }
    const int gi = item0 + blockIdx.y;
    const int o = gi % ia.n_out;
    if (ia.chans[o].xf_first < 0) {
        return;
    }
    T *nw = reinterpret_cast<T *>(ia.out_time) + (size_t)gi * L;
    nw[n] = xfade<T>(old_time[(size_t)blockIdx.y * L + n], nw[n], n, L);
}

// ======================================================================================================
// plan + launchers
// ======================================================================================================

bool fft_big_supported(int N, int realsize)
{
    if (N < 8 || (N & (N - 1)) != 0) {
        return false;
    }
    const int M = N / 2;
    return M >= (realsize == 4 ? 32768 : 16384) && M <= (1 << 22);
}

template <typename T>
static cudaError_t make_sub_table(void **out, int Mx)
{
    std::vector<T> h(2 * (size_t)Mx);
    const int Nx = 2 * Mx;
    for (int j = 0; j < Mx; j++) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)Nx;
        h[2 * (size_t)j] = (T)cosl(ang);
        h[2 * (size_t)j + 1] = (T)sinl(ang);
    }
    if (Nx >= 8) {
        h[2 * (size_t)(Nx / 4)] = (T)0;
        h[2 * (size_t)(Nx / 4) + 1] = (T)-1;
    }
    cudaError_t err = cudaMalloc(out, h.size() * sizeof(T));
    if (err != cudaSuccess) {
        return err;
    }
    return cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

cudaError_t fft_big_plan_create(FftPlan *plan, int max_items)
{
    plan->big_m1 = plan->big_m2 = 0;
    plan->big_tw1 = plan->big_tw2 = plan->big_scr0 = plan->big_scr1 = plan->big_old = nullptr;
    plan->big_items = 0;
    if (!fft_big_supported(plan->N, plan->realsize)) {
        return cudaSuccess;
    }
    const int M = plan->N / 2;
    int lg = 0;
    while ((1 << lg) < M) lg++;
    plan->big_m1 = 1 << (lg / 2);
    plan->big_m2 = M / plan->big_m1;
    cudaError_t err;
    if (plan->realsize == 4) {
        if ((err = make_sub_table<float>(&plan->big_tw1, plan->big_m1)) != cudaSuccess) return err;
        if ((err = make_sub_table<float>(&plan->big_tw2, plan->big_m2)) != cudaSuccess) return err;
    } else {
        if ((err = make_sub_table<double>(&plan->big_tw1, plan->big_m1)) != cudaSuccess) return err;
        if ((err = make_sub_table<double>(&plan->big_tw2, plan->big_m2)) != cudaSuccess) return err;
    }
    // two scratch arrays of M complex points per transform in flight; cap them at 256 MiB each and walk larger
    // launches in chunks
    const size_t per_item = (size_t)M * 2 * plan->realsize;
    size_t items = (size_t)(max_items < 1 ? 1 : max_items);
    const size_t cap = ((size_t)256 << 20) / per_item;
    if (items > cap) items = cap < 1 ? 1 : cap;
    plan->big_items = (int)items;
    if ((err = cudaMalloc(&plan->big_scr0, items * per_item)) != cudaSuccess) return err;
    if ((err = cudaMalloc(&plan->big_scr1, items * per_item)) != cudaSuccess) return err;
    if ((err = cudaMalloc(&plan->big_old, items * (per_item / 2))) != cudaSuccess) return err;     // L reals per item
    return cudaSuccess;
}

void fft_big_plan_destroy(FftPlan *plan)
{
    for (void **p : { &plan->big_tw1, &plan->big_tw2, &plan->big_scr0, &plan->big_scr1, &plan->big_old }) {
        if (*p != nullptr) {
            cudaFree(*p);
            *p = nullptr;
        }
    }
    plan->big_m1 = plan->big_m2 = plan->big_items = 0;
}

static void sub_launch_shape(int Mx, int realsize, int *threads, int *per_block, size_t *smem)
{
    const int nt = Mx / 8 > 0 ? Mx / 8 : 1;
    int pb = 1024 / nt;
    if (pb > 8) pb = 8;
    if (pb < 1) pb = 1;
    while (pb > 1 && (size_t)pb * fft_smem_reals(Mx) * realsize > 160 * 1024) {
        pb >>= 1;
    }
    *threads = pb * nt;
    *per_block = pb;
    *smem = (size_t)pb * fft_smem_reals(Mx) * realsize;
}

template <typename K>
static cudaError_t allow_big_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    return cudaSuccess;
}

// the M-point complex transform of `items` sequences: source (frame or complex) -> ... -> rows kernel's destination
template <typename T, bool INV>
static cudaError_t big_complex(const FftPlan &plan, BigColArgs<T> ca, BigRowArgs<T> ra, int items, cudaStream_t s)
{
    const int M = plan.N / 2;
    int th, pb;
    size_t smem;
    cudaError_t err;
    ca.M = ra.M = M;
    ca.M1 = ra.M1 = plan.big_m1;
    ca.M2 = ra.M2 = plan.big_m2;
    ca.tw_sub = reinterpret_cast<const T *>(plan.big_tw1);
    ca.tw_main = reinterpret_cast<const T *>(plan.tw);
    ca.dst = reinterpret_cast<cpair<T> *>(plan.big_scr0);
    ra.src = reinterpret_cast<const cpair<T> *>(plan.big_scr0);
    ra.tw_sub = reinterpret_cast<const T *>(plan.big_tw2);
    sub_launch_shape(plan.big_m1, plan.realsize, &th, &pb, &smem);
    if ((err = allow_big_smem(k_big_cols<T, INV>, smem)) != cudaSuccess) return err;
    k_big_cols<T, INV><<<dim3(plan.big_m2 / pb, items), th, smem, s>>>(ca);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    sub_launch_shape(plan.big_m2, plan.realsize, &th, &pb, &smem);
    if ((err = allow_big_smem(k_big_rows<T, INV>, smem)) != cudaSuccess) return err;
    k_big_rows<T, INV><<<dim3(plan.big_m1 / pb, items), th, smem, s>>>(ra);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t forward_big_t(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    const int M = plan.N / 2, L = M;
    const int total = a.n_in * a.batch;
    for (int i0 = 0; i0 < total; i0 += plan.big_items) {
        const int items = total - i0 < plan.big_items ? total - i0 : plan.big_items;
        // item gi = blk * n_in + c: this block at xt_cur + gi * L; the block before it one batch row earlier, or
        // xt_prev for the first block of the batch.  Rows of one chunk share a single (lo, hi) stride only within a
        // block row, so chunks are cut at block rows when the batch has several.
        for (int j0 = 0; j0 < items;) {
            const int gi = i0 + j0;
            const int blk = gi / a.n_in, c = gi % a.n_in;
            int run = a.n_in - c;
            if (run > items - j0) run = items - j0;
            BigColArgs<T> ca;
            BigRowArgs<T> ra;
            ca.frame = 1;
            ca.src = nullptr;
            ca.scale = (T)1;
            ca.hi = reinterpret_cast<const T *>(a.xt_cur) + (size_t)gi * L;
            ca.lo = blk == 0 ? reinterpret_cast<const T *>(a.xt_prev) + (size_t)c * L
                             : reinterpret_cast<const T *>(a.xt_cur) + (size_t)(gi - a.n_in) * L;
            ca.lo_stride = ca.hi_stride = (size_t)L;
            ra.dst = reinterpret_cast<cpair<T> *>(plan.big_scr1) + (size_t)j0 * M;
            ra.dst_stride = (size_t)M;
            ra.kmax = M;
            // the column kernel of this run writes scratch T rows [0, run): reuse the front of scratch 0 per run
            cudaError_t err = big_complex<T, false>(plan, ca, ra, run, s);
            if (err != cudaSuccess) return err;
            j0 += run;
        }
        BigSplitArgs<T> sa;
        sa.Z = reinterpret_cast<const cpair<T> *>(plan.big_scr1);
        sa.tw_main = reinterpret_cast<const T *>(plan.tw);
        sa.M = M;
        sa.fa = a;
        sa.item0 = i0;
        sa.coeff_dst = nullptr;
        sa.hbase = 0;
        sa.coeff_scale = (T)0;
        k_big_split<T><<<dim3((M / 2 + 1 + 255) / 256, items), 256, 0, s>>>(sa);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t launch_forward_big(const FftPlan &plan, const ForwardArgs &a, cudaStream_t s)
{
    if (a.n_in == 0) return cudaSuccess;
    return plan.realsize == 4 ? forward_big_t<float>(plan, a, s) : forward_big_t<double>(plan, a, s);
}

template <typename T>
static cudaError_t coeff_big_t(const FftPlan &plan, const void *taps, int n_blocks, double scale, void *H, int hbase,
                               cudaStream_t s)
{
    const int M = plan.N / 2, L = M;
    for (int b0 = 0; b0 < n_blocks; b0 += plan.big_items) {
        const int items = n_blocks - b0 < plan.big_items ? n_blocks - b0 : plan.big_items;
        BigColArgs<T> ca;
        BigRowArgs<T> ra;
        ca.frame = 1;
        ca.src = nullptr;
        ca.scale = (T)scale;
        ca.lo = nullptr;            // [0_L | scale * h]: the circular shift by L makes the first half of the inverse valid
        ca.hi = reinterpret_cast<const T *>(taps) + (size_t)b0 * L;
        ca.lo_stride = ca.hi_stride = (size_t)L;
        ra.dst = reinterpret_cast<cpair<T> *>(plan.big_scr1);
        ra.dst_stride = (size_t)M;
        ra.kmax = M;
        cudaError_t err = big_complex<T, false>(plan, ca, ra, items, s);
        if (err != cudaSuccess) return err;
        BigSplitArgs<T> sa;
        memset(&sa.fa, 0, sizeof(sa.fa));
        sa.Z = reinterpret_cast<const cpair<T> *>(plan.big_scr1);
        sa.tw_main = reinterpret_cast<const T *>(plan.tw);
        sa.M = M;
        sa.item0 = 0;
        sa.coeff_dst = reinterpret_cast<T *>(H);
        sa.hbase = hbase + b0;
        sa.coeff_scale = (T)(1.0 / (double)plan.N);
        k_big_split<T><<<dim3((M / 2 + 1 + 255) / 256, items), 256, 0, s>>>(sa);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t launch_coeff_fft_big(const FftPlan &plan, const void *taps, int n_blocks, double scale, void *H, int hbase,
                                 cudaStream_t s)
{
    return plan.realsize == 4 ? coeff_big_t<float>(plan, taps, n_blocks, scale, H, hbase, s)
                              : coeff_big_t<double>(plan, taps, n_blocks, scale, H, hbase, s);
}

template <typename T>
static cudaError_t inverse_big_t(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    const int M = plan.N / 2, L = M;
    const int total = a.n_out * a.batch;
    const int passes = a.any_xfade ? 2 : 1;
    for (int i0 = 0; i0 < total; i0 += plan.big_items) {
        const int items = total - i0 < plan.big_items ? total - i0 : plan.big_items;
        for (int pass = 0; pass < passes; pass++) {
            const bool old_pass = passes == 2 && pass == 0;
            BigMergeArgs<T> ma;
            ma.Z = reinterpret_cast<cpair<T> *>(plan.big_scr1);
            ma.tw_main = reinterpret_cast<const T *>(plan.tw);
            ma.M = M;
            ma.ia = a;
            ma.item0 = i0;
            ma.old_pass = old_pass ? 1 : 0;
            k_big_merge<T><<<dim3((M / 2 + 1 + 255) / 256, items), 256, 0, s>>>(ma);
            cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) return err;
            BigColArgs<T> ca;
            BigRowArgs<T> ra;
            ca.frame = 0;
            ca.lo = ca.hi = nullptr;
            ca.lo_stride = ca.hi_stride = 0;
            ca.scale = (T)1;
            ca.src = reinterpret_cast<const cpair<T> *>(plan.big_scr1);
            // the first L reals of the result = complex elements k < M/2 (overlap-save, fftw_convolver.c:498-501)
            T *dst = old_pass ? reinterpret_cast<T *>(plan.big_old)
                              : reinterpret_cast<T *>(a.out_time) + (size_t)i0 * L;
            ra.dst = reinterpret_cast<cpair<T> *>(dst);
            ra.dst_stride = (size_t)(L / 2);
            ra.kmax = M / 2;
            if ((err = big_complex<T, true>(plan, ca, ra, items, s)) != cudaSuccess) return err;
        }
        if (passes == 2) {
            k_big_blend<T><<<dim3((L + 255) / 256, items), 256, 0, s>>>(a, reinterpret_cast<const T *>(plan.big_old), i0, L);
            cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) return err;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_inverse_big(const FftPlan &plan, const InverseArgs &a, cudaStream_t s)
{
    if (a.n_out == 0) return cudaSuccess;
    return plan.realsize == 4 ? inverse_big_t<float>(plan, a, s) : inverse_big_t<double>(plan, a, s);
}

}  // namespace bf
