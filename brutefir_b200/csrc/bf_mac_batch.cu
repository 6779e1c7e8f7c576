// bf_mac_batch.cu -- the delay-line multiply-accumulate for B consecutive audio blocks per launch.
//
// Output block t+b needs  sum_i FDL[t+b-i] (*) H[i]  (/root/reference/bfrun.c:1737-1754 for every block of the
// batch).  Walking the partitions i upwards with B accumulators, step i uses ONE coefficient vector H[i] for all B
// blocks and the window FDL[t-i .. t-i+B-1], which differs from the previous step's window by one new slot.  So a
// step moves two operand vectors (H[i] and the new delay-line slot, real and imaginary part each) for B complex
// vector MACs instead of 2 B: the HBM traffic of a batch is rs*N*(P + (P+B-1) + B) per filter instead of
// B*rs*N*(2P+1), while every output block still accumulates its partitions in ascending order with the reference's
// separately rounded products and sums (convolver_xmm.c:25-30) -- bit-identical to B single-block launches.
//
// At B = 8 the kernel is no longer HBM-bound (0.5 -> ~3.8 flop/byte): what limits it is FP32 issue and, before
// this version, exposed load latency (ncu: 62 % of the stall samples waiting on the two-steps-ahead register
// prefetch).  Here every thread streams its operands through a private ring of S stages in shared memory with
// cp.async (LDGSTS): S-1 steps of loads in flight per thread without holding registers, no block-wide barrier
// (a thread only reads what it copied itself; cp.async.wait_group orders it).  The window lives in registers and is
// rotated by unrolling the partition loop.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "bf_kernels.h"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"
#include "bf_mac_acc.cuh"

namespace bf {


// Threads per block.  (The work is a power of two of identical threads, the machine 148 SMs: 1024 blocks of 256 on 296
// slots are 3.46 "waves" at the headline shape.  Measured: 64, 128 and 256 threads per block take the same time -- the
// blocks of the partial last wave run alone on their SMs and correspondingly faster.)
//
// Small shards (8 filters of a 64-filter job per GPU: 128 blocks of 256 threads for 148 SMs).  The first half of round 2 ran
// them one bin per thread (W = 1, PairAcc) in 64-thread blocks, because the engine's split heuristic cut the two-bin
// kernel's partition sum three ways there: 44-47 us per 8-block launch either way.  With the WHOLE sum in one thread (no
// split from four warps per SM up, choose_split in bf_engine.cu) the two-bin kernel takes 38.8 us, and 37.0 us in the
// 154-register build below -- a grid of at most one block per SM has the register file to itself -- in the reference's
// summation order.  mac_batch_lanes() is the one place that decides the lanes.
// What did NOT help at that size (profiles/r2_macsweep_*.txt, r2s_shard8_*): deeper rings (S = 12 .. 24: the kernel is
// not waiting for memory, its stalls are fixed-latency waits and math-pipe throttle with two warps per scheduler), block
// sizes that fill 148 SMs evenly (224 threads: the two-warp schedulers bound it, not the idle SMs), two steps per wait
// (PAIR = 2), two groups of four blocks per thread (59 us), the shared-ring kernels of bf_mac_tile.cu at 8 blocks per launch.
// PS: powersave instantiation (zero delay-line slots are not read); kept apart so that the plain kernel carries none of it
// PAIR = 2: two partition steps per wait -- one cp.async.wait_group, the shared loads of both steps, then the arithmetic
// of both in one scheduling window (the wait is a compiler barrier: with one step per wait every step starts with an
// exposed shared-memory load and ends with the tail of its dependent FFMA2 -> FADD2 -> FADD2 chains)
template <typename T, int W, int B, int S, int MINB, int MBT, bool PS, int PAIR = 1>
__global__ void __launch_bounds__(MBT, MINB) k_mac_batch2(MacArgs a, int N)
{
    static_assert(S >= 2, "at least one stage in flight");
    constexpr int VB = W * (int)sizeof(T);
    typedef typename VecB<T, W>::type V;
    typedef LanesB<T, W> L;
    extern __shared__ __align__(16) unsigned char mac_ring[];
    V *ring = reinterpret_cast<V *>(mac_ring) + threadIdx.x;       // [S][4][MBT] vectors, this thread's column
    auto stage_ptr = [&](int stage, int op) -> V * { return ring + (stage * 4 + op) * MBT; };

    const int M = N >> 1;
    const int vecs = M / W;
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long)a.n_jobs * vecs) {
        return;
    }
    const int job = (int)(g / vecs), v = (int)(g - (long)job * vecs);
    const MacJob jb = a.jobs[job];
    const int z = blockIdx.y;
    const int R = a.ring;
    // Batch groups (blockIdx.z): a batch larger than B is dealt over gridDim.z groups of B consecutive blocks each --
    // more threads for shards too small to fill the machine otherwise; the groups read the same coefficient vectors at
    // about the same time (second reader: L2), and every output block keeps its own left-to-right partition sum.
    const int bg = (int)blockIdx.z * B;         // first block of this thread's group
    const int t0 = (a.t + bg) % R;
    const int nb_here = a.batch - bg;           // blocks of the group that exist (<= B are computed)
    const T *X = reinterpret_cast<const T *>(a.fdl) + (size_t)jb.stream * R * N + (size_t)v * W;
    auto xslot = [&](int s) -> const T * {      // s in (-R, 2R)
        s += (s < 0) ? R : 0;
        s -= (s >= R) ? R : 0;
        return X + (size_t)s * N;
    };

    typename AccSel<T, W>::type acc[B];         // one output block's bins each
    const unsigned long long nz = a.neg_zero2;
#pragma unroll
    for (int b = 0; b < B; b++) {
        acc[b].zero();
    }

    if (jb.hbase < 0) {
        // coeff = -1: unit pulse in the shifted-coefficient convention = (+1/N, -1/N, ...) per bin
        // (fftw_convfuns.h:606-619)
        if (z == 0) {
            const T fr = (T)(1.0 / (T)N);
#pragma unroll
            for (int b = 0; b < B; b++) {
                if (b < nb_here) {
                    const T *xp = xslot(t0 + b);
                    const V xr = ldg_once(reinterpret_cast<const V *>(xp));
                    const V xi = ldg_once(reinterpret_cast<const V *>(xp + M));
                    const L lr = *reinterpret_cast<const L *>(&xr), li = *reinterpret_cast<const L *>(&xi);
#pragma unroll
                    for (int l = 0; l < W; l++) {
                        const T s = ((v * W + l) & 1) ? -fr : fr;     // sign by bin parity
                        acc[b].set(l, mul_rn(lr.v[l], s), mul_rn(li.v[l], s));
                    }
                }
            }
        }
    } else {
        const int chunk = (jb.n_parts + a.split - 1) / a.split;
        const int i0 = z * chunk;
        const int i1 = min(jb.n_parts, i0 + chunk);
        const int n = i1 - i0;                  // steps of this thread
        const T *H = reinterpret_cast<const T *>(a.H) + ((size_t)jb.hbase + i0) * N + (size_t)v * W;
        T dc[B], ny[B];
        V wr[B], wi[B];         // window: logical block b of step j sits in physical slot (b - j) mod B
#pragma unroll
        for (int b = 0; b < B; b++) {
            dc[b] = (T)0;
            ny[b] = (T)0;
        }
        if (n > 0) {
            // producer state: the next step to request.  Step j needs H[i0 + j] and the delay-line block that joins
            // the window at step j (ring slot t - i0 - j); step 0's is loaded again although the window already holds
            // it (one vector per thread and launch) so that every request is the same four copies, no branches.
            const T *hnext = H;
            int xs = t0 - i0;
            xs += (xs < 0) ? R : 0;
            const T *xnext = X + (size_t)xs * N;
            const size_t wrap = (size_t)(R - 1) * N;
            int jn = 0;
            // powersave (bfrun.c:1745-1754 skips convolve_add for zero delay-line blocks): a flagged slot is not read --
            // the zero-fill form of cp.async delivers its zeros -- and while the whole B-slot window of a step is
            // flagged its coefficient vector is not read either.  nzw = unflagged slots in the window of step jn.
            const uint8_t *zf = PS ? a.slot_zero + (size_t)jb.stream * R : nullptr;
            int nzw = 0;
            if (PS) {
#pragma unroll
                for (int b = 0; b < B; b++) {
                    int sl = xs + b;
                    sl -= (sl >= R) ? R : 0;
                    nzw += zf[sl] ? 0 : 1;
                }
            }
            auto issue = [&](int stage) {       // request step jn into `stage`; always closes a group
                const bool live = jn < n;
                bool xlive = live, hlive = live;
                if (PS && live) {
                    if (jn > 0) {               // the window moved down by one slot: xs joined, xs + B left
                        int old = xs + B;
                        old -= (old >= R) ? R : 0;
                        nzw += (zf[xs] ? 0 : 1) - (zf[old] ? 0 : 1);
                    }
                    xlive = !zf[xs];
                    hlive = nzw > 0;
                }
                const unsigned int szx = xlive ? (unsigned int)VB : 0u, szh = hlive ? (unsigned int)VB : 0u;
                const T *hp = live ? hnext : H;     // a zero-size copy reads nothing, but keep its address in bounds anyway
                cp_async<VB>(stage_ptr(stage, 0), hp, szh);
                cp_async<VB>(stage_ptr(stage, 1), hp + M, szh);
                cp_async<VB>(stage_ptr(stage, 2), xnext, szx);
                cp_async<VB>(stage_ptr(stage, 3), xnext + M, szx);
                cp_async_commit();
                hnext += N;
                xnext = xs == 0 ? xnext + wrap : xnext - N;
                xs = xs == 0 ? R - 1 : xs - 1;
                jn++;
            };
#pragma unroll
            for (int st = 0; st < S - 1; st++) {
                issue(st);
            }
            // the initial window (blocks t .. t+B-1 against partition i0) straight into registers
#pragma unroll
            for (int b = 0; b < B; b++) {
                const T *xp = xslot(t0 + b - i0);
                wr[b] = ldg_once(reinterpret_cast<const V *>(xp));
                wi[b] = ldg_once(reinterpret_cast<const V *>(xp + M));
            }
            // DC and Nyquist ride in lane 0 of the job's first vector and are REAL products (fftw_convfuns.h:546-547):
            // only the warp that holds that vector carries the two extra accumulators through the loop.
            const bool has0 = __any_sync(__activemask(), v == 0);
            auto body = [&](auto dcny_tag) {
                constexpr bool DCNY = decltype(dcny_tag)::value;
                {
                    // step 0: convolver_convolve, a plain product (peeled so that the main loop carries no
                    // assign/accumulate branch)
                    cp_async_wait<S - 2>();
                    const V hr = *stage_ptr(0, 0), hi = *stage_ptr(0, 1);
                    issue(S - 1);
                    const L cr = *reinterpret_cast<const L *>(&hr), ci = *reinterpret_cast<const L *>(&hi);
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        acc[b].template step<false>(wr[b], wi[b], hr, hi, nz);
                        if (DCNY) {
                            const L br = *reinterpret_cast<const L *>(&wr[b]), bi = *reinterpret_cast<const L *>(&wi[b]);
                            dc[b] = mul_rn(br.v[0], cr.v[0]);
                            ny[b] = mul_rn(bi.v[0], ci.v[0]);
                        }
                    }
                }
                // remaining steps: convolver_convolve_add.  The window rotation u = j % B is a compile-time constant
                // (the loop is unrolled over one rotation, base = 1 mod B); the ring stage is tracked at run time, so
                // the ring depth S is free of B and the unrolled body stays one rotation long whatever S is
                // (an S-fold unrolled body outgrew the instruction cache at S = 16 and 24).
                int st_c = 1 % S;           // stage step j reads:    j % S
                int st_p = 0;               // stage step j refills:  (j - 1) % S, consumed by the previous step
                auto fp_step = [&](int u, const V &hr, const V &hi) {     // the arithmetic of a step whose rotation is u
                    const L cr = *reinterpret_cast<const L *>(&hr), ci = *reinterpret_cast<const L *>(&hi);
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        acc[b].template step<true>(wr[(b - u + B) % B], wi[(b - u + B) % B], hr, hi, nz);
                        if (DCNY) {
                            const L br = *reinterpret_cast<const L *>(&wr[(b - u + B) % B]);
                            const L bi = *reinterpret_cast<const L *>(&wi[(b - u + B) % B]);
                            dc[b] = add_rn(dc[b], mul_rn(br.v[0], cr.v[0]));
                            ny[b] = add_rn(ny[b], mul_rn(bi.v[0], ci.v[0]));
                        }
                    }
                };
                for (int base = 1; base < n; base += B) {
#pragma unroll
                    for (int k = 0; k < B; k += PAIR) {
                        const int j = base + k;
                        if (PAIR == 2 && j + 1 < n) {
                            const int u0 = (k + 1) % B, u1 = (k + 2) % B;
                            const int st_n = st_c + 1 == S ? 0 : st_c + 1;
                            cp_async_wait<S - 3>();         // the two oldest groups have landed
                            const V hr0 = *stage_ptr(st_c, 0), hi0 = *stage_ptr(st_c, 1);
                            const V xr0 = *stage_ptr(st_c, 2), xi0 = *stage_ptr(st_c, 3);
                            const V hr1 = *stage_ptr(st_n, 0), hi1 = *stage_ptr(st_n, 1);
                            const V xr1 = *stage_ptr(st_n, 2), xi1 = *stage_ptr(st_n, 3);
                            issue(st_p);
                            issue(st_c);
                            st_p = st_n;
                            st_c = st_n + 1 == S ? 0 : st_n + 1;
                            wr[(B - u0) % B] = xr0;
                            wi[(B - u0) % B] = xi0;
                            fp_step(u0, hr0, hi0);
                            wr[(B - u1) % B] = xr1;
                            wi[(B - u1) % B] = xi1;
                            fp_step(u1, hr1, hi1);
                        } else if (j < n) {
                            const int u = (k + 1) % B;
                            cp_async_wait<S - 2>();
                            const V hr = *stage_ptr(st_c, 0), hi = *stage_ptr(st_c, 1);
                            // the block that left the window makes room for the new oldest-partition slot
                            wr[(B - u) % B] = *stage_ptr(st_c, 2);
                            wi[(B - u) % B] = *stage_ptr(st_c, 3);
                            issue(st_p);
                            st_p = st_c;
                            st_c = st_c + 1 == S ? 0 : st_c + 1;
                            fp_step(u, hr, hi);
                        }
                    }
                }
            };
            if (has0) {
                body(DcTrue());
            } else {
                body(DcFalse());
            }
            cp_async_wait<0>();
        }
        if (v == 0) {
#pragma unroll
            for (int b = 0; b < B; b++) {
                acc[b].set(0, dc[b], ny[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < B; b++) {
        if (b < nb_here) {
            T *out = reinterpret_cast<T *>(a.Y) + (((size_t)z * a.batch + bg + b) * a.n_slots + jb.out) * N + (size_t)v * W;
            V ore, oim;
            acc[b].get(ore, oim);
            *reinterpret_cast<V *>(out) = ore;
            *reinterpret_cast<V *>(out + M) = oim;
        }
    }
}

static int env_int_early(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v != nullptr ? atoi(v) : dflt;
}

template <typename T, int W, int B, int S, int REGS, int MBT, bool PS, int PAIR = 1>
static cudaError_t launch_one_ps(const MacArgs &a, int N, cudaStream_t s, int groups)
{
    // BFCUDA_MAC_SMEM_PAD (experiments): extra dynamic shared memory per block, i.e. fewer resident blocks per SM -- room
    // for the FFT-side kernels to run beside the MAC instead of after it
    static const size_t pad = (size_t)env_int_early("BFCUDA_MAC_SMEM_PAD", 0);
    const size_t smem = (size_t)S * 4 * MBT * W * sizeof(T) + pad;
    constexpr int MINB = 65536 / REGS / MBT;    // 512 threads per SM at 128 registers, 256 at 255
    static bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(k_mac_batch2<T, W, B, S, MINB, MBT, PS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
        if (err != cudaSuccess) {
            return err;
        }
        if (dev >= 0 && dev < 64) {
            configured[dev] = true;
        }
    }
    const long threads = (long)a.n_jobs * (N / 2 / W);
    dim3 grid((unsigned int)((threads + MBT - 1) / MBT), a.split, groups);
    MacArgs args = a;
    args.neg_zero2 = 0x8000000080000000ull;     // (-0.0f, -0.0f): see BinPairAcc
    g_last_func = (const void *)k_mac_batch2<T, W, B, S, MINB, MBT, PS, PAIR>;
    k_mac_batch2<T, W, B, S, MINB, MBT, PS, PAIR><<<grid, MBT, smem, s>>>(args, N);
    return cudaGetLastError();
}

template <typename T, int W, int B, int S, int REGS = 128, int MBT = 256, int PAIR = 1>
static cudaError_t launch_one(const MacArgs &a, int N, cudaStream_t s, int groups = 1)
{
    return a.slot_zero != nullptr ? launch_one_ps<T, W, B, S, REGS, MBT, true, PAIR>(a, N, s, groups)
                                  : launch_one_ps<T, W, B, S, REGS, MBT, false, PAIR>(a, N, s, groups);
}

static int sm_count_cached()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
    }
    return n;
}

// Instantiations: (lanes per thread W, batch B, ring stages S).  Larger batches use narrower vectors so that
// B accumulators + the B-slot window stay within 128 registers (2 blocks of 256 threads per SM).  A batch smaller
// than B leaves the surplus accumulators unused (their window slots are still read, always inside the ring).
static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v != nullptr ? atoi(v) : dflt;
}

// Bins per thread of the batched kernel for a launch of `n_jobs` jobs over M = N/2 bins: the one table both the
// launcher and the engine's split heuristic (choose_split, bf_engine.cu) read.
int mac_batch_lanes(int realsize, int batch, int n_jobs, int N)
{
    const long bins = (long)n_jobs * (N / 2);
#ifdef BF_MAC_SWEEP
    if (realsize == 4 && batch > 4 && batch <= 8 && getenv("BFCUDA_MAC_W") != nullptr) {
        return atoi(getenv("BFCUDA_MAC_W"));
    }
#endif
    if (realsize == 4) {
        if (batch <= 4) return 4;
        // 8 blocks per launch: two bins per thread at every size.  (Round 2 first ran shards of <= 8 filters with one bin
        // per thread because the engine's split rule cut the two-bin kernel's partition sum three ways there; with the
        // whole sum in one thread the two-bin kernel takes 37-39 us per launch on an 8-filter shard against 47 us --
        // profiles/r2_macsweep_w2split1.txt.  BFCUDA_MAC_NARROW_MAX_BINS brings the old rule back for A/B runs.)
        static const int narrow_max = env_int("BFCUDA_MAC_NARROW_MAX_BINS", 0);
        if (batch <= 8) return bins <= narrow_max ? 1 : 2;
        // 16 blocks per launch: two bins per thread (241 registers, one 256-thread block per SM) measured faster than one
        // bin per thread even at 8 filters (68 against 84 us per launch on a rank of an 8-GPU run)
        static const int narrow16 = env_int("BFCUDA_MAC_B16_NARROW", 0);
        return narrow16 ? 1 : 2;
    }
    if (batch <= 2) return 2;
    return 1;
}

cudaError_t launch_mac_batch2(const FftPlan &plan, const MacArgs &a, cudaStream_t s)
{
    const int lanes = mac_batch_lanes(plan.realsize, a.batch, a.n_jobs, plan.N);
    // the shared-ring kernels (bf_mac_tile.cu): by default for 16-block launches of small shards; BFCUDA_MAC_TILE = 1 / 2
    // forces the bulk-copy staged / the cooperative cp.async kernel wherever they apply, 0 turns them off
    static const int tile_mode = env_int("BFCUDA_MAC_TILE", -1);
    if ((tile_mode >= 1 && mac_tile_applicable(plan, a)) || mac_coop_by_default(plan, a)) {
        return launch_mac_tile(plan, a, s);
    }
    if (plan.realsize == 4) {
#ifdef BF_MAC_SWEEP
        if (a.batch == 1) {
            const int S = env_int("BFCUDA_MAC_S", 8);
            if (S == 16) return launch_one<float, 4, 1, 16, 64, 256>(a, plan.N, s);
            if (S == 12) return launch_one<float, 4, 1, 12, 64, 256>(a, plan.N, s);
            return launch_one<float, 4, 1, 8, 64, 256>(a, plan.N, s);
        }
#endif
        if (a.batch <= 2) return launch_one<float, 4, 2, 4>(a, plan.N, s);
        if (a.batch <= 4) return launch_one<float, 4, 4, 4>(a, plan.N, s);
        if (a.batch <= 8) {
#ifdef BF_MAC_SWEEP
            // experiment build: (ring stages, threads per block) from the environment
            const int S = env_int("BFCUDA_MAC_S", 8), TPB = env_int("BFCUDA_MAC_TPB", lanes == 1 ? 64 : 256);
            if (lanes == 1) {
                if (S == 8 && TPB == 64) return launch_one<float, 1, 8, 8, 96, 64>(a, plan.N, s);
                if (S == 16 && TPB == 64) return launch_one<float, 1, 8, 16, 96, 64>(a, plan.N, s);
                if (S == 24 && TPB == 64) return launch_one<float, 1, 8, 24, 96, 64>(a, plan.N, s);
                if (S == 16 && TPB == 128) return launch_one<float, 1, 8, 16, 96, 128>(a, plan.N, s);
                return cudaErrorInvalidValue;
            }
            if (env_int("BFCUDA_MAC_GROUPS", 1) == 2) {
                // two groups of four blocks
                if (lanes == 4) return launch_one<float, 4, 4, 4>(a, plan.N, s, 2);
                if (TPB == 128) return launch_one<float, 2, 4, 8, 64, 128>(a, plan.N, s, 2);
                return launch_one<float, 2, 4, 8, 64, 256>(a, plan.N, s, 2);
            }
            if (S == 8 && TPB == 256) return launch_one<float, 2, 8, 8>(a, plan.N, s);
            if (S == 8 && TPB == 128) return launch_one<float, 2, 8, 8, 128, 128>(a, plan.N, s);
            if (S == 8 && TPB == 224) return launch_one<float, 2, 8, 8, 128, 224>(a, plan.N, s);
            if (S == 8 && TPB == 1256) return launch_one<float, 2, 8, 8, 255, 256>(a, plan.N, s);
            if (S == 8 && TPB == 3256) return launch_one<float, 2, 8, 8, 255, 256, 2>(a, plan.N, s);
            if (S == 8 && TPB == 4256) return launch_one<float, 2, 8, 8, 128, 256, 2>(a, plan.N, s);
            if (S == 12 && TPB == 3256) return launch_one<float, 2, 8, 12, 255, 256, 2>(a, plan.N, s);
            if (S == 8 && TPB == 1224) return launch_one<float, 2, 8, 8, 255, 224>(a, plan.N, s);
            if (S == 8 && TPB == 1192) return launch_one<float, 2, 8, 8, 255, 192>(a, plan.N, s);
            if (S == 8 && TPB == 1160) return launch_one<float, 2, 8, 8, 255, 160>(a, plan.N, s);
            if (S == 10 && TPB == 1224) return launch_one<float, 2, 8, 10, 255, 224>(a, plan.N, s);
            if (S == 12 && TPB == 1224) return launch_one<float, 2, 8, 12, 255, 224>(a, plan.N, s);
            if (S == 16 && TPB == 1224) return launch_one<float, 2, 8, 16, 255, 224>(a, plan.N, s);
            if (S == 8 && TPB == 1128) return launch_one<float, 2, 8, 8, 255, 128>(a, plan.N, s);
            if (S == 8 && TPB == 2256) return launch_one<float, 2, 8, 8, 168, 256>(a, plan.N, s);
            if (S == 8 && TPB == 192) return launch_one<float, 2, 8, 8, 128, 192>(a, plan.N, s);
            if (S == 8 && TPB == 160) return launch_one<float, 2, 8, 8, 128, 160>(a, plan.N, s);
            if (S == 8 && TPB == 96) return launch_one<float, 2, 8, 8, 128, 96>(a, plan.N, s);
            if (S == 8 && TPB == 448) return launch_one<float, 2, 8, 8, 128, 448>(a, plan.N, s);
            if (S == 10 && TPB == 256) return launch_one<float, 2, 8, 10>(a, plan.N, s);
            if (S == 12 && TPB == 256) return launch_one<float, 2, 8, 12>(a, plan.N, s);
            if (S == 12 && TPB == 128) return launch_one<float, 2, 8, 12, 128, 128>(a, plan.N, s);
            if (S == 16 && TPB == 128) return launch_one<float, 2, 8, 16, 128, 128>(a, plan.N, s);
            if (S == 16 && TPB == 256) return launch_one<float, 2, 8, 16, 128, 256>(a, plan.N, s);
            if (S == 24 && TPB == 256) return launch_one<float, 2, 8, 24, 128, 256>(a, plan.N, s);
            return cudaErrorInvalidValue;
#else
            if (lanes == 1) return launch_one<float, 1, 8, 8, 96, 64>(a, plan.N, s);
            // a grid of at most one 256-thread block per SM has the register file to itself: 154 registers instead of 128
            // give the scheduler the temporaries to keep dependent FFMA2 / FADD2 pairs apart (37.0 against 38.8 us at 8 filters)
            if ((long)a.n_jobs * (plan.N / 4) * a.split <= 256L * sm_count_cached()) {
                return launch_one<float, 2, 8, 8, 255, 256>(a, plan.N, s);
            }
            return launch_one<float, 2, 8, 8>(a, plan.N, s);
#endif
        }
        if (a.batch <= 16) {
            return lanes == 1 ? launch_one<float, 1, 16, 16, 128, 64>(a, plan.N, s)
                              : launch_one<float, 2, 16, 16, 256>(a, plan.N, s);
        }
        return cudaErrorInvalidValue;
    }
    if (a.batch <= 2) return launch_one<double, 2, 2, 4>(a, plan.N, s);
    if (a.batch <= 4) return launch_one<double, 1, 4, 8>(a, plan.N, s);
    if (a.batch <= 8) return launch_one<double, 1, 8, 8>(a, plan.N, s);
    return cudaErrorInvalidValue;
}

}  // namespace bf
