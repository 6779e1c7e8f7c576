// bf_mac_tile.cu -- the batched delay-line multiply-accumulate with block-cooperative operand staging: the TMA
// engine's bulk copies fill a shared-memory ring per thread block, and the B blocks of a batch are dealt over G
// thread groups that read the SAME staged operands.
//
// Why (profiles/r2_ncu_full_summary.json, capture shard8_mac_b8): in k_mac_batch2 every thread runs its own cp.async
// ring -- 4 LDGSTS, a group commit / wait and ~35 address and bookkeeping instructions per partition step next to
// the 48-64 floating-point ones -- and all B accumulators of a bin live in ONE thread.  A small shard (8 filters of
// the 64-filter job per GPU) then has only 7-14 warps per SM, each a chain of 135 dependent steps: 44 us per launch
// where HBM needs 22 and the FP32 pipe 15.  Here
//   * one producer lane per block issues four bulk copies per step (Re H | Im H | Re X | Im X rows of the block's
//     bin tile; SASS UBLKCP, completion counted in bytes on an mbarrier) -- the consumers carry no address
//     arithmetic and no copy instructions at all;
//   * step j's coefficient row H[j] serves all B blocks, and the delay-line row that joins group 0's window at step
//     j joins group g's window BG*g steps later, straight from the ring: HBM and L2 see every operand once per
//     thread block, as before, but G times as many warps share the work (B = G * BG accumulators per bin);
//   * ring entries are handed back through a second set of mbarriers (one arrival per consumer warp), so there
//     is no block-wide barrier anywhere in the loop.
// Arithmetic, summation order and therefore every bit of the result are those of k_mac_batch2 / k_mac
// (bf_mac_acc.cuh: the reference's separately rounded products and sums, convolver_xmm.c:25-30).
//
// Ring entry q (q = -(B-1) .. n-1) holds H[i0 + q] (q >= 0) and the delay-line slot (t - i0 - q): the block that
// meets partition i0 + q in output block t.  Output block t + b at step j needs slot t + b - i0 - j = entry j - b.
// Group g (blocks g*BG .. g*BG+BG-1) therefore reads, at step j, H from entry j and its new window member from
// entry j - g*BG; entry q is dead once every warp has finished step q + (G-1)*BG.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "bf_kernels.h"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"
#include "bf_mac_acc.cuh"

namespace bf {

__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// G groups x BG blocks, TPG threads per group (2 bins each), S ring entries (power of two)
// MODE (experiments only): 0 = the kernel; 1 = operands staged and read but not multiplied (memory side alone);
// 2 = nothing staged, no waits (arithmetic side alone, on whatever the ring holds)
template <int G, int BG, int TPG, int S, int MINB, int MODE = 0>
__global__ void __launch_bounds__(G * TPG + 32, MINB) k_mac_tile(MacArgs a, int N, int tiles)
{
    constexpr int B = G * BG;
    constexpr int W = 2;
    constexpr int NC = G * TPG;                 // consumer threads
    constexpr int RB = TPG * W * 4;             // bytes per staged row
    constexpr int LAG = (G - 1) * BG;
    static_assert((S & (S - 1)) == 0 && S >= B + LAG + 2, "ring: power of two, deeper than the entries in use");
    typedef float2 V;
    extern __shared__ __align__(128) unsigned char tile_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(tile_smem + (size_t)S * 4 * RB);
    uint64_t *empty = full + S;

    const int tid = threadIdx.x;
    const int M = N >> 1;
    const int R = a.ring;
    const int job = blockIdx.x / tiles, tile = blockIdx.x - job * tiles;
    const MacJob jb = a.jobs[job];
    const int z = blockIdx.y;
    const int t0 = a.t % R;
    const int chunk = (jb.n_parts + a.split - 1) / a.split;
    const int i0 = z * chunk;
    const int n = jb.hbase < 0 ? 0 : min(jb.n_parts, i0 + chunk) - i0;     // steps of this block
    const float *X = reinterpret_cast<const float *>(a.fdl) + (size_t)jb.stream * R * N + (size_t)tile * (TPG * W);

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC) {
        // ---- producer: one lane keeps the ring full ---------------------------------------------------------------
        if (tid == NC && n > 0 && MODE != 2) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            const float *H = reinterpret_cast<const float *>(a.H) + ((size_t)jb.hbase + i0) * N + (size_t)tile * (TPG * W);
            int xs = (t0 - i0 + (B - 1)) % R;       // slot of entry q = -(B-1)
            xs += xs < 0 ? R : 0;
            for (int q = -(B - 1); q < n; q++) {
                const int k = q + (B - 1);          // entries are used in this order
                const int e = k & (S - 1);
                unsigned char *dst = tile_smem + (size_t)e * 4 * RB;
                if (k >= S) {
                    const uint32_t par = (uint32_t)((k / S - 1) & 1);
                    while (!mbar_try(&empty[e], par)) {
                    }
                }
                const float *xp = X + (size_t)xs * N;
                if (q >= 0) {
                    const float *hp = H + (size_t)q * N;
                    mbar_expect_tx(&full[e], 4u * RB);
                    bulk_g2s(dst, hp, RB, &full[e], policy);
                    bulk_g2s(dst + RB, hp + M, RB, &full[e], policy);
                } else {
                    mbar_expect_tx(&full[e], 2u * RB);
                }
                bulk_g2s(dst + 2 * RB, xp, RB, &full[e], policy);
                bulk_g2s(dst + 3 * RB, xp + M, RB, &full[e], policy);
                xs = xs == 0 ? R - 1 : xs - 1;
            }
        }
        return;
    }

    // ---- consumers --------------------------------------------------------------------------------------------------
    const int g = tid / TPG, lt = tid - g * TPG;        // group, thread within the group
    const int b0 = g * BG;
    const int v = tile * TPG + lt;                      // bin pair of the job
    const int lane = tid & 31;
    const unsigned long long nz = a.neg_zero2;
    BinPairAcc<W> acc[BG];
#pragma unroll
    for (int b = 0; b < BG; b++) {
        acc[b].zero();
    }
    if (jb.hbase < 0) {
        // coeff = -1: unit pulse = (+1/N, -1/N, ...) per bin (fftw_convfuns.h:606-619), one partition, from global memory
        if (z == 0) {
            const float fr = (float)(1.0 / (float)N);
#pragma unroll
            for (int b = 0; b < BG; b++) {
                if (b0 + b < a.batch) {
                    int s = t0 + b0 + b;
                    s -= s >= R ? R : 0;
                    const float *xp = X + (size_t)s * N + (size_t)lt * W;
                    const V xr = __ldg(reinterpret_cast<const V *>(xp)), xi = __ldg(reinterpret_cast<const V *>(xp + M));
                    acc[b].set(0, mul_rn(xr.x, fr), mul_rn(xi.x, fr));
                    acc[b].set(1, mul_rn(xr.y, -fr), mul_rn(xi.y, -fr));
                }
            }
        }
    } else if (n > 0) {
        float dc[BG], ny[BG];
        V wr[BG], wi[BG];       // window: block b0 + b of step j sits in physical slot (b - j) mod BG
#pragma unroll
        for (int b = 0; b < BG; b++) {
            dc[b] = 0.f;
            ny[b] = 0.f;
        }
        const unsigned char *mine = tile_smem + (size_t)lt * (W * 4);
        auto row = [&](unsigned int e, int op) -> V { return *reinterpret_cast<const V *>(mine + ((size_t)e * 4 + op) * RB); };
        // the B entries of step 0 (first use of every entry: parity 0)
#pragma unroll
        for (int k = 0; k < B; k++) {
            while (MODE != 2 && !mbar_try(&full[k], 0u)) {
            }
        }
#pragma unroll
        for (int b = 0; b < BG; b++) {
            const int k = (B - 1) - b0 - b;         // entry q = -(b0 + b)
            wr[b] = row(k, 2);
            wi[b] = row(k, 3);
        }
        const bool has0 = __any_sync(0xffffffffu, v == 0);
        auto release = [&](unsigned int k) {                 // this warp is through with entry number k
            __syncwarp();
            if (lane == 0 && MODE != 2) {
                mbar_arrive(&empty[k & (S - 1)]);
            }
        };
        auto body = [&](auto dcny_tag) {
            constexpr bool DCNY = decltype(dcny_tag)::value;
            {
                // step 0: convolver_convolve, a plain product
                const V hr = row(B - 1, 0), hi = row(B - 1, 1);
#pragma unroll
                for (int b = 0; b < BG; b++) {
                    acc[b].template step<false>(wr[b], wi[b], hr, hi, nz);
                    if (DCNY) {
                        dc[b] = mul_rn(wr[b].x, hr.x);
                        ny[b] = mul_rn(wi[b].x, hi.x);
                    }
                }
                // entries q = -(B-1) .. -LAG are read at step 0 only
#pragma unroll
                for (int k = 0; k <= B - 1 - LAG; k++) {
                    release(k);
                }
            }
            // steps 1 .. n-1: convolver_convolve_add; kh = entry number of step j, the group's new window member
            // comes from entry number kh - b0
            unsigned int kh = B;
            for (int base = 1; base < n; base += BG) {
#pragma unroll
                for (int u1 = 0; u1 < BG; u1++) {
                    const int j = base + u1;
                    if (j < n) {
                        const int u = (u1 + 1) % BG;            // j % BG: base = 1 mod BG
                        const unsigned int e = kh & (S - 1);
                        const uint32_t par = (kh / S) & 1u;
                        while (MODE != 2 && !mbar_try(&full[e], par)) {
                        }
                        const V hr = row(e, 0), hi = row(e, 1);
                        const unsigned int ex = (kh - (unsigned int)b0) & (S - 1);
                        wr[(BG - u) % BG] = row(ex, 2);
                        wi[(BG - u) % BG] = row(ex, 3);
#pragma unroll
                        for (int b = 0; b < BG; b++) {
                            const V xr = wr[(b - u + BG) % BG], xi = wi[(b - u + BG) % BG];
                            if (MODE == 1) {
                                if (b == (BG - u) % BG) {
                                    acc[b].re[0] = add_pair(acc[b].re[0], add_pair(xr, hr));
                                    acc[b].im[0] = add_pair(acc[b].im[0], add_pair(xi, hi));
                                }
                                continue;
                            }
                            acc[b].template step<true>(xr, xi, hr, hi, nz);
                            if (DCNY) {
                                dc[b] = add_rn(dc[b], mul_rn(xr.x, hr.x));
                                ny[b] = add_rn(ny[b], mul_rn(xi.x, hi.x));
                            }
                        }
                        release(kh - LAG);
                        kh++;
                    }
                }
            }
        };
        if (has0) {
            body(DcTrue());
        } else {
            body(DcFalse());
        }
        if (v == 0) {
#pragma unroll
            for (int b = 0; b < BG; b++) {
                acc[b].set(0, dc[b], ny[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < BG; b++) {
        if (b0 + b < a.batch) {
            float *out = reinterpret_cast<float *>(a.Y) + (((size_t)z * a.batch + b0 + b) * a.n_slots + jb.out) * N + (size_t)v * W;
            V ore, oim;
            acc[b].get(ore, oim);
            *reinterpret_cast<V *>(out) = ore;
            *reinterpret_cast<V *>(out + M) = oim;
        }
    }
}

// ---- the same ring filled by the consumers themselves: cooperative cp.async -----------------------------------------
// tools/ubench_bulk.cu (gpurun_out/r2_ubench_bulk*.txt -> profiles/): a bulk copy costs its issuing warp ~60 cycles
// whatever its size, so rows of 256-1024 bytes cap ONE producer warp at 8-32 GB/s where an SM needs 46; per-thread
// cp.async has no such limit.  Here every consumer thread copies one 16-byte piece of each ring entry
// (cp.async.cg, SASS LDGSTS.128) and lets the copy arrive on the entry's mbarrier when it lands
// (cp.async.mbarrier.arrive.noinc): the shared ring and the block groups of k_mac_tile without a producer.
// Small tiles (TPG = 32: 64 bins, rows of 256 bytes) deal 8 filters x 128 tiles over 148 SMs as 7 / 6 blocks per SM.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int G, int BG, int TPG, int S, int MINB, int MODE = 0>
__global__ void __launch_bounds__(G * TPG, MINB) k_mac_coop(MacArgs a, int N, int tiles)
{
    constexpr int B = G * BG;
    constexpr int W = 2;
    constexpr int NC = G * TPG;                 // threads
    constexpr int RB = TPG * W * 4;             // bytes per staged row
    constexpr int LAG = (G - 1) * BG;
    constexpr int CHUNKS = 4 * RB / 16;         // 16-byte pieces per entry = 2 * TPG
    constexpr int NI = NC < CHUNKS ? NC : CHUNKS;       // issuing threads
    constexpr int CH = CHUNKS / NI;             // pieces per issuing thread
    constexpr int D = (S - LAG - 3) < (S - B + 1) ? (S - LAG - 3) : (S - B + 1);   // entries requested ahead
    static_assert((S & (S - 1)) == 0 && D >= 2, "ring: power of two, deeper than the entries in use");
    static_assert(RB % 256 == 0, "a warp's pieces lie in at most two rows");
    typedef float2 V;
    extern __shared__ __align__(128) unsigned char tile_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(tile_smem + (size_t)S * 4 * RB);
    uint64_t *empty = full + S;

    const int tid = threadIdx.x;
    const int M = N >> 1;
    const int R = a.ring;
    const int job = blockIdx.x / tiles, tile = blockIdx.x - job * tiles;
    const MacJob jb = a.jobs[job];
    const int z = blockIdx.y;
    const int t0 = a.t % R;
    const int chunk = (jb.n_parts + a.split - 1) / a.split;
    const int i0 = z * chunk;
    const int n = jb.hbase < 0 ? 0 : min(jb.n_parts, i0 + chunk) - i0;     // steps of this block
    const float *X = reinterpret_cast<const float *>(a.fdl) + (size_t)jb.stream * R * N + (size_t)tile * (TPG * W);

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], NI);
            mbar_init(&empty[s], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (MODE == 2) {        // arithmetic side alone: finite operands in the ring
        for (int i = tid; i < S * 4 * RB / 4; i += NC) {
            reinterpret_cast<float *>(tile_smem)[i] = 1.0f;
        }
        __syncthreads();
    }
    const int g = tid / TPG, lt = tid - g * TPG;        // group, thread within the group
    const int b0 = g * BG;
    const int v = tile * TPG + lt;                      // bin pair of the job
    const int lane = tid & 31;
    const unsigned long long nz = a.neg_zero2;
    BinPairAcc<W> acc[BG];
#pragma unroll
    for (int b = 0; b < BG; b++) {
        acc[b].zero();
    }
    if (jb.hbase < 0) {
        // coeff = -1: unit pulse = (+1/N, -1/N, ...) per bin (fftw_convfuns.h:606-619), one partition, from global memory
        if (z == 0) {
            const float fr = (float)(1.0 / (float)N);
#pragma unroll
            for (int b = 0; b < BG; b++) {
                if (b0 + b < a.batch) {
                    int s = t0 + b0 + b;
                    s -= s >= R ? R : 0;
                    const float *xp = X + (size_t)s * N + (size_t)lt * W;
                    const V xr = __ldg(reinterpret_cast<const V *>(xp)), xi = __ldg(reinterpret_cast<const V *>(xp + M));
                    acc[b].set(0, mul_rn(xr.x, fr), mul_rn(xi.x, fr));
                    acc[b].set(1, mul_rn(xr.y, -fr), mul_rn(xi.y, -fr));
                }
            }
        }
    } else if (n > 0) {
        // ---- this thread's share of the copies: piece c = tid + i * NI of every entry ------------------------------
        const float *src[CH];           // where piece i of the NEXT entry to request comes from
        bool is_h[CH];
        unsigned int dst_off[CH];
        const unsigned int total = (unsigned int)(n + B - 1);       // entries of this block
        unsigned int k_next = 0;        // next entry number to request
        int xs = (t0 - i0 + (B - 1)) % R;       // slot of entry q = -(B-1)
        xs += xs < 0 ? R : 0;
        if (tid < NI) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int c = tid + i * NI;
                const int op = c / (RB / 16), off = (c - op * (RB / 16)) * 4;      // row, float offset within it
                is_h[i] = op < 2;
                dst_off[i] = (unsigned int)c * 16u;
                if (is_h[i]) {
                    src[i] = reinterpret_cast<const float *>(a.H) + ((size_t)jb.hbase + i0) * N + (size_t)tile * (TPG * W) +
                             (op == 1 ? M : 0) + off;
                } else {
                    src[i] = X + (size_t)xs * N + (op == 3 ? M : 0) + off;
                }
            }
        }
        const size_t wrap = (size_t)(R - 1) * N;
        auto request = [&]() {          // entry k_next; its ring slot is free
            const unsigned int e = k_next & (S - 1);
            unsigned char *dst = tile_smem + (size_t)e * 4 * RB;
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (is_h[i]) {
                    if (k_next >= (unsigned int)(B - 1)) {
                        if (MODE != 2) cp_async16(dst + dst_off[i], src[i]);
                        src[i] += N;
                    }
                } else {
                    if (MODE != 2) cp_async16(dst + dst_off[i], src[i]);
                    src[i] = xs == 0 ? src[i] + wrap : src[i] - N;
                }
            }
            xs = xs == 0 ? R - 1 : xs - 1;
            if (MODE != 2) cp_async_arrive(&full[e]);
            k_next++;
        };
        if (tid < NI) {
            for (int k = 0; k < B - 1 + D && k_next < total; k++) {     // B - 1 + D <= S: first use of every slot
                request();
            }
        }

        float dc[BG], ny[BG];
        V wr[BG], wi[BG];       // window: block b0 + b of step j sits in physical slot (b - j) mod BG
#pragma unroll
        for (int b = 0; b < BG; b++) {
            dc[b] = 0.f;
            ny[b] = 0.f;
        }
        const unsigned char *mine = tile_smem + (size_t)lt * (W * 4);
        auto row = [&](unsigned int e, int op) -> V { return *reinterpret_cast<const V *>(mine + ((size_t)e * 4 + op) * RB); };
        auto refill = [&]() {           // called once per step: request the entry D steps ahead
            if (tid < NI && k_next < total) {
                if (k_next >= (unsigned int)S) {
                    const uint32_t par = ((k_next / S) - 1u) & 1u;
                    while (MODE != 2 && !mbar_try(&empty[k_next & (S - 1)], par)) {
                    }
                }
                request();
            }
        };
        // the B entries of step 0 (first use of every entry: parity 0)
#pragma unroll
        for (int k = 0; k < B; k++) {
            while (MODE != 2 && !mbar_try(&full[k], 0u)) {
            }
        }
#pragma unroll
        for (int b = 0; b < BG; b++) {
            const int k = (B - 1) - b0 - b;         // entry q = -(b0 + b)
            wr[b] = row(k, 2);
            wi[b] = row(k, 3);
        }
        const bool has0 = __any_sync(0xffffffffu, v == 0);
        auto release = [&](unsigned int k) {        // this warp is through with entry number k
            __syncwarp();
            if (lane == 0 && MODE != 2) {
                mbar_arrive(&empty[k & (S - 1)]);
            }
        };
        auto body = [&](auto dcny_tag) {
            constexpr bool DCNY = decltype(dcny_tag)::value;
            {
                // step 0: convolver_convolve, a plain product
                const V hr = row(B - 1, 0), hi = row(B - 1, 1);
#pragma unroll
                for (int b = 0; b < BG; b++) {
                    acc[b].template step<false>(wr[b], wi[b], hr, hi, nz);
                    if (DCNY) {
                        dc[b] = mul_rn(wr[b].x, hr.x);
                        ny[b] = mul_rn(wi[b].x, hi.x);
                    }
                }
                // entries q = -(B-1) .. -LAG are read at step 0 only
#pragma unroll
                for (int k = 0; k <= B - 1 - LAG; k++) {
                    release(k);
                }
                refill();               // entry B - 1 + D may reuse the slot of entry 0, released just now
            }
            unsigned int kh = B;
            for (int base = 1; base < n; base += BG) {
#pragma unroll
                for (int u1 = 0; u1 < BG; u1++) {
                    const int j = base + u1;
                    if (j < n) {
                        const int u = (u1 + 1) % BG;            // j % BG: base = 1 mod BG
                        refill();
                        const unsigned int e = kh & (S - 1);
                        const uint32_t par = (kh / S) & 1u;
                        while (MODE != 2 && !mbar_try(&full[e], par)) {
                        }
                        const V hr = row(e, 0), hi = row(e, 1);
                        const unsigned int ex = (kh - (unsigned int)b0) & (S - 1);
                        wr[(BG - u) % BG] = row(ex, 2);
                        wi[(BG - u) % BG] = row(ex, 3);
#pragma unroll
                        for (int b = 0; b < BG; b++) {
                            const V xr = wr[(b - u + BG) % BG], xi = wi[(b - u + BG) % BG];
                            if (MODE == 1) {
                                if (b == (BG - u) % BG) {
                                    acc[b].re[0] = add_pair(acc[b].re[0], add_pair(xr, hr));
                                    acc[b].im[0] = add_pair(acc[b].im[0], add_pair(xi, hi));
                                }
                                continue;
                            }
                            acc[b].template step<true>(xr, xi, hr, hi, nz);
                            if (DCNY) {
                                dc[b] = add_rn(dc[b], mul_rn(xr.x, hr.x));
                                ny[b] = add_rn(ny[b], mul_rn(xi.x, hi.x));
                            }
                        }
                        release(kh - LAG);
                        kh++;
                    }
                }
            }
        };
        if (has0) {
            body(DcTrue());
        } else {
            body(DcFalse());
        }
        if (v == 0) {
#pragma unroll
            for (int b = 0; b < BG; b++) {
                acc[b].set(0, dc[b], ny[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < BG; b++) {
        if (b0 + b < a.batch) {
            float *out = reinterpret_cast<float *>(a.Y) + (((size_t)z * a.batch + b0 + b) * a.n_slots + jb.out) * N + (size_t)v * W;
            V ore, oim;
            acc[b].get(ore, oim);
            *reinterpret_cast<V *>(out) = ore;
            *reinterpret_cast<V *>(out + M) = oim;
        }
    }
}

template <int G, int BG, int TPG, int S, int MINB, int MODE = 0>
static cudaError_t launch_coop(const MacArgs &a, int N, cudaStream_t s)
{
    constexpr int RB = TPG * 2 * 4;
    constexpr size_t smem = (size_t)S * 4 * RB + 2 * S * sizeof(uint64_t);
    static bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(k_mac_coop<G, BG, TPG, S, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) {
            return err;
        }
        if (dev >= 0 && dev < 64) {
            configured[dev] = true;
        }
    }
    const int tiles = (N / 2) / (TPG * 2);
    dim3 grid((unsigned int)(a.n_jobs * tiles), a.split, 1);
    MacArgs args = a;
    args.neg_zero2 = 0x8000000080000000ull;
    g_last_func = (const void *)k_mac_coop<G, BG, TPG, S, MINB, MODE>;
    k_mac_coop<G, BG, TPG, S, MINB, MODE><<<grid, G * TPG, smem, s>>>(args, N, tiles);
    return cudaGetLastError();
}

template <int G, int BG, int TPG, int S, int MINB, int MODE = 0>
static cudaError_t launch_tile(const MacArgs &a, int N, cudaStream_t s)
{
    constexpr int RB = TPG * 2 * 4;
    constexpr size_t smem = (size_t)S * 4 * RB + 2 * S * sizeof(uint64_t);
    static bool configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(k_mac_tile<G, BG, TPG, S, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) {
            return err;
        }
        if (dev >= 0 && dev < 64) {
            configured[dev] = true;
        }
    }
    const int tiles = (N / 2) / (TPG * 2);
    dim3 grid((unsigned int)(a.n_jobs * tiles), a.split, 1);
    MacArgs args = a;
    args.neg_zero2 = 0x8000000080000000ull;
    g_last_func = (const void *)k_mac_tile<G, BG, TPG, S, MINB, MODE>;
    k_mac_tile<G, BG, TPG, S, MINB, MODE><<<grid, G * TPG + 32, smem, s>>>(args, N, tiles);
    return cudaGetLastError();
}

static int tile_env(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v != nullptr ? atoi(v) : dflt;
}

// float_bits 32, no powersave flags, whole partition sums, tiles of 128 bins dividing the spectrum
bool mac_tile_applicable(const FftPlan &plan, const MacArgs &a)
{
    return plan.realsize == 4 && a.slot_zero == nullptr && a.batch > 4 && a.batch <= 16 && (plan.N / 2) % 256 == 0 &&
           plan.N >= 512 && a.head == 0 && a.z_count == 0;
}

// Which launches take a shared-ring kernel by default.  Measured (profiles/r2_macsweep_coop.txt, r2_coop_ab.txt, the
// headline filter shape, 8 / 16 / 32 / 64 filters per GPU): at 8 blocks per launch k_mac_batch2 is as fast or faster
// everywhere (46.9 / 66.5 / 120 / 201 us against 45-47 / 77 / 119-134 / 209-252 us); at 16 blocks per launch its
// 241-register threads leave one block of 8 warps per SM, and k_mac_coop<2 groups x 8 blocks> wins up to 32 filters
// (75 / 120 / 212 us against 97 / 134 / 227 us; at 64 filters 403 against 376 us).
bool mac_coop_by_default(const FftPlan &plan, const MacArgs &a)
{
    static const int mode = tile_env("BFCUDA_MAC_TILE", -1);        // 0 forces k_mac_batch2
    return mode != 0 && a.batch > 8 && mac_tile_applicable(plan, a) && (long)a.n_jobs * (plan.N / 2) <= 32L * 8192;
}

cudaError_t launch_mac_tile(const FftPlan &plan, const MacArgs &a, cudaStream_t s)
{
    const int N = plan.N;
    const int which = tile_env("BFCUDA_MAC_TILE", -1);      // -1: the default rule; 1: bulk-copy staged; 2: cooperative cp.async
    if (which < 1) {
        return launch_coop<2, 8, 64, 32, 2>(a, N, s);
    }
    const int G = tile_env("BFCUDA_TILE_G", 2), TPG = tile_env("BFCUDA_TILE_TPG", 64);
    if (which == 2) {
        if (a.batch <= 8) {
#ifdef BF_MAC_SWEEP
            const int S = tile_env("BFCUDA_TILE_S", 16);
            const int mode = tile_env("BFCUDA_TILE_MODE", 0);
            if (mode == 1 && G == 2 && TPG == 32) return launch_coop<2, 4, 32, 16, 8, 1>(a, N, s);
            if (mode == 2 && G == 2 && TPG == 32) return launch_coop<2, 4, 32, 16, 8, 2>(a, N, s);
            if (mode == 1 && G == 1 && TPG == 64) return launch_coop<1, 8, 64, 16, 4, 1>(a, N, s);
            if (mode == 2 && G == 1 && TPG == 64) return launch_coop<1, 8, 64, 16, 4, 2>(a, N, s);
            if (mode == 1 && G == 4 && TPG == 32) return launch_coop<4, 2, 32, 16, 4, 1>(a, N, s);
            if (mode == 2 && G == 4 && TPG == 32) return launch_coop<4, 2, 32, 16, 4, 2>(a, N, s);
            if (G == 2 && TPG == 32 && S == 32) return launch_coop<2, 4, 32, 32, 8>(a, N, s);
            if (G == 2 && TPG == 64 && S == 32) return launch_coop<2, 4, 64, 32, 4>(a, N, s);
            if (G == 4 && TPG == 32 && S == 32) return launch_coop<4, 2, 32, 32, 4>(a, N, s);
            if (G == 1 && TPG == 32) return launch_coop<1, 8, 32, 16, 8>(a, N, s);
            if (G == 4 && TPG == 64) return launch_coop<4, 2, 64, 16, 2>(a, N, s);
            if (G == 2 && TPG == 32) return launch_coop<2, 4, 32, 16, 8>(a, N, s);
            if (G == 4 && TPG == 32) return launch_coop<4, 2, 32, 16, 4>(a, N, s);
#endif
            if (G == 1) return launch_coop<1, 8, 64, 16, 4>(a, N, s);
            return launch_coop<2, 4, 64, 16, 4>(a, N, s);
        }
#ifdef BF_MAC_SWEEP
        if (G == 2 && TPG == 32) return launch_coop<2, 8, 32, 32, 4>(a, N, s);
        if (G == 4 && TPG == 32) return launch_coop<4, 4, 32, 32, 4>(a, N, s);
        if (G == 4 && TPG == 64) return launch_coop<4, 4, 64, 32, 2>(a, N, s);
#endif
        return launch_coop<2, 8, 64, 32, 2>(a, N, s);
    }
    if (a.batch <= 8) {
#ifdef BF_MAC_SWEEP
        const int mode = tile_env("BFCUDA_TILE_MODE", 0);
        if (mode == 1 && G == 2) return launch_tile<2, 4, 64, 16, 4, 1>(a, N, s);
        if (mode == 2 && G == 2) return launch_tile<2, 4, 64, 16, 4, 2>(a, N, s);
        if (mode == 1 && G == 1) return launch_tile<1, 8, 64, 16, 4, 1>(a, N, s);
        if (mode == 2 && G == 1) return launch_tile<1, 8, 64, 16, 4, 2>(a, N, s);
        if (G == 1 && TPG == 64) return launch_tile<1, 8, 64, 16, 4>(a, N, s);
        if (G == 2 && TPG == 32) return launch_tile<2, 4, 32, 16, 8>(a, N, s);
        if (G == 2 && TPG == 128) return launch_tile<2, 4, 128, 16, 2>(a, N, s);
        if (G == 4 && TPG == 32) return launch_tile<4, 2, 32, 32, 4>(a, N, s);
        if (G == 4 && TPG == 64) return launch_tile<4, 2, 64, 32, 2>(a, N, s);
#endif
        if (G == 1) return launch_tile<1, 8, 128, 16, 2>(a, N, s);
        return launch_tile<2, 4, 64, 16, 4>(a, N, s);
    }
#ifdef BF_MAC_SWEEP
    if (G == 4 && TPG == 32) return launch_tile<4, 4, 32, 32, 4>(a, N, s);
    if (G == 4 && TPG == 64) return launch_tile<4, 4, 64, 32, 2>(a, N, s);
#endif
    return launch_tile<2, 8, 64, 32, 2>(a, N, s);
}

}  // namespace bf
