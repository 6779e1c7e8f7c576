// bf_mac_tma.cu -- the delay-line multiply-accumulate with operands staged through shared memory by the
// TMA engine's bulk copies (cp.async.bulk + mbarrier transaction counts), sm_100a.
//
// Persistent grid: every block walks work items (job, tile of TB bins [, partition split]) round robin.
// One producer lane keeps a ring of STAGES stages full; a stage holds, for one partition of the tile,
// the four contiguous runs  Re X | Im X | Re H | Im H  (PB bytes each) of the planar layout.  NC = PB/16
// consumer threads each own 16 bytes (4 float / 2 double bins) of the tile, read their operands with
// conflict-free 128-bit shared loads, and accumulate over the partitions in registers in the
// reference's order (bfrun.c:1737-1754), with the reference's roundings (bf_common.cuh).  The ring runs
// across item boundaries, so the memory system never drains between items.
//
// Same arithmetic and same results, bit for bit, as k_mac in bf_kernels.cu; only the data path differs.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bf_kernels.h"
#include "bf_sample.cuh"
#include "bf_dev_utils.cuh"

namespace bf {

template <typename T, int W>
struct __align__(16) LanesT {
    T v[W];
};

template <typename T>
__device__ __forceinline__ void cprod_t(T br, T bi, T cr, T ci, T &re, T &im)
{
    re = sub_rn(mul_rn(br, cr), mul_rn(bi, ci));
    im = add_rn(mul_rn(br, ci), mul_rn(bi, cr));
}

struct ItemRange {
    int job, tile, z, i0, i1;
};

__device__ __forceinline__ ItemRange decode_item(int item, int tiles, const MacArgs &a, const MacJob &jb)
{
    ItemRange r;
    r.job = item / (tiles * a.split);
    const int rem = item - r.job * tiles * a.split;
    r.z = rem / tiles;
    r.tile = rem - r.z * tiles;
    const int chunk = (jb.n_parts + a.split - 1) / a.split;
    r.i0 = r.z * chunk;
    r.i1 = min(jb.n_parts, r.i0 + chunk);
    return r;
}

template <typename T, int PB, int STAGES>
__global__ void __launch_bounds__(PB / 16 + 32) k_mac_tma(MacArgs a, int N, int tiles, int n_items)
{
    constexpr int W = 16 / (int)sizeof(T);
    constexpr int TB = PB / (int)sizeof(T);     // bins per tile
    constexpr int NC = PB / 16;                 // consumer threads
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * 4 * PB);
    uint64_t *empty = full + STAGES;
    const int tid = threadIdx.x;
    const int M = N >> 1;
    const int P = a.ring;
    const int slot0 = a.t;
    const T *Xall = reinterpret_cast<const T *>(a.fdl);
    const T *Hall = reinterpret_cast<const T *>(a.H);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC) {
        // ---- producer: one lane feeds the ring ----------------------------------------------------
        if (tid == NC) {
            uint64_t policy;
            // every operand byte is read exactly once per block: do not let it push useful lines out of L2
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int job = item / (tiles * a.split);
                const MacJob jb = a.jobs[job];
                if (jb.hbase < 0) {
                    continue;
                }
                const ItemRange r = decode_item(item, tiles, a, jb);
                const T *X = Xall + (size_t)jb.stream * P * N + (size_t)r.tile * TB;
                const T *H = Hall + (size_t)jb.hbase * N + (size_t)r.tile * TB;
                for (int i = r.i0; i < r.i1; i++) {
                    int slot = slot0 - i;
                    slot += (slot < 0) ? P : 0;
                    const T *xp = X + (size_t)slot * N;
                    const T *hp = H + (size_t)i * N;
                    unsigned char *dst = smem + (size_t)stage * 4 * PB;
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], 4u * PB);
                    bulk_g2s(dst, xp, PB, &full[stage], policy);
                    bulk_g2s(dst + PB, xp + M, PB, &full[stage], policy);
                    bulk_g2s(dst + 2 * PB, hp, PB, &full[stage], policy);
                    bulk_g2s(dst + 3 * PB, hp + M, PB, &full[stage], policy);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------
    typedef LanesT<T, W> L;
    int stage = 0;
    uint32_t phase = 0;
    const int lane = tid & 31;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int job = item / (tiles * a.split);
        const MacJob jb = a.jobs[job];
        const ItemRange r = decode_item(item, tiles, a, jb);
        const size_t off = (size_t)r.tile * TB + (size_t)tid * W;   // first bin of this thread
        T *out = reinterpret_cast<T *>(a.Y) + ((size_t)r.z * a.n_slots + jb.out) * N + off;
        L are, aim;
#pragma unroll
        for (int l = 0; l < W; l++) {
            are.v[l] = (T)0;
            aim.v[l] = (T)0;
        }
        if (jb.hbase < 0) {
            // dirac short cut: one partition, straight from global memory
            if (r.z == 0) {
                const T *xp = Xall + ((size_t)jb.stream * P + slot0) * N + off;
                const L xr = *reinterpret_cast<const L *>(xp), xi = *reinterpret_cast<const L *>(xp + M);
                const T fr = (T)(1.0 / (T)N);
#pragma unroll
                for (int l = 0; l < W; l++) {
                    const T s = (l & 1) ? -fr : fr;
                    are.v[l] = mul_rn(xr.v[l], s);
                    aim.v[l] = mul_rn(xi.v[l], s);
                }
            }
        } else {
            T dc = (T)0, ny = (T)0;
            for (int i = r.i0; i < r.i1; i++) {
                const unsigned char *src = smem + (size_t)stage * 4 * PB + (size_t)tid * 16;
                mbar_wait(&full[stage], phase);
                const L br = *reinterpret_cast<const L *>(src);
                const L bi = *reinterpret_cast<const L *>(src + PB);
                const L cr = *reinterpret_cast<const L *>(src + 2 * PB);
                const L ci = *reinterpret_cast<const L *>(src + 3 * PB);
                if (i == r.i0) {
#pragma unroll
                    for (int l = 0; l < W; l++) {
                        cprod_t<T>(br.v[l], bi.v[l], cr.v[l], ci.v[l], are.v[l], aim.v[l]);
                    }
                    dc = mul_rn(br.v[0], cr.v[0]);
                    ny = mul_rn(bi.v[0], ci.v[0]);
                } else {
#pragma unroll
                    for (int l = 0; l < W; l++) {
                        T re, im;
                        cprod_t<T>(br.v[l], bi.v[l], cr.v[l], ci.v[l], re, im);
                        are.v[l] = add_rn(are.v[l], re);
                        aim.v[l] = add_rn(aim.v[l], im);
                    }
                    dc = add_rn(dc, mul_rn(br.v[0], cr.v[0]));
                    ny = add_rn(ny, mul_rn(bi.v[0], ci.v[0]));
                }
                // operands are in registers and consumed: hand the stage back to the producer
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&empty[stage]);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (off == 0) {
                are.v[0] = dc;      // DC and Nyquist are real products (fftw_convfuns.h:546-547, 559-560)
                aim.v[0] = ny;
            }
        }
        *reinterpret_cast<L *>(out) = are;
        *reinterpret_cast<L *>(out + M) = aim;
    }
}

bool mac_tma_applicable(const FftPlan &plan)
{
    const int M = plan.N / 2;
    return M * plan.realsize >= 2048 && (M * plan.realsize) % 2048 == 0;
}

cudaError_t launch_mac_tma(const FftPlan &plan, const MacArgs &a, cudaStream_t s)
{
    constexpr int PB = 2048, STAGES = 8;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    const int M = plan.N / 2;
    const int tiles = (M * plan.realsize) / PB;
    const int n_items = a.n_jobs * tiles * a.split;
    const size_t smem = (size_t)STAGES * 4 * PB + 2 * STAGES * sizeof(uint64_t);
    const int per_sm = 3;
    int grid = sm_count * per_sm;
    if (grid > n_items) {
        grid = n_items;
    }
    cudaError_t e;
    if (plan.realsize == 4) {
        e = cudaFuncSetAttribute(k_mac_tma<float, PB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_mac_tma<float, PB, STAGES><<<grid, PB / 16 + 32, smem, s>>>(a, plan.N, tiles, n_items);
    } else {
        e = cudaFuncSetAttribute(k_mac_tma<double, PB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_mac_tma<double, PB, STAGES><<<grid, PB / 16 + 32, smem, s>>>(a, plan.N, tiles, n_items);
    }
    return cudaGetLastError();
}

}  // namespace bf
