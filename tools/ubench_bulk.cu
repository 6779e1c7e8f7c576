// ubench_bulk.cu -- the memory side of the batched MAC alone, on the MAC's own access pattern: F filters, each a
// stream of n = P + B - 1 steps; a step fetches two rows of H (Re | Im of one partition) and two rows of the delay
// line for every bin tile.  Which way of issuing those copies sustains what, at 8 filters (one rank of 8) and 64?
//   variant 0: one producer lane, four bulk copies per ring entry (k_mac_tile as first written)
//   variant 1: one producer warp, lanes 0..3 issue one of the four copies each
//   variant 2: NP producer warps, entries dealt round robin
//   variant 3: no producer: every consumer warp's lane 0 issues the copies of the entries k = warp (mod warps)
//   variant 4: per-thread cp.async rings (k_mac_batch2's data path)
// Consumers wait, read their 8 bytes of each of the four rows, add them up, hand the entry back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_bulk tools/ubench_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Args {
    const float *H, *X;     // [F][P][N], [F][R][N]
    float *out;
    int N, P, R, n, tiles, RB, S, variant, np;
    long rsH, fsH, rsX, fsX;   // row / filter strides in floats
    int rot;                   // diagnostic: tiles start `rot * tile` steps apart (not the MAC's order)
};

// consumers: RB / 8 threads; producers: np warps after them
__global__ void k_bulk(Args a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int RB = a.RB, S = a.S, NC = RB / 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * 4 * RB);
    uint64_t *empty = full + S;
    const int tid = threadIdx.x;
    const int f = blockIdx.x / a.tiles, tile = blockIdx.x - f * a.tiles;
    const int M = a.N / 2;
    const float *H = a.H + (size_t)f * a.fsH + (size_t)tile * (RB / 4);
    const float *X = a.X + (size_t)f * a.fsX + (size_t)tile * (RB / 4);
    const int koff = a.rot * tile;
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], a.variant == 1 ? 4 : 1);
            mbar_init(&empty[s], NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int k, int part) {     // part: -1 = all four copies, 0..3 = one of them
        const int e = k % S;
        unsigned char *dst = smem + (size_t)e * 4 * RB;
        const float *hp = H + (size_t)((k + koff) % a.P) * a.rsH;
        const float *xp = X + (size_t)((a.R - 1 - (k + koff) % a.R)) * a.rsX;
        if (part < 0) {
            mbar_expect_tx(&full[e], 4u * RB);
            bulk_g2s(dst, hp, RB, &full[e]);
            bulk_g2s(dst + RB, hp + M, RB, &full[e]);
            bulk_g2s(dst + 2 * RB, xp, RB, &full[e]);
            bulk_g2s(dst + 3 * RB, xp + M, RB, &full[e]);
        } else {
            mbar_expect_tx(&full[e], (uint32_t)RB);
            const float *src = part == 0 ? hp : part == 1 ? hp + M : part == 2 ? xp : xp + M;
            bulk_g2s(dst + part * RB, src, RB, &full[e]);
        }
    };
    if (tid >= NC) {
        const int pw = (tid - NC) / 32, lane = tid & 31;
        if (a.variant == 0 && pw == 0 && lane == 0) {
            for (int k = 0; k < a.n; k++) {
                if (k >= S) {
                    while (!mbar_try(&empty[k % S], (uint32_t)((k / S - 1) & 1))) {
                    }
                }
                issue(k, -1);
            }
        } else if (a.variant == 1 && pw == 0 && lane < 4) {
            for (int k = 0; k < a.n; k++) {
                if (k >= S) {
                    while (!mbar_try(&empty[k % S], (uint32_t)((k / S - 1) & 1))) {
                    }
                }
                issue(k, lane);
            }
        } else if (a.variant == 2 && lane == 0) {
            for (int k = pw; k < a.n; k += a.np) {
                if (k >= S) {
                    while (!mbar_try(&empty[k % S], (uint32_t)((k / S - 1) & 1))) {
                    }
                }
                issue(k, -1);
            }
        }
        return;
    }
    const int warp = tid / 32, lane = tid & 31, nw = NC / 32;
    float2 acc = make_float2(0.f, 0.f);
    if (a.variant == 3) {
        // prologue: the first S entries, dealt over the warps
        if (lane == 0) {
            for (int k = warp; k < S && k < a.n; k += nw) {
                issue(k, -1);
            }
        }
    }
    for (int k = 0; k < a.n; k++) {
        const int e = k % S;
        while (!mbar_try(&full[e], (uint32_t)((k / S) & 1))) {
        }
        const unsigned char *src = smem + (size_t)e * 4 * RB + (size_t)tid * 8;
#pragma unroll
        for (int op = 0; op < 4; op++) {
            const float2 v = *reinterpret_cast<const float2 *>(src + op * RB);
            acc.x += v.x;
            acc.y += v.y;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[e]);
        }
        if (a.variant == 3 && lane == 0) {
            // refill: entry k2 = k - LAGR + S is this warp's if k2 % nw == warp; it reuses the slot of entry k - LAGR,
            // which every warp released LAGR steps ago (almost always without waiting)
            const int LAGR = 2;
            const int k2 = k - LAGR + S;
            if (k >= LAGR && k2 < a.n && k2 % nw == warp) {
                while (!mbar_try(&empty[k2 % S], (uint32_t)((k2 / S - 1) & 1))) {
                }
                issue(k2, -1);
            }
        }
    }
    a.out[(size_t)blockIdx.x * NC + tid] = acc.x + acc.y;
}

// variant 6: warp tiles.  A unit = 64 bins of one filter (rows of 256 bytes) owned by ONE consumer warp; U units per
// block, one producer warp whose lane u runs unit u's ring.  Fine-grained units deal out evenly over 148 SMs.
template <int S>
__global__ void k_units(Args a, int U, int units)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = 256;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)U * S * 4 * RB);     // [U][S]
    uint64_t *empty = full + U * S;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
    const int M = a.N / 2;
    const int tiles = M / 64;
    if (tid < U * S) {
        mbar_init(&full[tid], 1);
        mbar_init(&empty[tid], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp == U) {
        const int unit = blockIdx.x * U + lane;
        if (lane < U && unit < units) {
            const int f = unit / tiles, tile = unit - f * tiles;
            const float *hp = a.H + (size_t)f * a.fsH + (size_t)tile * 64;
            const float *xp = a.X + (size_t)f * a.fsX + (size_t)(a.R - 1) * a.rsX + (size_t)tile * 64;
            unsigned char *base = smem + (size_t)lane * S * 4 * RB;
            uint64_t *fu = full + lane * S, *eu = empty + lane * S;
            for (int k = 0; k < a.n; k++) {
                const int e = k & (S - 1);
                if (k >= S) {
                    while (!mbar_try(&eu[e], (uint32_t)((k / S - 1) & 1))) {
                    }
                }
                unsigned char *dst = base + (size_t)e * 4 * RB;
                mbar_expect_tx(&fu[e], 4u * RB);
                bulk_g2s(dst, hp, RB, &fu[e]);
                bulk_g2s(dst + RB, hp + M, RB, &fu[e]);
                bulk_g2s(dst + 2 * RB, xp, RB, &fu[e]);
                bulk_g2s(dst + 3 * RB, xp + M, RB, &fu[e]);
                hp += a.rsH;
                xp -= a.rsX;
            }
        }
        return;
    }
    const int unit = blockIdx.x * U + warp;
    if (unit >= units) {
        return;
    }
    const unsigned char *base = smem + (size_t)warp * S * 4 * RB + (size_t)lane * 8;
    uint64_t *fu = full + warp * S, *eu = empty + warp * S;
    float2 acc = make_float2(0.f, 0.f);
    for (int k = 0; k < a.n; k++) {
        const int e = k & (S - 1);
        while (!mbar_try(&fu[e], (uint32_t)((k / S) & 1))) {
        }
#pragma unroll
        for (int op = 0; op < 4; op++) {
            const float2 v = *reinterpret_cast<const float2 *>(base + ((size_t)e * 4 + op) * RB);
            acc.x += v.x;
            acc.y += v.y;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&eu[e]);
        }
    }
    a.out[(size_t)unit * 32 + lane] = acc.x + acc.y;
}

// evict the operands from L2 WITHOUT leaving dirty lines behind (a memset would: their write-back then competes with
// the measured reads)
__global__ void k_flush(const float4 *p, size_t n, float *out)
{
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(p + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 12345.f) {
        *out = acc;
    }
}

// variant 5: the same bytes as one plain grid-stride stream (what the memory system gives a kernel with no structure)
__global__ void k_stream(const float4 *h, size_t nh, const float4 *x, size_t nx, float *out)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < nh; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = i + u * stride < nh ? __ldg(h + i + u * stride) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    for (size_t i = i0; i < nx; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = i + u * stride < nx ? __ldg(x + i + u * stride) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 12345.f) {
        *out = acc;
    }
}

// variant 4: per-thread cp.async ring, 8 bytes x 4 rows per step, S stages
template <int S>
__global__ void k_cpasync(Args a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int NT = blockDim.x;
    const int M = a.N / 2;
    const long g = (long)blockIdx.x * NT + threadIdx.x;
    const int vecs = M / 2;
    const int f = (int)(g / vecs), v = (int)(g - (long)f * vecs);
    const float *H = a.H + (size_t)f * a.fsH + (size_t)v * 2;
    const float *X = a.X + (size_t)f * a.fsX + (size_t)v * 2;
    const int koff = a.rot * (int)(blockIdx.x % (vecs / NT));
    float2 *ring = reinterpret_cast<float2 *>(smem) + threadIdx.x;
    auto issue = [&](int k, int st) {
        const float *hp = H + (size_t)((k + koff) % a.P) * a.rsH;
        const float *xp = X + (size_t)((a.R - 1 - (k + koff) % a.R)) * a.rsX;
        const unsigned sz = k < a.n ? 8u : 0u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(ring + (st * 4 + 0) * NT)), "l"(hp), "r"(sz) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(ring + (st * 4 + 1) * NT)), "l"(hp + M), "r"(sz) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(ring + (st * 4 + 2) * NT)), "l"(xp), "r"(sz) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(ring + (st * 4 + 3) * NT)), "l"(xp + M), "r"(sz) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int k = 0; k < S - 1; k++) {
        issue(k, k);
    }
    float2 acc = make_float2(0.f, 0.f);
    int st = 0;
    for (int k = 0; k < a.n; k++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 2) : "memory");
#pragma unroll
        for (int op = 0; op < 4; op++) {
            const float2 v2 = ring[(st * 4 + op) * NT];
            acc.x += v2.x;
            acc.y += v2.y;
        }
        issue(k + S - 1, (st + S - 1) % S);
        st = st + 1 == S ? 0 : st + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    a.out[g] = acc.x + acc.y;
}

int main(int argc, char **argv)
{
    const int N = 16384, P = 128, B = 8, R = 2 * P + 2 * B, n = P + B - 1;
    const int maxF = 64;
    float *H, *X, *out;
    const size_t slack = 64u << 20;
    CK(cudaMalloc(&H, (size_t)maxF * P * N * 4 + slack));
    CK(cudaMalloc(&X, (size_t)maxF * R * N * 4 + slack));
    CK(cudaMalloc(&out, (size_t)maxF * N * 4));
    CK(cudaMemset(H, 0, (size_t)maxF * P * N * 4 + slack));
    CK(cudaMemset(X, 0, (size_t)maxF * R * N * 4 + slack));
    float *flush;
    const size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_units<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_units<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_units<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_cpasync<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_cpasync<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    auto run = [&](const char *label, int F, int variant, int RB, int S, int np, int nt4, long rpad, long fpadH, long fpadX, int rot) {
        Args a{ H, X, out, N, P, R, n, 0, RB, S, variant, np };
        a.rsH = N + rpad; a.rsX = N + rpad;
        a.fsH = (long)P * a.rsH + fpadH; a.fsX = (long)R * a.rsX + fpadX;
        a.rot = rot;
        float best = 1e9f;
        for (int rep = 0; rep < 7; rep++) {
            k_flush<<<148 * 8, 256>>>(reinterpret_cast<const float4 *>(flush), flush_bytes / 16, out);
            CK(cudaEventRecord(e0));
            if (variant == 6) {
                const int U = np, units = F * (N / 2 / 64);
                const size_t sm = (size_t)U * S * 1024 + 2 * (size_t)U * S * 8;
                if (S == 16) k_units<16><<<(units + U - 1) / U, (U + 1) * 32, sm>>>(a, U, units);
                else if (S == 32) k_units<32><<<(units + U - 1) / U, (U + 1) * 32, sm>>>(a, U, units);
                else k_units<8><<<(units + U - 1) / U, (U + 1) * 32, sm>>>(a, U, units);
            } else if (variant == 5) {
                k_stream<<<148 * np, 256>>>(reinterpret_cast<const float4 *>(H), (size_t)F * P * N / 4, reinterpret_cast<const float4 *>(X), (size_t)F * (size_t)n * N / 4, out);
            } else if (variant == 4) {
                const long threads = (long)F * (N / 4);
                const size_t sm = (size_t)S * 4 * nt4 * 8;
                if (S == 8) k_cpasync<8><<<(unsigned)(threads / nt4), nt4, sm>>>(a);
                else k_cpasync<16><<<(unsigned)(threads / nt4), nt4, sm>>>(a);
            } else {
                a.tiles = (N / 2 * 4) / RB;
                const int NC = RB / 8;
                const int nprod = variant == 3 ? 0 : (variant == 2 ? np : 1);
                const size_t sm = (size_t)S * 4 * RB + 2 * S * 8;
                k_bulk<<<F * a.tiles, NC + 32 * nprod, sm>>>(a);
            }
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double bytes = (double)F * n * 2.0 * N * 4;
        printf("%-34s F %2d v%d RB %4d S %2d np %d nt %3d : %7.1f us  %5.2f TB/s\n", label, F, variant, RB, S, np, nt4,
               best * 1e3, bytes / best / 1e9);
        fflush(stdout);
    };
    CK(cudaMemset(flush, 0, flush_bytes));
    for (int F : { 8, 16, 32, 64 }) {
        run("plain stream, 8 blocks/SM", F, 5, 0, 0, 8, 0, 0, 0, 0, 0);
        run("cp.async rings S 8 nt 256", F, 4, 0, 8, 0, 256, 0, 0, 0, 0);
        for (int S : { 8, 16, 32 }) {
            for (int U : { 4, 7, 8, 14 }) {
                if ((size_t)U * S * 1024 > 200 * 1024) continue;
                run("warp-tile units (np = U)", F, 6, 256, S, U, 0, 0, 0, 0, 0);
            }
        }
    }
    return 0;
}
