// ubench_fp32x2.cu -- issue rate of the FP32 instructions the batched MAC is made of, per SM sub-partition (SMSP).
// Each warp runs CH independent dependency chains of one instruction kind for ITERS rounds; cycles per warp-instruction
// per SMSP = elapsed clocks * 4 SMSPs / (warps per SM * instructions per warp).  Build: see tools/README or
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench_fp32x2 tools/ubench_fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int CH = 16, ITERS = 4096;

template <int KIND>
__global__ void k(u64 *out, u64 seed, long long *clk)
{
    u64 x[CH], y = seed + threadIdx.x, nz = seed;
    float f[CH], g = (float)threadIdx.x;
#pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = seed * (i + 1) + threadIdx.x; f[i] = (float)(i + threadIdx.x); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (KIND == 0) x[i] = ffma2(x[i], y, nz);
            if (KIND == 1) x[i] = fadd2(x[i], y);
            if (KIND == 2) f[i] = ffma(f[i], g, g);
            if (KIND == 3) f[i] = fmul(f[i], g);
            if (KIND == 4) f[i] = fadd(f[i], g);
            if (KIND == 5) { if (i & 1) x[i] = fadd2(x[i], x[i - 1]); else x[i] = ffma2(y, nz, x[i]); }    // the MAC's 1:1 mix
            if (KIND == 6) { if (i & 1) f[i] = fadd(f[i], g); else f[i] = fmul(f[i], g); }
        }
    }
    long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i] + (u64)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// The batched MAC's inner step without its memory traffic: 8 blocks x (4 FFMA2 + 4 FADD2) on a register window,
// distinct operands per instruction as in k_mac_batch2<float, 2, 8, 8> -- does the register file keep up?
__global__ void k_macstep(u64 *out, u64 seed, long long *clk)
{
    u64 wr[8], wi[8], re[8], im[8], cr = seed + threadIdx.x, ci = seed * 3 + threadIdx.x, nci = ci ^ 0x8000000080000000ull;
    const u64 nz = seed;
#pragma unroll
    for (int b = 0; b < 8; b++) { wr[b] = seed * (b + 2); wi[b] = seed * (b + 11) + threadIdx.x; re[b] = 0; im[b] = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const u64 p1 = ffma2(wr[b], cr, nz), p2 = ffma2(wi[b], nci, nz);
            const u64 p3 = ffma2(wr[b], ci, nz), p4 = ffma2(wi[b], cr, nz);
            re[b] = fadd2(re[b], fadd2(p1, p2));
            im[b] = fadd2(im[b], fadd2(p3, p4));
        }
        cr = re[7];         // the next step's coefficient: new values every step (a rename, no instruction), or the
        ci = im[7];         // compiler hoists the products out of the loop
        nci = im[3];
    }
    long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) s += re[b] + im[b];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static void run_macstep(int warps_per_sm)
{
    int sms = 148;
    u64 *out; long long *clk;
    cudaMalloc(&out, sizeof(u64) * sms * 1024); cudaMalloc(&clk, sizeof(long long) * sms);
    for (int r = 0; r < 2; r++) {
        k_macstep<<<sms, warps_per_sm * 32>>>(out, 0x8000000080000000ull, clk);
        cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms; i++) mean += h[i]; mean /= sms;
    printf("%-22s warps/SM %2d: %.2f cycles per warp-instruction per SMSP\n", "MAC step (registers)", warps_per_sm,
           mean * 4.0 / ((double)warps_per_sm * 64 * ITERS));
    cudaFree(out); cudaFree(clk);
}

template <int KIND>
static void run(const char *name, int warps_per_sm)
{
    int sms = 148;
    u64 *out; long long *clk;
    cudaMalloc(&out, sizeof(u64) * sms * 1024); cudaMalloc(&clk, sizeof(long long) * sms);
    k<KIND><<<sms, warps_per_sm * 32>>>(out, 0x8000000080000000ull, clk);
    cudaDeviceSynchronize();
    k<KIND><<<sms, warps_per_sm * 32>>>(out, 0x8000000080000000ull, clk);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms; i++) mean += h[i]; mean /= sms;
    double per = mean * 4.0 / ((double)warps_per_sm * CH * ITERS);
    printf("%-22s warps/SM %2d: %.2f cycles per warp-instruction per SMSP\n", name, warps_per_sm, per);
    cudaFree(out); cudaFree(clk);
}

int main()
{
    for (int w : {4, 8, 16}) {
        run<0>("FFMA2", w); run<1>("FADD2", w); run<2>("FFMA", w); run<3>("FMUL", w); run<4>("FADD", w);
        run<5>("FFMA2+FADD2 1:1", w); run<6>("FMUL+FADD 1:1", w);
        run_macstep(w);
    }
    return 0;
}
