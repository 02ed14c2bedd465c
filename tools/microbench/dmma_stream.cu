// dmma_stream.cu -- what DMMA rate does the contraction loop of profile_kernel sustain by itself?
// Every warp owns a table buffer in shared memory (32 rows x (34 + 34) doubles, as profile_kernel<4>) and loops over
// "groups": 8 k-steps of 4 LDS.128 + 16 DMMA.8x8x4 with the fragments of the next step fetched ahead -- the kernel's
// contraction without any table build.  Variants: operands from shared memory or from registers only; table
// contents zero / ordinary / mixed magnitudes (is the DMMA rate data dependent?); 1, 2 or 3 warps per sub-partition;
// a __syncwarp per group; n extra dependent DMULs per group (a stand-in for the table build).
// Reported: pipe cycles per DMMA per sub-partition at the measured SM clock (16 = the pipe's issue interval).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_stream dmma_stream.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int chunk_pair(int g) { return (g & 4) | ((g & 1) << 1) | ((g >> 1) & 1); }

constexpr int US = 34, WS = 34, kWarpDoubles = 32 * (US + WS);

template <bool SMEM, int NMUL, bool SYNC>
__global__ void __launch_bounds__(128, 3) stream(double* sink, int groups, int fill, long long* cycles)
{
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Us = smem + warp * kWarpDoubles;
    double* Ws = Us + 32 * US;
    for (int i = lane; i < kWarpDoubles; i += 32) {
        double v = 0.0;
        if (fill == 1) v = 1e-3 * (1 + (i % 97));
        if (fill == 2) v = ldexp(1.0 + 1e-3 * (i % 89), -(i % 61) * 13);  // magnitudes from 1 down to 2^-780
        Us[i] = v;
    }
    __syncwarp();
    const int fk = lane & 3, fg = lane >> 2;
    const double* fragU = Us + fk * US + 4 * chunk_pair(fg);
    const double* fragW = Ws + fk * WS + 4 * chunk_pair(fg);
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a0[4], b0[4], a1[4], b1[4];
    auto fetch = [&](double (&a)[4], double (&b)[4], int step) {
        if constexpr (SMEM) {
            const double* up = fragU + step * (4 * US);
            const double* wp = fragW + step * (4 * WS);
            const double2 u01 = *reinterpret_cast<const double2*>(up), u23 = *reinterpret_cast<const double2*>(up + 2);
            const double2 w01 = *reinterpret_cast<const double2*>(wp), w23 = *reinterpret_cast<const double2*>(wp + 2);
            a[0] = u01.x; a[1] = u01.y; a[2] = u23.x; a[3] = u23.y;
            b[0] = w01.x; b[1] = w01.y; b[2] = w23.x; b[3] = w23.y;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = a[i] * 1.0000001; b[i] = b[i]; }
        }
    };
    auto mma_step = [&](const double (&a)[4], const double (&b)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    };
    if constexpr (!SMEM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { a0[i] = fragU[i]; b0[i] = fragW[i]; a1[i] = fragU[4 * US + i]; b1[i] = fragW[4 * WS + i]; }
    }
    double x = 1.0 + 1e-9 * lane;
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
        if constexpr (NMUL > 0) {
#pragma unroll
            for (int k = 0; k < NMUL; ++k) x = x * 1.000000001;
            if (x == 12345.0) Us[lane] = x;
        }
        if constexpr (SYNC) __syncwarp();
        if constexpr (SMEM) fetch(a0, b0, 0);
#pragma unroll 1
        for (int s = 0; s < 8; s += 2) {
            if constexpr (SMEM) fetch(a1, b1, s + 1);
            mma_step(a0, b0);
            if constexpr (SMEM) fetch(a0, b0, min(s + 2, 7));
            mma_step(a1, b1);
        }
        if constexpr (SYNC) __syncwarp();
    }
    const long long t1 = clock64();
    double s = x;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x < 148) cycles[blockIdx.x] = t1 - t0;
}

template <bool SMEM, int NMUL, bool SYNC>
void run(const char* name, int ctas_per_sm, int fill, double* sink, long long* d_cyc)
{
    const int groups = 2000;
    const size_t smem = 4 * kWarpDoubles * sizeof(double);
    cudaFuncSetAttribute(stream<SMEM, NMUL, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    stream<SMEM, NMUL, SYNC><<<148 * ctas_per_sm, 128, smem>>>(sink, groups, fill, d_cyc);
    cudaEventRecord(e0);
    stream<SMEM, NMUL, SYNC><<<148 * ctas_per_sm, 128, smem>>>(sink, groups, fill, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int i = 0; i < 148; ++i) cyc += double(h[i]) / 148.0;
    const double dmma_per_smsp = double(groups) * 128.0 * ctas_per_sm;  // one warp of each CTA per sub-partition
    const double tflops = 148.0 * 4.0 * dmma_per_smsp * 512.0 / (ms * 1e-3) * 1e-12;
    printf("%-34s warps/SMSP %d  fill %d   %8.3f ms  %6.2f TFLOP/s   %6.2f cycles/DMMA (clock64 of CTA 0..147: %.0f cycles, %.3f GHz)\n", name,
           ctas_per_sm, fill, ms, tflops, cyc / dmma_per_smsp, cyc, cyc / (ms * 1e-3) * 1e-9);
}

int main()
{
    double* sink; long long* d_cyc;
    cudaMalloc(&sink, 64); cudaMalloc(&d_cyc, 148 * sizeof(long long));
    for (int w = 1; w <= 3; ++w) run<false, 0, false>("registers only", w, 1, sink, d_cyc);
    for (int w = 1; w <= 3; ++w) run<true, 0, false>("fragments from shared memory", w, 1, sink, d_cyc);
    run<true, 0, false>("fragments from shared memory", 3, 0, sink, d_cyc);
    run<true, 0, false>("fragments from shared memory", 3, 2, sink, d_cyc);
    for (int w = 1; w <= 3; ++w) run<true, 0, true>("+ __syncwarp per group", w, 1, sink, d_cyc);
    for (int w = 1; w <= 3; ++w) run<true, 32, true>("+ 32 dependent DMUL per group", w, 1, sink, d_cyc);
    for (int w = 1; w <= 3; ++w) run<true, 96, true>("+ 96 dependent DMUL per group", w, 1, sink, d_cyc);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
