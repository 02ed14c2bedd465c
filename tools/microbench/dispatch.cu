// dispatch.cu -- do non-FP64 instructions hide in the shadow of the half-rate FP64 pipe on B200?
// Each loop iteration issues 64 independent DFMAs (or 16 DMMAs) plus NI independent integer IMADs (and, in a
// second variant, NI shared-memory loads).  If time does not grow with NI the integer work is free.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dispatch dispatch.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, bool LDS>
__global__ void __launch_bounds__(128) dfma_int(double* sink, int iters, const double* in, int mul)
{
    __shared__ int sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i;
    __syncthreads();
    double acc[8][4], u[8], w[4];
    int z[NI > 0 ? NI : 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = in[i] + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = in[16 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;
#pragma unroll
    for (int k = 0; k < NI; ++k) z[k] = threadIdx.x + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = fma(u[i], w[j], acc[i][j]);
                constexpr int per = (NI + 31) / 32;
#pragma unroll
                for (int k = 0; k < per; ++k) {
                    const int q = (i * 4 + j) * per + k;
                    if (q < NI) {
                        if constexpr (LDS) z[q] = sh[(z[q] + it) & 1023];
                        else z[q] = z[q] * mul + it;
                    }
                }
            }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(u[i], w[j], acc[i][j]);
    }
    double s = 0;
    int zi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
#pragma unroll
    for (int k = 0; k < NI; ++k) zi += z[k];
    if (s == 123.456 || zi == 0x7fffffff) sink[0] = s + zi;
}

template <class F>
float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *sink, *in;
    cudaMalloc(&sink, 64); cudaMalloc(&in, 64 * 8);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + i * 1e-3;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 4096, blocks = sms * 3, threads = 128;   // 12 warps / SM
    printf("128 DFMA per iteration per thread, 12 warps/SM; extra ops per iteration -> ms\n");
#define RUN(NI, L) printf("%-6s NI=%3d  %.4f ms\n", L ? "lds" : "imad", NI, time_ms([&] { dfma_int<NI, L><<<blocks, threads>>>(sink, iters, in, 3); }))
    RUN(0, false); RUN(16, false); RUN(32, false); RUN(64, false); RUN(128, false);
    RUN(16, true); RUN(32, true); RUN(64, true);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
