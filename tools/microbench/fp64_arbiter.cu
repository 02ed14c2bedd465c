// fp64_arbiter.cu -- how does an SM sub-partition share its FP64 pipe between a warp that streams DMMA.8x8x4
// (16 pipe cycles each, issued back to back) and warps that issue plain FP64 multiplies (2 pipe cycles each)?
// One CTA per SM; warp 0 streams N DMMAs, warps 4 and 8 (same sub-partition: warp % 4 == 0) run DMUL streams with
// a given number of independent chains until warp 0 raises a flag.  Reported: cycles per DMMA for warp 0, and DMULs
// each helper got in per DMMA.  If a helper gets about one DMUL per DMMA whatever its ILP, the arbiter alternates
// warps; if it gets almost none, the streaming warp is favoured; if it scales with ILP, the pipe is shared by need.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_arbiter fp64_arbiter.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void vmul(double& x, double q)
{
    asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(q));
}

struct Result { long long dmma_cycles; long long helper_muls[2]; };

template <int ILP>
__global__ void __launch_bounds__(288) arbiter(Result* out, int n_dmma_iters, int n_helpers, double seed, int n_streamers)
{
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0 || (warp == 4 && n_streamers == 2)) {
        double acc[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
        const double a = seed + lane, b = seed * 0.5;
        const long long t0 = clock64();
        for (int it = 0; it < n_dmma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], a, b);
        }
        const long long t1 = clock64();
        if (warp == 0) stop = 1;
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
        if (s == 12345.0) out[blockIdx.x].dmma_cycles = -1;
        else if (lane == 0 && warp == 0) out[blockIdx.x].dmma_cycles = t1 - t0;
    } else if ((warp == 8 && n_helpers >= 1) || (warp == 4 && n_helpers >= 2)) {
        double x[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = 1.0 + 1e-9 * (lane + k);
        const double q = 1.0 + seed * 1e-12;
        long long n = 0;
        while (!stop) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < ILP; ++k) vmul(x[k], q);
            n += 4 * ILP;
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < ILP; ++k) s += x[k];
        if (s == 12345.0) n = -1;
        if (lane == 0) out[blockIdx.x].helper_muls[warp == 8 ? 0 : 1] = n;
    }
}

template <int ILP>
void run(Result* d, int helpers, int streamers = 1)
{
    const int iters = 4000;
    cudaMemset(d, 0, 148 * sizeof(Result));
    arbiter<ILP><<<148, 288>>>(d, iters, helpers, 1.0, streamers);
    arbiter<ILP><<<148, 288>>>(d, iters, helpers, 1.0, streamers);
    cudaDeviceSynchronize();
    Result h[148];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0, m0 = 0, m1 = 0;
    for (int i = 0; i < 148; ++i) { cyc += double(h[i].dmma_cycles); m0 += double(h[i].helper_muls[0]); m1 += double(h[i].helper_muls[1]); }
    const double n_dmma = 148.0 * iters * 16;
    printf("streamers %d  helpers %d  ILP %2d   cycles/DMMA %6.2f   DMUL per DMMA (of warp 0): helper A %5.2f  helper B %5.2f\n", streamers, helpers, ILP, cyc / n_dmma,
           m0 / n_dmma, m1 / n_dmma);
}

int main()
{
    Result* d;
    cudaMalloc(&d, 148 * sizeof(Result));
    run<1>(d, 0);
    run<1>(d, 1); run<2>(d, 1); run<4>(d, 1); run<8>(d, 1); run<16>(d, 1);
    run<1>(d, 2); run<4>(d, 2); run<8>(d, 2);
    run<1>(d, 0, 2); run<1>(d, 1, 2); run<4>(d, 1, 2); run<8>(d, 1, 2);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
