// pcie_duplex.cu -- can host-to-device and device-to-host copies of pinned memory overlap on this box?
// Times 20 MB H2D alone, 40 MB D2H alone, and both at once on two streams (the per-step traffic of config 2).
// build: nvcc -O3 -o pcie_duplex pcie_duplex.cu
#include <cstdio>
#include <cuda_runtime.h>

int main()
{
    const size_t in_bytes = 20u << 20, out_bytes = 40u << 20;
    void *h_in, *h_out, *d_in, *d_out;
    cudaMallocHost(&h_in, in_bytes); cudaMallocHost(&h_out, out_bytes);
    cudaMalloc(&d_in, in_bytes); cudaMalloc(&d_out, out_bytes);
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    auto run = [&](bool in, bool out) {
        float best = 1e30f;
        for (int r = 0; r < 6; ++r) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0, 0);
            if (in) cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s1);
            if (out) cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s2);
            cudaEventRecord(e2, s2);
            cudaStreamWaitEvent(s1, e2, 0);
            cudaEventRecord(e1, s1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        return best;
    };
    const float a = run(true, false), b = run(false, true), c = run(true, true);
    printf("H2D 20 MB alone      %.3f ms  (%.1f GB/s)\n", a, in_bytes / a / 1e6);
    printf("D2H 40 MB alone      %.3f ms  (%.1f GB/s)\n", b, out_bytes / b / 1e6);
    printf("both, two streams    %.3f ms  (sum of the two alone: %.3f, max: %.3f)\n", c, a + b, a > b ? a : b);
    return cudaDeviceSynchronize() != cudaSuccess;
}
