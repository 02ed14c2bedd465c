// pcie_multi.cu -- host-link ceiling of one box: concurrent pinned D2H / H2D / duplex copies from N GPUs (one host thread,
// one pinned buffer and two streams per GPU, all started together), per-GPU and aggregate GB/s, for N = 1, 2, 4, 8 and for
// pairs of GPUs (which ones share an uplink).  This is what bounds tapir_cuda_run end to end at N > 1: a step copies
// 8 B x T per locus back to the host (the `net` table's profile), the kernels need a fraction of that time.
//   nvcc -O2 -std=c++17 -o pcie_multi pcie_multi.cu && ./pcie_multi [MB per copy] [repeats]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Gpu { int dev; void *h_out, *h_in, *d_out, *d_in; cudaStream_t s_out, s_in; };
static std::atomic<int> arrived{0}, go{0};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// mode 0: D2H, 1: H2D, 2: both
static void worker(Gpu* g, size_t bytes, int reps, int mode, int n, int round, double* gbs)
{
    CK(cudaSetDevice(g->dev));
    arrived.fetch_add(1);
    while (go.load() < round) std::this_thread::yield();
    const double t0 = now();
    for (int r = 0; r < reps; ++r) {
        if (mode != 1) CK(cudaMemcpyAsync(g->h_out, g->d_out, bytes, cudaMemcpyDeviceToHost, g->s_out));
        if (mode != 0) CK(cudaMemcpyAsync(g->d_in, g->h_in, bytes, cudaMemcpyHostToDevice, g->s_in));
    }
    CK(cudaStreamSynchronize(g->s_out));
    CK(cudaStreamSynchronize(g->s_in));
    const double dt = now() - t0;
    *gbs = double(bytes) * reps / dt / 1e9;   // per direction
    (void)n;
}

int main(int argc, char** argv)
{
    const size_t bytes = size_t(argc > 1 ? atoi(argv[1]) : 256) << 20;
    const int reps = argc > 2 ? atoi(argv[2]) : 8;
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    std::vector<Gpu> gpus(n_dev);
    for (int i = 0; i < n_dev; ++i) {
        Gpu& g = gpus[i];
        g.dev = i;
        CK(cudaSetDevice(i));
        CK(cudaHostAlloc(&g.h_out, bytes, cudaHostAllocPortable));
        CK(cudaHostAlloc(&g.h_in, bytes, cudaHostAllocPortable));
        CK(cudaMalloc(&g.d_out, bytes));
        CK(cudaMalloc(&g.d_in, bytes));
        CK(cudaMemset(g.d_out, 1, bytes));
        CK(cudaStreamCreateWithFlags(&g.s_out, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g.s_in, cudaStreamNonBlocking));
        CK(cudaDeviceSynchronize());
    }
    printf("# %d GPUs, %zu MB per copy, %d copies per direction; GB/s per direction\n", n_dev, bytes >> 20, reps);
    int round = 0;
    auto run = [&](const std::vector<int>& set, int mode) {
        std::vector<double> gbs(set.size(), 0.0);
        std::vector<std::thread> th;
        arrived = 0;
        ++round;
        for (size_t k = 0; k < set.size(); ++k) th.emplace_back(worker, &gpus[set[k]], bytes, reps, mode, int(set.size()), round, &gbs[k]);
        while (arrived.load() < int(set.size())) std::this_thread::yield();
        go = round;
        for (auto& t : th) t.join();
        double sum = 0, mn = 1e30;
        for (double v : gbs) { sum += v; mn = v < mn ? v : mn; }
        printf("%-7s GPUs {", mode == 0 ? "D2H" : mode == 1 ? "H2D" : "duplex");
        for (size_t k = 0; k < set.size(); ++k) printf("%s%d", k ? "," : "", set[k]);
        printf("}: per GPU min %.1f, aggregate %.1f   [", mn, sum);
        for (size_t k = 0; k < set.size(); ++k) printf("%s%.1f", k ? " " : "", gbs[k]);
        printf("]\n");
    };
    for (int n = 1; n <= n_dev; n *= 2) {
        std::vector<int> set;
        for (int i = 0; i < n; ++i) set.push_back(i);
        for (int mode = 0; mode < 3; ++mode) run(set, mode);
    }
    // pairs: which GPUs share an uplink
    for (int j = 1; j < n_dev; ++j) run({0, j}, 0);
    if (n_dev >= 4) { run({0, 2, 4, 6}, 0); run({0, 1, 2, 3}, 0); }
    return 0;
}
