// pcie_bench.cu -- does splitting a 1 GiB pinned copy over several streams (copy engines) beat one cudaMemcpyAsync?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/pcie_bench bench_tools/pcie_bench.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    const size_t bytes = (size_t)1 << 30;
    char *h, *d;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d, bytes));
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (char)i;
    cudaStream_t st[8];
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (int dir = 0; dir < 2; ++dir) {
        for (int parts : {1, 2, 4, 8}) {
            double best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaDeviceSynchronize());
                const double t0 = now();
                const size_t chunk = bytes / parts;
                for (int p = 0; p < parts; ++p) {
                    if (dir == 0) CK(cudaMemcpyAsync(d + p * chunk, h + p * chunk, chunk, cudaMemcpyHostToDevice, st[p]));
                    else CK(cudaMemcpyAsync(h + p * chunk, d + p * chunk, chunk, cudaMemcpyDeviceToHost, st[p]));
                }
                CK(cudaDeviceSynchronize());
                const double t = now() - t0;
                if (rep > 0 && t < best) best = t;
            }
            printf("%s 1 GiB in %d part(s) on %d stream(s): %.2f ms = %.1f GB/s\n", dir == 0 ? "H2D" : "D2H", parts, parts, best * 1e3,
                   bytes / best / 1e9);
        }
    }
    // both directions at once (what two pipelined host-buffer sorts do)
    char *h2, *d2;
    CK(cudaHostAlloc(&h2, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d2, bytes));
    double best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaDeviceSynchronize());
        const double t0 = now();
        CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st[0]));
        CK(cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, st[1]));
        CK(cudaDeviceSynchronize());
        const double t = now() - t0;
        if (rep > 0 && t < best) best = t;
    }
    printf("H2D + D2H of 1 GiB each at once: %.2f ms = %.1f GB/s per direction\n", best * 1e3, bytes / best / 1e9);
    return 0;
}
