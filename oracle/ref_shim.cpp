/*
 * ref_shim.cpp -- extern "C" doorway onto the UNMODIFIED reference build.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/lsd_oracle.c header).  oracle/Makefile
 * compiles /root/reference/LSDRadixSort/{LSDRadixSort.cu,Utils.cpp,CudaUtils.cpp}
 * in place (with -Dmain=lsd_ref_main so the benchmark driver is inert) and links
 * this file against those objects into oracle/_ref/libref_lsd.so.  No reference
 * source is copied into the repo: the reference's algorithms have external C++
 * linkage, so declaring their prototypes here is enough to call them.
 *
 * Prototypes restated from the reference (LSDRadixSort.cu): :62 LSDRadixSort,
 * :25 LSDRadixSortPass, :128 PrefixSum, :643 BuildHistogramsCPU,
 * :265 GetGPUPrefixSumBlockSumsCount, :286 GPUPrefixSum, :660
 * BuildHistogramsKernel, :839 GPULSDRadixSort.
 */
#include <cstdint>
#include <cuda_runtime.h>

void LSDRadixSortPass(uint32_t* in, uint32_t* out, int count, uint32_t* histogram, int r, int bit_group);
void LSDRadixSort(uint32_t* in, uint32_t* out, int count, uint32_t* histogram, int r);
void PrefixSum(uint32_t* a, int count);
void BuildHistogramsCPU(uint32_t* a, uint32_t* h, int count, int r, int bit_group, int grid, int block);
int GetGPUPrefixSumBlockSumsCount(int count, int threads_per_block);
void GPUPrefixSum(uint32_t* d_a, int count, int threads_per_block, uint32_t* d_block_sums, cudaStream_t s);
__global__ void BuildHistogramsKernel(uint32_t* a, uint32_t* h, int count, int r, int bit_group);
void GPULSDRadixSort(uint32_t* a, uint32_t* b, uint32_t* h, uint32_t* block_sums, uint32_t* d, int grid, int block,
                     int block_sums_count, int count, int h_count, int r);

#define REF_API extern "C" __attribute__((visibility("default")))

/* ---- CPU path (runs anywhere) ------------------------------------------------ */
REF_API void ref_cpu_sort_pass(uint32_t* in, uint32_t* out, int count, uint32_t* histogram, int r, int bit_group)
{
    LSDRadixSortPass(in, out, count, histogram, r, bit_group);
}
REF_API void ref_cpu_sort(uint32_t* in, uint32_t* out, int count, uint32_t* histogram, int r)
{
    LSDRadixSort(in, out, count, histogram, r);
}
REF_API void ref_cpu_prefix_sum(uint32_t* a, int count) { PrefixSum(a, count); }
REF_API void ref_cpu_build_histograms(uint32_t* a, uint32_t* h, int count, int r, int bit_group, int grid, int block)
{
    BuildHistogramsCPU(a, h, count, r, bit_group, grid, block);
}
REF_API int ref_block_sums_count(int count, int threads_per_block)
{
    return GetGPUPrefixSumBlockSumsCount(count, threads_per_block);
}

/* ---- GPU path (B200 box only; reported baseline, never the product) ---------- */
REF_API int ref_gpu_prefix_sum(uint32_t* d_a, int count, int threads_per_block, uint32_t* d_block_sums)
{
    GPUPrefixSum(d_a, count, threads_per_block, d_block_sums, 0);
    return (int)cudaGetLastError();
}
REF_API int ref_gpu_build_histograms(uint32_t* d_a, uint32_t* d_h, int count, int r, int bit_group, int block)
{
    const int grid = (count + block - 1) / block;
    const size_t smem = sizeof(uint32_t) << r;
    BuildHistogramsKernel<<<grid, block, smem>>>(d_a, d_h, count, r, bit_group);
    return (int)cudaGetLastError();
}
/* Caller sizes h (3*grid*2^r words) and block_sums exactly as TestGPULSDRadixSort does (:919-929). */
REF_API int ref_gpu_sort(uint32_t* d_a, uint32_t* d_b, uint32_t* d_h, uint32_t* d_block_sums, int count, int block, int r)
{
    const int grid = (count + block - 1) / block;
    const int h_count = 1 << r;
    GPULSDRadixSort(d_a, d_b, d_h, d_block_sums, nullptr, grid, block, 0, count, h_count, r);
    return (int)cudaGetLastError();
}
REF_API int ref_device_synchronize() { return (int)cudaDeviceSynchronize(); }
