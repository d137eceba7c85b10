// common.cuh -- shared device/host helpers for liblsdsort (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lsdsort.h"

namespace lsd {

constexpr int kWarp = 32;
constexpr uint32_t kFullMask = 0xFFFFFFFFu;

// Digit `bit_group` of width RB (0 = least significant).  Same digit order as the
// reference's GET_R_BITS (Utils.h:22).
template <int RB>
__host__ __device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift)
{
    return (key >> shift) & ((1u << RB) - 1u);
}

// Order-preserving bijections between typed keys and unsigned order (lsd_key_type), branch-free and driven by two
// masks: (flip, sign) = (0, 0) u32 -- identity; (0, 0x80000000) i32; (0x80000000, 0x80000000) f32.
struct KeyXform {
    uint32_t flip;  // bit whose value selects "flip everything"
    uint32_t sign;  // bits that are always flipped
};
__host__ __device__ __forceinline__ KeyXform key_xform_of(uint32_t key_type)
{
    return KeyXform{key_type == 2u ? 0x80000000u : 0u, key_type != 0u ? 0x80000000u : 0u};
}
__host__ __device__ __forceinline__ uint32_t key_to_unsigned(uint32_t k, KeyXform x)
{
    return k ^ ((uint32_t)((int32_t)(k & x.flip) >> 31) | x.sign);
}
__host__ __device__ __forceinline__ uint32_t key_from_unsigned(uint32_t u, KeyXform x)
{
    return u ^ ((uint32_t)((int32_t)(~u & x.flip) >> 31) | x.sign);
}

__device__ __forceinline__ uint32_t lane_id()
{
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t lanemask_gt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}

// Streaming 128-bit / 32-bit global loads that do not allocate in L1 (keys are read once).
__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Look-back words: flag and value share one 32-bit (or 64-bit) word, so a relaxed
// gpu-scope store/load is all the ordering the protocol needs.
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ---- host-side error plumbing -------------------------------------------------------
void set_last_cuda_error(cudaError_t e);

#define LSD_CUDA_TRY(expr)                         \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) {                   \
            ::lsd::set_last_cuda_error(_e);        \
            return LSD_ERR_CUDA;                   \
        }                                          \
    } while (0)

#define LSD_LAUNCH_CHECK() LSD_CUDA_TRY(cudaGetLastError())

inline bool valid_radix(int r) { return r == 1 || r == 2 || r == 4 || r == 8; }  // digit widths with their own pass kernels
// Composite digit widths: every other r up to 16 (16 is the one factor of 32 the reference's CPU path takes beyond these,
// LSDRadixSort.cu:56-69; 11 is the classic 11-11-10 split).  Digit i is bits [i*r, min(32, (i+1)*r)); there are
// ceil(32/r) of them.  A stable pass on such a digit IS a stable pass on its low 8 bits followed by one on the rest, so
// these widths run on the 8-bit pass kernel (sort.cu: pass_enqueue_wide); a full sort's result does not depend on r at
// all and runs the 8-bit schedule.
inline bool composite_radix(int r) { return r >= 3 && r <= 16 && !valid_radix(r); }
inline bool accepted_radix(int r) { return valid_radix(r) || composite_radix(r); }
inline int exec_radix(int r) { return composite_radix(r) ? 8 : r; }   // the digit width the full sort executes
inline int digit_count(int r) { return (32 + r - 1) / r; }
inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int sm_count();          // cached multiprocessor count of the current device
int smem_optin_bytes();  // cached max opt-in shared memory per block

// ---- entry points implemented per translation unit (called by api.cu) ----------------
// zero_ptr / zero_bytes (16-byte granular, optional): memory the kernel zeroes on the side (the sort's look-back records)
int launch_digit_histograms(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s,
                            uint32_t key_type = 0, void* zero_ptr = nullptr, size_t zero_bytes = 0);
int launch_top_digit_histogram(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s);
int launch_one_digit_histogram(const uint32_t* keys, uint64_t n, int r, int digit, uint64_t* hist, cudaStream_t s);
int launch_tile_histograms(const uint32_t* keys, uint64_t n, int r, int bit_group, int block, uint32_t* hist,
                           cudaStream_t s);
int launch_field_histogram(const uint32_t* keys, uint64_t n, int shift, int bits, uint64_t* hist, cudaStream_t s);
int launch_field_scan(uint64_t* a, int bits, cudaStream_t s);
size_t scan_workspace_bytes(uint64_t n, int block);
int launch_prefix_sum(uint32_t* a, uint64_t n, int block, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace lsd
