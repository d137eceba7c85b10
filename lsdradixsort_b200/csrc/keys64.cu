// keys64.cu -- lsd_sort64: LSD radix sort of 64-bit keys (uint64 / int64 / float64 order).
//
// No reference counterpart: the reference sorts uint32 only (LSDRadixSort.cu:62, :839); this is SURVEY 8(f)4, the
// width extension of the same pass.  An LSD sort of 64-bit keys with 8-bit digits is eight stable passes; a stable pass
// on a digit of the LOW word moves (low, high) together with the low word as the key, a pass on a digit of the HIGH
// word the other way round.  Both are exactly the key-value pass the library already has (16 B per pair per pass =
// 16 B per 64-bit key per pass, the same traffic a native 64-bit pass would move), so the sort is:
//     split   keys[i] -> lo[i], hi[i]                       (structure of arrays, typed keys mapped to unsigned order)
//     pairs   sort (key = lo, value = hi)  -- digits 0..3, stable
//     pairs   sort (key = hi, value = lo)  -- digits 4..7, stable
//     merge   keys[i] <- hi[i] << 32 | lo[i]                (mapped back)
// Each pairs sort keeps its own device-side plan, so constant digits are skipped per word (keys below 2^32 cost the four
// low passes only).  The caller's key buffer is the ping-pong space of the pairs sorts, `scratch` holds lo / hi.
// The key-value passes stage tiles with TMA and need 16-byte aligned arrays, so the split covers the first m = n & ~3
// keys (four word arrays of m entries tile keys + scratch exactly); the n - m <= 3 tail keys wait in the workspace and
// the merge kernel interleaves them (a two-way merge where one side has at most three elements).
#include "sort.h"

namespace lsd {

// unsigned-order image of a 64-bit key: type 3 = uint64 (identity), 4 = int64 (flip the sign bit),
// 5 = float64 in IEEE total order (negatives: flip everything, others: flip the sign bit)
__host__ __device__ __forceinline__ uint64_t key64_to_unsigned(uint64_t k, uint32_t key_type)
{
    if (key_type == LSD_KEY_I64) return k ^ 0x8000000000000000ull;
    if (key_type == LSD_KEY_F64) return k ^ ((uint64_t)((int64_t)k >> 63) | 0x8000000000000000ull);
    return k;
}
__host__ __device__ __forceinline__ uint64_t key64_from_unsigned(uint64_t u, uint32_t key_type)
{
    if (key_type == LSD_KEY_I64) return u ^ 0x8000000000000000ull;
    if (key_type == LSD_KEY_F64) return u ^ ((uint64_t)((int64_t)~u >> 63) | 0x8000000000000000ull);
    return u;
}

constexpr int kSplitThreads = 256;

// lo / hi: m entries each (m % 4 == 0); tail: the n - m keys behind them (unsigned images, ascending).
// Two keys per thread and step: one 128-bit load, two 64-bit stores.
__global__ void __launch_bounds__(kSplitThreads)
split64_kernel(const uint64_t* __restrict__ keys, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi, uint64_t m, uint64_t n,
               uint64_t* __restrict__ tail, uint32_t key_type)
{
    const uint64_t stride = (uint64_t)gridDim.x * kSplitThreads;
    const uint64_t pairs = m >> 1;
    for (uint64_t i = (uint64_t)blockIdx.x * kSplitThreads + threadIdx.x; i < pairs; i += stride) {
        const ulonglong2 k = __ldcs(reinterpret_cast<const ulonglong2*>(keys) + i);
        const uint64_t u0 = key64_to_unsigned(k.x, key_type), u1 = key64_to_unsigned(k.y, key_type);
        reinterpret_cast<uint2*>(lo)[i] = make_uint2((uint32_t)u0, (uint32_t)u1);
        reinterpret_cast<uint2*>(hi)[i] = make_uint2((uint32_t)(u0 >> 32), (uint32_t)(u1 >> 32));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t t[3] = {~0ull, ~0ull, ~0ull};
        const uint32_t cnt = (uint32_t)(n - m);
        for (uint32_t j = 0; j < cnt; ++j) t[j] = key64_to_unsigned(keys[m + j], key_type);
        // three-element sorting network (unused slots hold the maximum and stay last)
        if (t[0] > t[1]) { const uint64_t x = t[0]; t[0] = t[1]; t[1] = x; }
        if (t[1] > t[2]) { const uint64_t x = t[1]; t[1] = t[2]; t[2] = x; }
        if (t[0] > t[1]) { const uint64_t x = t[0]; t[0] = t[1]; t[1] = x; }
        for (uint32_t j = 0; j < 3; ++j) tail[j] = t[j];
    }
}

// Two-way merge of the m sorted (hi, lo) keys with the <= 3 sorted tail keys: main key i goes to i + #{tail < key},
// tail key j to j + #{main <= tail_j} (ties: main first; equal 64-bit keys are indistinguishable).  Two keys per thread
// and step; without a tail (n % 4 == 0) they leave as one 128-bit store.
__global__ void __launch_bounds__(kSplitThreads)
merge64_kernel(uint64_t* __restrict__ keys, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, uint64_t m, uint64_t n,
               const uint64_t* __restrict__ tail, uint32_t key_type)
{
    const uint32_t cnt = (uint32_t)(n - m);
    const uint64_t t0 = cnt > 0 ? tail[0] : 0, t1 = cnt > 1 ? tail[1] : 0, t2 = cnt > 2 ? tail[2] : 0;
    const uint64_t stride = (uint64_t)gridDim.x * kSplitThreads;
    const uint64_t pairs = m >> 1;
    for (uint64_t i = (uint64_t)blockIdx.x * kSplitThreads + threadIdx.x; i < pairs; i += stride) {
        const uint2 l = __ldcs(reinterpret_cast<const uint2*>(lo) + i);
        const uint2 h = __ldcs(reinterpret_cast<const uint2*>(hi) + i);
        const uint64_t u0 = ((uint64_t)h.x << 32) | l.x, u1 = ((uint64_t)h.y << 32) | l.y;
        if (cnt == 0) {
            reinterpret_cast<ulonglong2*>(keys)[i] = make_ulonglong2(key64_from_unsigned(u0, key_type), key64_from_unsigned(u1, key_type));
        } else {
            const uint32_t b0 = (t0 < u0 ? 1u : 0u) + (cnt > 1 && t1 < u0 ? 1u : 0u) + (cnt > 2 && t2 < u0 ? 1u : 0u);
            const uint32_t b1 = (t0 < u1 ? 1u : 0u) + (cnt > 1 && t1 < u1 ? 1u : 0u) + (cnt > 2 && t2 < u1 ? 1u : 0u);
            keys[2 * i + b0] = key64_from_unsigned(u0, key_type);
            keys[2 * i + 1 + b1] = key64_from_unsigned(u1, key_type);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < cnt) {
        const uint64_t t = tail[threadIdx.x];
        uint64_t a = 0, b = m;  // upper bound: first main key > t
        while (a < b) {
            const uint64_t mid = a + ((b - a) >> 1);
            const uint64_t u = ((uint64_t)hi[mid] << 32) | lo[mid];
            if (u <= t) a = mid + 1; else b = mid;
        }
        keys[a + threadIdx.x] = key64_from_unsigned(t, key_type);
    }
}

static size_t align256_(size_t v) { return (v + 255) / 256 * 256; }

size_t sort64_workspace_bytes(uint64_t n)
{
    const uint64_t m = n & ~3ull;
    SortLayout L;
    if (make_layout(m, 8, 0, nullptr, &L, true) != LSD_OK) return 0;
    return align256_(L.total_bytes) + 256;
}

int sort64_enqueue(uint64_t* keys, uint64_t* scratch, uint64_t n, uint32_t key_type, void* ws, size_t ws_bytes, cudaStream_t s)
{
    if (key_type < LSD_KEY_U64 || key_type > LSD_KEY_F64) return LSD_ERR_INVALID_VALUE;
    const uint64_t m = n & ~3ull;
    SortLayout L;
    const int st = make_layout(m, 8, 0, nullptr, &L, true);
    if (st != LSD_OK) return st;
    if (n == 0) return LSD_OK;
    if (!keys || !scratch || !ws) return LSD_ERR_INVALID_VALUE;
    const size_t pairs_ws = align256_(L.total_bytes);
    if (ws_bytes < pairs_ws + 256) return LSD_ERR_WORKSPACE_TOO_SMALL;
    if (!aligned_to(keys, 16) || !aligned_to(scratch, 16) || !aligned_to(ws, 256)) return LSD_ERR_ALIGNMENT;

    uint32_t* lo = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* hi = lo + m;  // m % 4 == 0: 16-byte aligned
    uint32_t* pp0 = reinterpret_cast<uint32_t*>(keys);
    uint32_t* pp1 = pp0 + m;
    uint64_t* tail = reinterpret_cast<uint64_t*>(static_cast<char*>(ws) + pairs_ws);

    const uint64_t want = ((m >> 1) + kSplitThreads - 1) / kSplitThreads;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    split64_kernel<<<grid, kSplitThreads, 0, s>>>(keys, lo, hi, m, n, tail, key_type);
    LSD_LAUNCH_CHECK();
    if (m > 0) {
        int rc = sort_enqueue(lo, pp0, m, 8, 0, ws, pairs_ws, nullptr, s, nullptr, nullptr, hi, pp1);
        if (rc != LSD_OK) return rc;
        rc = sort_enqueue(hi, pp0, m, 8, 0, ws, pairs_ws, nullptr, s, nullptr, nullptr, lo, pp1);
        if (rc != LSD_OK) return rc;
    }
    merge64_kernel<<<grid, kSplitThreads, 0, s>>>(keys, lo, hi, m, n, tail, key_type);
    LSD_LAUNCH_CHECK();
    return LSD_OK;
}

}  // namespace lsd
