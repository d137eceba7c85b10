// multi.cu -- lsd_sort_multi: the multi-GPU sort behind the C ABI (one process or thread per GPU, one node).
//
// No reference counterpart: the reference is single-GPU.  This is BASELINE.json's partitioning
// (SURVEY 8(e)): digit histograms per rank -> the rows are all-gathered (their sum is the all-reduced MSD histogram)
// -> every rank derives the same contiguous bucket -> rank map -> one pass kernel partitions the local keys by
// destination and stores them straight into the owners' receive buffers over NVLink peer memory (CUDA IPC) -> local
// LSD sort of what arrived.  The map is planned ON THE DEVICE (multi_plan_kernel); the host waits for 64 bytes of its
// result (which digit to partition on, this rank's share) and enqueues the exchange pass and the local sort.
// Skew: the exchange partitions on the 8-bit WINDOW THAT ENDS AT THE HIGHEST BIT THAT VARIES over the whole input (all bits
// above it are constant, so contiguous bucket ranges of the window are contiguous key ranges): keys in any range -- below
// 2^24, in [0, 2^17), in [2^31, 2^31 + 1000) -- are balanced over their own 8 most significant varying bits instead of
// landing on one rank, and an input whose keys are all equal stays where it is.
//
// The two collectives (a 2 KiB all-gather and a barrier) come from the caller as callbacks, so liblsdsort has no link
// dependency on a communication library; include/lsdsort_nccl.h supplies them for an ncclComm_t, lsdradixsort_b200/
// multi.py for torch.distributed.
#include <new>
#include <vector>

#include <cstring>
#include <unistd.h>

#include "sort.h"

namespace lsd {

constexpr int kMultiBuckets = 256;  // the exchange partitions on one 8-bit digit
constexpr int kMultiDigits = 4;
constexpr int kMultiMaxRanks = 64;

struct MultiResult {              // written by multi_plan_kernel, copied to the host
    uint64_t n_out;               // keys this rank owns after the exchange
    uint64_t n_out_max;           // the largest share of any rank
    uint64_t sent;                // keys this rank sends to other ranks
    uint32_t overflow;            // some rank's share exceeds the capacity: nothing is moved
    uint32_t first_bucket;        // this rank's bucket range [first, last]; first > last when it owns nothing
    uint32_t last_bucket;
    uint32_t shift;               // the exchange partitions on bits [shift, shift + 8): the window that ends at the highest
                                  // bit that varies over the input (24 = the top digit)
    uint32_t keep_local;          // every key of every rank is the same: nothing to exchange, each rank keeps its keys
    uint32_t need_window_hist;    // shift is not a multiple of 8: the map needs the histogram of that window first
    uint32_t pad[4];
};
static_assert(sizeof(MultiResult) == 64, "copied as 64 bytes");

// Bucket -> rank map for one 8-bit window.  rows: the window's histogram of rank s is rows[s * stride + b].
// Bucket b goes to rank floor(nranks * (keys before b + half of b) / total), made monotone: every rank owns a contiguous
// run of buckets whose total is as close to total / nranks as whole buckets allow (the same rule as multi.assign_buckets).
// Called by all 256 threads of the plan CTA; tot = this thread's bucket total, grand = number of keys (> 0).
__device__ void multi_plan_map(const uint64_t* __restrict__ rows, size_t stride, uint64_t tot, uint64_t grand, int nranks, int rank,
                               const uint64_t* __restrict__ peer_ptrs, uint64_t capacity, uint64_t* __restrict__ dst_ptrs,
                               uint32_t* __restrict__ dst_seg, uint32_t* __restrict__ abort_flag, MultiResult* __restrict__ result,
                               uint32_t shift)
{
    __shared__ uint64_t s_scan[kMultiBuckets];
    __shared__ int s_owner[kMultiBuckets];
    __shared__ uint64_t s_share[kMultiMaxRanks];   // keys owned by rank d
    __shared__ uint64_t s_before[kMultiMaxRanks];  // keys that sources ahead of me put into d's buffer
    __shared__ uint64_t s_mine[kMultiMaxRanks];    // keys I send to d
    __shared__ uint32_t s_first[kMultiMaxRanks], s_last[kMultiMaxRanks];
    const int b = threadIdx.x;
    s_scan[b] = tot;
    if (b < kMultiMaxRanks) {
        s_share[b] = 0;
        s_before[b] = 0;
        s_mine[b] = 0;
        s_first[b] = kMultiBuckets;
        s_last[b] = 0;
    }
    __syncthreads();
    for (int o = 1; o < kMultiBuckets; o <<= 1) {  // inclusive scan of the global histogram
        const uint64_t t = b >= o ? s_scan[b - o] : 0;
        __syncthreads();
        s_scan[b] += t;
        __syncthreads();
    }
    int owner;
    {
        const uint64_t twice_mid = 2 * (s_scan[b] - tot) + tot;  // < 2^41: the product below fits 64 bits for nranks <= 64
        const uint64_t o = twice_mid * (uint64_t)nranks / (2 * grand);
        owner = (int)(o < (uint64_t)(nranks - 1) ? o : (uint64_t)(nranks - 1));
    }
    s_owner[b] = owner;
    __syncthreads();
    // monotone even with empty buckets: running maximum
    for (int o = 1; o < kMultiBuckets; o <<= 1) {
        const int t = b >= o ? s_owner[b - o] : 0;
        __syncthreads();
        if (t > s_owner[b]) s_owner[b] = t;
        __syncthreads();
    }
    owner = s_owner[b];
    // per destination: total share, what the sources ahead of me send, what I send, its bucket range
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_share[owner]), (unsigned long long)tot);
    uint64_t ahead = 0;
    for (int s = 0; s < rank; ++s) ahead += rows[s * stride + b];
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_before[owner]), (unsigned long long)ahead);
    atomicAdd(reinterpret_cast<unsigned long long*>(&s_mine[owner]), (unsigned long long)rows[(size_t)rank * stride + b]);
    atomicMin(&s_first[owner], (uint32_t)b);
    atomicMax(&s_last[owner], (uint32_t)b);
    __syncthreads();
    // all buckets owned by one rank form ONE destination segment: the pass kernel appends the segment's keys of a tile
    // as one block to the owner's buffer, where the block of source `rank` starts behind the blocks of the ranks ahead
    dst_ptrs[b] = peer_ptrs[owner] + 4ull * s_before[owner];
    dst_seg[b] = s_first[owner] | (s_last[owner] << 16);
    if (b == 0) {
        uint64_t mx = 0, sent = 0;
        for (int d = 0; d < nranks; ++d) {
            mx = s_share[d] > mx ? s_share[d] : mx;
            if (d != rank) sent += s_mine[d];
        }
        result->n_out = s_share[rank];
        result->n_out_max = mx;
        result->sent = sent;
        result->overflow = mx > capacity ? 1u : 0u;
        result->first_bucket = s_first[rank];
        result->last_bucket = s_last[rank];
        result->shift = shift;
        result->keep_local = 0;
        result->need_window_hist = 0;
        *abort_flag = mx > capacity ? 1u : 0u;
    }
}

// One CTA of 256 threads (one per bucket).  per_rank: [nranks][4][256] digit counts (the all-gathered histograms).
// Finds the highest bit h that varies over the whole input (from the highest digit whose global histogram is not a single
// bucket: the bits in which its live bucket numbers differ) and the exchange window [shift, shift + 8), shift = max(0, h - 7).
// All bits above the window are constant, so contiguous bucket ranges of the window are contiguous key ranges, and the
// window holds the 8 most significant varying bits: 256-way balance whatever the range of the keys.  A window that is a
// digit (shift = 0, 8, 16, 24) is mapped at once; otherwise the caller histograms the window and runs multi_window_kernel.
__global__ void __launch_bounds__(kMultiBuckets)
multi_plan_kernel(const uint64_t* __restrict__ per_rank, int nranks, int rank, const uint64_t* __restrict__ peer_ptrs,
                  uint64_t capacity, uint64_t* __restrict__ dst_ptrs, uint32_t* __restrict__ dst_seg,
                  uint32_t* __restrict__ abort_flag, MultiResult* __restrict__ result)
{
    __shared__ uint64_t s_red[kMultiBuckets];
    __shared__ uint32_t s_or, s_and;
    const int b = threadIdx.x;
    constexpr size_t kRow = (size_t)kMultiDigits * kMultiBuckets;  // one rank's histograms
    // total number of keys (any digit's histogram sums to it)
    uint64_t tot = 0;
    for (int s = 0; s < nranks; ++s) tot += per_rank[s * kRow + (size_t)(kMultiDigits - 1) * kMultiBuckets + b];
    s_red[b] = tot;
    if (b == 0) { s_or = 0u; s_and = 0xFFu; }
    __syncthreads();
    for (int o = kMultiBuckets / 2; o > 0; o >>= 1) {
        if (b < o) s_red[b] += s_red[b + o];
        __syncthreads();
    }
    const uint64_t grand = s_red[0];
    // the highest digit that varies
    int digit = kMultiDigits - 1;
    int constant = 1;
    for (int p = kMultiDigits - 1; p >= 0; --p) {
        uint64_t t = 0;
        for (int s = 0; s < nranks; ++s) t += per_rank[s * kRow + (size_t)p * kMultiBuckets + b];
        constant = __syncthreads_or(t == grand);  // one bucket holds every key (also true for an empty input)
        if (!constant) {
            digit = p;
            tot = t;
            break;
        }
    }
    if (constant) {
        // all keys are equal (or there are none): every rank keeps what it has
        dst_ptrs[b] = peer_ptrs[rank];
        dst_seg[b] = 0u | ((uint32_t)(kMultiBuckets - 1) << 16);
        if (b == 0) {
            uint64_t mine = 0, mx = 0;
            for (int s = 0; s < nranks; ++s) {
                uint64_t ns = 0;
                for (int q = 0; q < kMultiBuckets; ++q) ns += per_rank[s * kRow + q];
                mx = ns > mx ? ns : mx;
                if (s == rank) mine = ns;
            }
            result->n_out = mine;
            result->n_out_max = mx;
            result->sent = 0;
            result->overflow = mx > capacity ? 1u : 0u;
            result->first_bucket = 0;
            result->last_bucket = kMultiBuckets - 1;
            result->shift = 0;
            result->keep_local = 1;
            result->need_window_hist = 0;
            *abort_flag = 1u;
        }
        return;
    }
    // bits of this digit in which the live buckets differ
    if (tot != 0) {
        atomicOr(&s_or, (uint32_t)b);
        atomicAnd(&s_and, (uint32_t)b);
    }
    __syncthreads();
    const uint32_t varying = s_or & ~s_and;                       // != 0: at least two live buckets
    const int h = 8 * digit + (31 - __clz((int)varying));         // highest varying bit of the keys
    const uint32_t shift = h >= 7 ? (uint32_t)(h - 7) : 0u;
    if ((shift & 7u) != 0u) {  // the window straddles two digits: its histogram is not among the four
        if (b == 0) {
            result->shift = shift;
            result->keep_local = 0;
            result->need_window_hist = 1;
            result->overflow = 0;
            result->n_out = result->n_out_max = result->sent = 0;
        }
        return;
    }
    const int wd = (int)(shift >> 3);  // == digit, or digit - 1 ... the aligned window that ends at or above h
    uint64_t t = 0;
    for (int s = 0; s < nranks; ++s) t += per_rank[s * kRow + (size_t)wd * kMultiBuckets + b];
    multi_plan_map(per_rank + (size_t)wd * kMultiBuckets, kRow, t, grand, nranks, rank, peer_ptrs, capacity, dst_ptrs, dst_seg,
                   abort_flag, result, shift);
}

// The map for a window whose histogram was taken separately: rows = [nranks][256] counts of ((key >> shift) & 255).
__global__ void __launch_bounds__(kMultiBuckets)
multi_window_kernel(const uint64_t* __restrict__ rows, int nranks, int rank, const uint64_t* __restrict__ peer_ptrs,
                    uint64_t capacity, uint64_t* __restrict__ dst_ptrs, uint32_t* __restrict__ dst_seg,
                    uint32_t* __restrict__ abort_flag, MultiResult* __restrict__ result, uint32_t shift)
{
    __shared__ uint64_t s_red[kMultiBuckets];
    const int b = threadIdx.x;
    uint64_t tot = 0;
    for (int s = 0; s < nranks; ++s) tot += rows[(size_t)s * kMultiBuckets + b];
    s_red[b] = tot;
    __syncthreads();
    for (int o = kMultiBuckets / 2; o > 0; o >>= 1) {
        if (b < o) s_red[b] += s_red[b + o];
        __syncthreads();
    }
    multi_plan_map(rows, kMultiBuckets, tot, s_red[0], nranks, rank, peer_ptrs, capacity, dst_ptrs, dst_seg, abort_flag, result, shift);
}

}  // namespace lsd

using namespace lsd;

struct lsd_multi_ctx {
    lsd_multi_comm comm;
    int device = 0;
    int r = 8;
    uint32_t* recv = nullptr;
    uint64_t capacity = 0;
    uint64_t max_n_local = 0;
    std::vector<uint64_t> peer_ptrs;   // every rank's receive buffer as seen from this process
    std::vector<uint64_t> peer_offs;   // offset of the buffer inside its IPC allocation (for lsd_ipc_close)
    std::vector<char> peer_ipc;        // 1: mapped with lsd_ipc_open (another process)
    // device workspace owned by the context
    char* ws = nullptr;
    size_t ws_bytes = 0;
    size_t off_hist = 0, off_gather = 0, off_peers = 0, off_dst = 0, off_seg = 0, off_abort = 0, off_result = 0, off_sort = 0;
    size_t sort_ws_bytes = 0;
    MultiResult* host_result = nullptr;  // pinned
    cudaStream_t side = nullptr;
    cudaEvent_t ev_plan = nullptr, ev_copied = nullptr;
    cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};  // stage boundaries when timing is on
    bool timing = false, timed = false;
    lsd_multi_stats last = {};
};

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" {

LSD_API int lsd_multi_ctx_create(const lsd_multi_comm* comm, uint32_t* recv, uint64_t capacity, uint64_t max_n_local, int r,
                                 lsd_multi_ctx** out, lsd_stream_t stream)
{
    if (!comm || !out || !recv || capacity == 0) return LSD_ERR_INVALID_VALUE;
    if (comm->struct_bytes != sizeof(lsd_multi_comm) || !comm->all_gather || !comm->barrier) return LSD_ERR_INVALID_VALUE;
    if (comm->nranks < 1 || comm->nranks > kMultiMaxRanks || comm->rank < 0 || comm->rank >= comm->nranks) return LSD_ERR_INVALID_VALUE;
    if (r != 8) return LSD_ERR_UNSUPPORTED;  // the exchange partitions on the top 8-bit digit
    if (!aligned_to(recv, 16)) return LSD_ERR_ALIGNMENT;
    lsd_multi_ctx* c = new (std::nothrow) lsd_multi_ctx();
    if (!c) return LSD_ERR_INVALID_VALUE;
    c->comm = *comm;
    c->r = r;
    c->recv = recv;
    c->capacity = capacity;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = comm->nranks;
    int rc = LSD_OK;
    auto fail = [&](int code) {
        lsd_multi_ctx_destroy(c);
        return code;
    };
    if (cudaGetDevice(&c->device) != cudaSuccess) return fail(LSD_ERR_CUDA);
    c->max_n_local = max_n_local;
    c->sort_ws_bytes = lsd_sort_workspace_bytes(capacity > max_n_local ? capacity : max_n_local, r, 0);
    if (c->sort_ws_bytes == 0) return fail(LSD_ERR_UNSUPPORTED);
    size_t off = 0;
    c->off_hist = off;    off = align256(off + sizeof(uint64_t) * (32 / r) * kMultiBuckets);
    c->off_gather = off;  off = align256(off + sizeof(uint64_t) * (size_t)N * kMultiDigits * kMultiBuckets);
    c->off_peers = off;   off = align256(off + sizeof(uint64_t) * N);
    c->off_dst = off;     off = align256(off + sizeof(uint64_t) * kMultiBuckets);
    c->off_seg = off;     off = align256(off + sizeof(uint32_t) * kMultiBuckets);
    c->off_abort = off;   off = align256(off + 256);
    c->off_result = off;  off = align256(off + sizeof(MultiResult));
    c->off_sort = off;    off = align256(off + c->sort_ws_bytes);
    c->ws_bytes = off;
    if (cudaMalloc(&c->ws, c->ws_bytes) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (cudaMallocHost(&c->host_result, sizeof(MultiResult)) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming) != cudaSuccess) return fail(LSD_ERR_CUDA);

    // exchange the receive buffers' addresses through the all-gather callback (96 bytes per rank, staged in the gather
    // area, which holds 8 KiB per rank): a CUDA IPC handle for ranks in other processes, the plain pointer (plus peer
    // access) for ranks that are threads of this process
    struct Handle { unsigned char h[64]; uint64_t off; uint64_t raw; int64_t pid; int32_t device; int32_t pad; };
    static_assert(sizeof(Handle) == 96, "96 bytes per rank");
    Handle mine;
    memset(&mine, 0, sizeof(mine));
    rc = lsd_ipc_export(recv, mine.h, &mine.off);
    if (rc != LSD_OK) return fail(rc);
    mine.raw = (uint64_t)(uintptr_t)recv;
    mine.pid = (int64_t)getpid();
    mine.device = c->device;
    char* stage_send = c->ws + c->off_hist;
    char* stage_recv = c->ws + c->off_gather;
    if (cudaMemcpyAsync(stage_send, &mine, sizeof(mine), cudaMemcpyHostToDevice, s) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (comm->all_gather(comm->ctx, stage_send, stage_recv, sizeof(Handle), stream) != 0) return fail(LSD_ERR_COMM);
    std::vector<Handle> all(N);
    if (cudaMemcpyAsync(all.data(), stage_recv, sizeof(Handle) * N, cudaMemcpyDeviceToHost, s) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LSD_ERR_CUDA);
    c->peer_ptrs.assign(N, 0);
    c->peer_offs.assign(N, 0);
    c->peer_ipc.assign(N, 0);
    for (int p = 0; p < N; ++p) {
        if (p == comm->rank) {
            c->peer_ptrs[p] = (uint64_t)(uintptr_t)recv;
        } else if (all[p].pid == mine.pid) {  // a thread of this process: the pointer is valid here, enable peer access
            if (all[p].device != c->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(all[p].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_last_cuda_error(e);
                    return fail(LSD_ERR_CUDA);
                }
                (void)cudaGetLastError();
            }
            c->peer_ptrs[p] = all[p].raw;
        } else {
            void* ptr = nullptr;
            rc = lsd_ipc_open(all[p].h, all[p].off, &ptr);
            if (rc != LSD_OK) return fail(rc);
            c->peer_ptrs[p] = (uint64_t)(uintptr_t)ptr;
            c->peer_offs[p] = all[p].off;
            c->peer_ipc[p] = 1;
        }
    }
    if (cudaMemcpyAsync(c->ws + c->off_peers, c->peer_ptrs.data(), sizeof(uint64_t) * N, cudaMemcpyHostToDevice, s) != cudaSuccess)
        return fail(LSD_ERR_CUDA);
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LSD_ERR_CUDA);
    if (comm->barrier(comm->ctx, stream) != 0) return fail(LSD_ERR_COMM);  // every rank has mapped every buffer
    *out = c;
    return LSD_OK;
}

LSD_API int lsd_multi_ctx_destroy(lsd_multi_ctx* c)
{
    if (!c) return LSD_OK;
    for (size_t p = 0; p < c->peer_ptrs.size(); ++p)
        if (p < c->peer_ipc.size() && c->peer_ipc[p]) lsd_ipc_close((void*)(uintptr_t)c->peer_ptrs[p], c->peer_offs[p]);
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    for (cudaEvent_t e : c->ev_t)
        if (e) cudaEventDestroy(e);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->host_result) cudaFreeHost(c->host_result);
    if (c->ws) cudaFree(c->ws);
    delete c;
    return LSD_OK;
}

LSD_API int lsd_sort_multi(lsd_multi_ctx* c, const uint32_t* keys, uint64_t n_local, uint32_t* scratch, uint64_t* n_out,
                           lsd_stream_t stream)
{
    if (!c || !n_out || (n_local > 0 && !keys) || !scratch) return LSD_ERR_INVALID_VALUE;
    if (n_local > c->max_n_local) return LSD_ERR_INVALID_VALUE;
    if (n_local > 0 && !aligned_to(keys, 16)) return LSD_ERR_ALIGNMENT;
    if (!aligned_to(scratch, 16)) return LSD_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const lsd_multi_comm& comm = c->comm;
    const int N = comm.nranks;
    uint64_t* hist = reinterpret_cast<uint64_t*>(c->ws + c->off_hist);
    uint64_t* gathered = reinterpret_cast<uint64_t*>(c->ws + c->off_gather);
    const uint64_t* peers = reinterpret_cast<const uint64_t*>(c->ws + c->off_peers);
    uint64_t* dst = reinterpret_cast<uint64_t*>(c->ws + c->off_dst);
    uint32_t* seg = reinterpret_cast<uint32_t*>(c->ws + c->off_seg);
    uint32_t* abort_flag = reinterpret_cast<uint32_t*>(c->ws + c->off_abort);
    MultiResult* result = reinterpret_cast<MultiResult*>(c->ws + c->off_result);
    void* sort_ws = c->ws + c->off_sort;

    c->timed = false;
    if (c->timing) LSD_CUDA_TRY(cudaEventRecord(c->ev_t[0], s));
    // 1. digit histograms of the local keys ([4][256]: one read; the exchange digit is chosen from all four)
    int rc = launch_digit_histograms(keys, n_local, c->r, hist, s);
    if (rc != LSD_OK) return rc;
    // 2. the histograms of all ranks (their sum is the all-reduced histogram of every digit)
    if (comm.all_gather(comm.ctx, hist, gathered, sizeof(uint64_t) * kMultiDigits * kMultiBuckets, stream) != 0) return LSD_ERR_COMM;
    // 3. exchange digit, bucket -> rank map, destination pointers and segments, shares: on the device
    multi_plan_kernel<<<1, kMultiBuckets, 0, s>>>(gathered, N, comm.rank, peers, c->capacity, dst, seg, abort_flag, result);
    LSD_LAUNCH_CHECK();
    // the host needs 64 bytes of the plan: where the exchange window sits (the pass kernel is instantiated per digit
    // position, with a run-time form for the others) and this rank's share (to enqueue the local sort)
    LSD_CUDA_TRY(cudaMemcpyAsync(c->host_result, result, sizeof(MultiResult), cudaMemcpyDeviceToHost, s));
    LSD_CUDA_TRY(cudaEventRecord(c->ev_copied, s));
    LSD_CUDA_TRY(cudaEventSynchronize(c->ev_copied));
    if (c->host_result->need_window_hist) {
        // the 8 most significant varying bits straddle two digits (keys in a range like [0, 2^17)): histogram that window
        // (one more read of the local keys), gather the 2 KiB rows, map
        const uint32_t shift = c->host_result->shift;
        rc = launch_field_histogram(keys, n_local, (int)shift, 8, hist, s);
        if (rc != LSD_OK) return rc;
        if (comm.all_gather(comm.ctx, hist, gathered, sizeof(uint64_t) * kMultiBuckets, stream) != 0) return LSD_ERR_COMM;
        multi_window_kernel<<<1, kMultiBuckets, 0, s>>>(gathered, N, comm.rank, peers, c->capacity, dst, seg, abort_flag, result, shift);
        LSD_LAUNCH_CHECK();
        LSD_CUDA_TRY(cudaMemcpyAsync(c->host_result, result, sizeof(MultiResult), cudaMemcpyDeviceToHost, s));
        LSD_CUDA_TRY(cudaEventRecord(c->ev_copied, s));
        LSD_CUDA_TRY(cudaEventSynchronize(c->ev_copied));
    }
    if (c->timing) LSD_CUDA_TRY(cudaEventRecord(c->ev_t[1], s));
    const MultiResult res = *c->host_result;
    c->last.n_in = n_local;
    c->last.n_out = res.n_out;
    c->last.n_out_max = res.n_out_max;
    c->last.sent_bytes = 4 * res.sent;
    c->last.first_bucket = res.first_bucket;
    c->last.last_bucket = res.last_bucket;
    c->last.exchange_shift = res.keep_local ? 0xFFFFFFFFu : res.shift;
    *n_out = res.n_out;
    if (res.overflow) {  // identical on every rank: all return the same status, nothing is moved
        *n_out = res.n_out_max;
        return LSD_ERR_CAPACITY;
    }
    // 4. fused partition + exchange on the chosen digit: every rank is done reading what the previous exchange left in its
    //    buffer, then the pass stores each destination segment into its owner's buffer, then every rank's stores have landed
    if (comm.barrier(comm.ctx, stream) != 0) return LSD_ERR_COMM;
    if (res.keep_local) {  // all keys equal: nothing to exchange, the rank's keys are its slice
        if (n_local > 0) LSD_CUDA_TRY(cudaMemcpyAsync(c->recv, keys, sizeof(uint32_t) * n_local, cudaMemcpyDeviceToDevice, s));
    } else {
        rc = pass_enqueue(keys, nullptr, n_local, c->r, 0, 0, sort_ws, c->sort_ws_bytes, nullptr, s, dst, seg, abort_flag, (int)res.shift);
        if (rc != LSD_OK) return rc;
    }
    if (comm.barrier(comm.ctx, stream) != 0) return LSD_ERR_COMM;
    if (c->timing) LSD_CUDA_TRY(cudaEventRecord(c->ev_t[2], s));
    // 5. local LSD sort of the owned key range
    rc = lsd_sort(c->recv, scratch, res.n_out, c->r, 0, sort_ws, c->sort_ws_bytes, stream);
    if (rc != LSD_OK) return rc;
    if (c->timing) {
        LSD_CUDA_TRY(cudaEventRecord(c->ev_t[3], s));
        c->timed = true;
    }
    return LSD_OK;
}

LSD_API int lsd_multi_last_stats(lsd_multi_ctx* c, lsd_multi_stats* out)
{
    if (!c || !out) return LSD_ERR_INVALID_VALUE;
    c->last.plan_ms = c->last.exchange_ms = c->last.sort_ms = 0.f;
    if (c->timed) {
        LSD_CUDA_TRY(cudaEventSynchronize(c->ev_t[3]));
        LSD_CUDA_TRY(cudaEventElapsedTime(&c->last.plan_ms, c->ev_t[0], c->ev_t[1]));
        LSD_CUDA_TRY(cudaEventElapsedTime(&c->last.exchange_ms, c->ev_t[1], c->ev_t[2]));
        LSD_CUDA_TRY(cudaEventElapsedTime(&c->last.sort_ms, c->ev_t[2], c->ev_t[3]));
    }
    *out = c->last;
    return LSD_OK;
}

LSD_API int lsd_multi_set_timing(lsd_multi_ctx* c, int enabled)
{
    if (!c) return LSD_ERR_INVALID_VALUE;
    if (enabled)
        for (cudaEvent_t& e : c->ev_t)
            if (!e) LSD_CUDA_TRY(cudaEventCreate(&e));
    c->timing = enabled != 0;
    c->timed = false;
    return LSD_OK;
}

}  // extern "C"
