/*
 * lsdsort.h -- C ABI of liblsdsort (B200 / sm_100a LSD radix sort of uint32 keys).
 *
 * This is the drop-in boundary for the one hot path of emanuele-xyz/LSDRadixSort.
 * The reference has no header or FFI layer; its boundary is the set of free functions
 * in LSDRadixSort/LSDRadixSort.cu.  Every entry point below names the reference
 * interface it replaces (file:line, relative to the reference's LSDRadixSort/ dir).
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - return an int status (LSD_OK == 0); nothing aborts or prints
 *     (replaces CUDA_CALL / MYASSERT crash-on-error, CudaUtils.h:7-8, Utils.h:6-15).
 *   - device pointers unless the name says "host"; caller owns all memory;
 *     no hidden allocation: scratch comes from the *_workspace_bytes queries
 *     (replaces GetGPUPrefixSumBlockSumsCount sizing, LSDRadixSort.cu:265-276, and the
 *     3*G*H histogram scratch sized at :919-929).
 *   - asynchronous on the caller's stream (`lsd_stream_t` is a cudaStream_t / CUstream);
 *     the reference used the legacy default stream plus two streams it created per call
 *     (LSDRadixSort.cu:841-842).
 *   - n is 64-bit (the reference uses `int count` everywhere); any n >= 0 works, including
 *     0, 1 and non-multiples of the tile (the reference requires count % block == 0).
 *   - r = radix bits per digit, one of 1, 2, 4, 8 (reference: r in {1,2,4,8}, GPU path
 *     rejects r > 10 at :953).  block = CUDA threads per block with the reference's
 *     meaning; 0 picks the tuned default.
 */
#ifndef LSDSORT_H
#define LSDSORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LSD_API __declspec(dllexport)
#else
#define LSD_API __attribute__((visibility("default")))
#endif

typedef struct CUstream_st *lsd_stream_t; /* == cudaStream_t */

enum lsd_status {
    LSD_OK = 0,
    LSD_ERR_INVALID_VALUE = 1,       /* bad r / block / bit_group / NULL pointer with n > 0 */
    LSD_ERR_WORKSPACE_TOO_SMALL = 2, /* ws_bytes < the matching *_workspace_bytes query */
    LSD_ERR_CUDA = 3,                /* a CUDA call or launch failed; see lsd_last_cuda_error */
    LSD_ERR_UNSUPPORTED = 4,         /* size beyond what this build addresses, or a tuning variant without the requested form */
    LSD_ERR_ALIGNMENT = 5,           /* key pointers must be 16-byte aligned, workspace 256-byte */
    LSD_ERR_CAPACITY = 6,            /* lsd_sort_multi: some rank would own more keys than its receive buffer holds */
    LSD_ERR_COMM = 7                 /* lsd_sort_multi: a communication callback reported failure */
};

#define LSD_VERSION 100 /* 0.1.0 */

LSD_API int lsd_version(void);
LSD_API const char *lsd_status_string(int status);
/* cudaError_t value behind the most recent LSD_ERR_CUDA on the calling thread (0 if none). */
LSD_API int lsd_last_cuda_error(void);
/* Bind the calling thread to a device (one process per GPU: call once with LOCAL_RANK). */
LSD_API int lsd_set_device(int device);
/* Device facts the host side sizes grids with: sm_count, max opt-in shared memory per block. */
LSD_API int lsd_device_info(int *sm_count, int *smem_optin_bytes, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------
 * build_histogram
 * Replaces: BuildHistogramsKernel<<<G,B,H*4>>>(a, h, count, r, bit_group)
 *           (LSDRadixSort.cu:660-702; launched at :770 and :850), CPU twin
 *           BuildHistogramsCPU (:643-658).
 * hist is [G][2^r] tile-major uint32, G = ceil(n / block); every cell is overwritten
 * (like the kernel; no pre-zeroing needed).  block is the keys-per-histogram of the
 * reference (its threads per block), any value in [1, 2^20].
 * ------------------------------------------------------------------------------------- */
LSD_API size_t lsd_build_histogram_bytes(uint64_t n, int r, int block);
LSD_API int lsd_build_histogram(const uint32_t *keys, uint64_t n, int r, int bit_group, int block,
                                uint32_t *hist, lsd_stream_t stream);

/* Whole-array histograms of ALL digits in one read of the keys:
 * hist is [32/r][2^r] uint64, overwritten.  No reference counterpart as a function: it is
 * the column sum of build_histogram over all tiles for every bit group (what the reference
 * recomputes per pass at :850), hoisted out of the pass loop. */
LSD_API int lsd_digit_histograms(const uint32_t *keys, uint64_t n, int r, uint64_t *hist, lsd_stream_t stream);
/* Same layout ([32/r][2^r] uint64, overwritten) but only the TOP digit's row is counted, the others are zero: one
 * shared-memory atomic per key instead of 32/r.  The planning step of the multi-GPU sort (bucket -> rank map). */
LSD_API int lsd_top_digit_histogram(const uint32_t *keys, uint64_t n, int r, uint64_t *hist, lsd_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * prefix_sum
 * Replaces: void GPUPrefixSum(uint32_t* d_a, int count, int threads_per_block,
 *                             uint32_t* d_block_sums, cudaStream_t s)  (LSDRadixSort.cu:286-302)
 *           + int GetGPUPrefixSumBlockSumsCount(int count, int tpb)    (:265-276)
 *           CPU twin PrefixSum (:128-139).
 * In place, EXCLUSIVE, uint32 wrap-around (mod 2^32).  Single pass, decoupled look-back.
 * ------------------------------------------------------------------------------------- */
LSD_API size_t lsd_prefix_sum_workspace_bytes(uint64_t n, int block);
LSD_API int lsd_prefix_sum(uint32_t *a, uint64_t n, int block, void *ws, size_t ws_bytes, lsd_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * LSD sort
 * Replaces: void GPULSDRadixSort(uint32_t* a, uint32_t* b, uint32_t* h, uint32_t* block_sums,
 *                                uint32_t* d, int grid, int block, int block_sums_count,
 *                                int count, int h_count, int r)       (LSDRadixSort.cu:839-910)
 *           CPU twin LSDRadixSort (:62-69).
 * keys  : n keys in, ascending keys out (the reference leaves its result in `a`, :1005).
 * scratch: n keys of ping-pong space (the reference's `b`); contents undefined afterwards.
 * 32/r passes, least significant digit first, each pass stable.  Passes whose digit is
 * constant over the whole input are skipped (decided on the device; the call stays async).
 * ------------------------------------------------------------------------------------- */
typedef struct lsd_sort_options {
    uint32_t struct_bytes;  /* sizeof(lsd_sort_options); 0-initialise the rest */
    uint32_t portion_keys;  /* 0 = default; max keys per look-back portion (tests use small values) */
    uint32_t disable_skip;  /* 1 = run every pass even if its digit is constant */
    uint32_t variant;       /* 0 = default kernel shape; other values select tuning variants */
    uint64_t debug_trace;   /* 0, or a device pointer to 16 uint64 per tile of the LAST pass: per-phase SM clocks
                               (tuning aid, only honoured by kernels that support it) */
    uint32_t key_type;      /* enum lsd_key_type: how the 32-bit words are ORDERED (default LSD_KEY_U32, the reference) */
    uint32_t reserved;      /* 0 */
} lsd_sort_options;

/* Key orderings beyond the reference's (its keys are uint32 only, LSDRadixSort.cu:839).  The 32-bit
 * patterns are mapped to unsigned order by a bijection applied when the first executed pass reads the keys and undone
 * when the last executed pass writes them (and inside the digit histogram), so it costs no extra pass over the data:
 *   LSD_KEY_I32: two's-complement order (flip the sign bit);
 *   LSD_KEY_F32: IEEE-754 total order: -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN (flip all bits of negatives,
 *                the sign bit of the others) -- equal to `<` on floats wherever `<` decides.
 * Supported by the default kernel shapes (variant 0, any r, any block) of lsd_sort_ex / lsd_sort_pairs / *_timed;
 * other variants return LSD_ERR_UNSUPPORTED. */
enum lsd_key_type {
    LSD_KEY_U32 = 0, LSD_KEY_I32 = 1, LSD_KEY_F32 = 2,
    LSD_KEY_U64 = 3, LSD_KEY_I64 = 4, LSD_KEY_F64 = 5 /* lsd_sort64 only */
};

LSD_API size_t lsd_sort_workspace_bytes(uint64_t n, int r, int block);
LSD_API size_t lsd_sort_workspace_bytes_ex(uint64_t n, int r, int block, const lsd_sort_options *opt);
LSD_API int lsd_sort(uint32_t *keys, uint32_t *scratch, uint64_t n, int r, int block, void *ws, size_t ws_bytes,
                     lsd_stream_t stream);
LSD_API int lsd_sort_ex(uint32_t *keys, uint32_t *scratch, uint64_t n, int r, int block, void *ws, size_t ws_bytes,
                        const lsd_sort_options *opt, lsd_stream_t stream);

/* Key-value form of lsd_sort_ex: every key carries a 32-bit value (a row id, an index, any payload) that follows it
 * through every pass.  keys/vals: n entries in, out; keys_scratch/vals_scratch: n entries of ping-pong space each.
 * Result: keys ascending, and vals[i] is the value that came with keys[i]; equal keys keep their INPUT order (the
 * reference states its passes are stable, LSDRadixSort.cu:25-54 -- with keys alone that is unobservable, SURVEY 8(f)2;
 * with vals = 0..n-1 the output vals are the stable sorting permutation).  The reference's scatter moves keys only
 * (LSDRadixSort.cu:836); this is the payload extension of the same pass.  Same passes, plan and skipping as lsd_sort;
 * workspace from lsd_sort_pairs_workspace_bytes (opt may be NULL).  16 B per pair per pass of HBM traffic. */
LSD_API size_t lsd_sort_pairs_workspace_bytes(uint64_t n, int r, int block, const lsd_sort_options *opt);
LSD_API int lsd_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_scratch, uint32_t *vals_scratch, uint64_t n,
                           int r, int block, void *ws, size_t ws_bytes, const lsd_sort_options *opt,
                           lsd_stream_t stream);
/* lsd_sort_pairs with per-stage device times, as lsd_sort_timed (stage_ms[1+passes] = both copy-backs). */
LSD_API int lsd_sort_pairs_timed(uint32_t *keys, uint32_t *vals, uint32_t *keys_scratch, uint32_t *vals_scratch,
                                 uint64_t n, int r, int block, void *ws, size_t ws_bytes, const lsd_sort_options *opt,
                                 lsd_stream_t stream, float *stage_ms, int stage_cap, int *stages_written);

/* 64-bit keys (the reference sorts uint32 only, LSDRadixSort.cu:62, :839).  keys: n 64-bit keys in,
 * ascending out in the order key_type names (LSD_KEY_U64: unsigned; LSD_KEY_I64: two's complement; LSD_KEY_F64: IEEE
 * total order, as LSD_KEY_F32).  scratch: n 64-bit entries of ping-pong space.  Eight stable 8-bit passes: the keys are
 * split into low / high word arrays, every pass is the key-value pass of lsd_sort_pairs with the digit's word as the key
 * and the other word as its value (16 B per key per pass, what a native 64-bit pass moves), and a final kernel merges
 * the words back.  Digits that are constant over the input are skipped per word (keys below 2^32: four passes).
 * n <= 2^32.  Workspace from lsd_sort64_workspace_bytes. */
LSD_API size_t lsd_sort64_workspace_bytes(uint64_t n);
LSD_API int lsd_sort64(uint64_t *keys, uint64_t *scratch, uint64_t n, uint32_t key_type, void *ws, size_t ws_bytes,
                       lsd_stream_t stream);

/* One stable counting-sort pass on digit `bit_group`: out <- in reordered by that digit, keys with
 * equal digits keeping their input order.  Replaces one iteration of the reference's pass loop
 * (LSDRadixSort.cu:845-906; CPU twin LSDRadixSortPass, :25-54, without its copy-back at :53).
 * in and out must not overlap.  If hist_out is non-NULL it receives the 2^r bucket START offsets
 * (uint64, what the CPU twin leaves in `histogram`).  Workspace: lsd_sort_workspace_bytes(n, r, block).
 * Also the MSD partition step of the multi-GPU sort (bit_group = 32/r - 1). */
LSD_API int lsd_sort_pass(const uint32_t *in, uint32_t *out, uint64_t n, int r, int bit_group, int block, void *ws,
                          size_t ws_bytes, uint64_t *hist_out, lsd_stream_t stream);

/* Peer-scatter form of lsd_sort_pass, the fused partition + exchange step of the multi-GPU sort.  Buckets of digit
 * `bit_group` are grouped into destination SEGMENTS of consecutive buckets: dst_seg[d] = first | last << 16 names the
 * segment of bucket d (NULL: every bucket is its own segment) and dst_ptrs[d] its destination (equal for all buckets
 * of a segment).  The keys of a segment are appended to ((uint32_t*)dst_ptrs[d]) tile after tile, inside a tile in
 * (bucket, input position) order: deterministic, stable inside a bucket, and with one bucket per segment exactly the
 * output of lsd_sort_pass.  dst_ptrs / dst_seg are DEVICE arrays of 2^r entries; a pointer may aim into this GPU's
 * memory or into a peer GPU's buffer opened with lsd_ipc_open, in which case the keys cross NVLink as the pass
 * kernel's own stores (no separate all-to-all, long contiguous runs per destination).  No reference counterpart (the
 * reference is single-GPU).  The caller orders the peers (a barrier before the buffers are reused and
 * after the pass, before they are read).  Workspace: lsd_sort_workspace_bytes(n, r, 0). */
LSD_API int lsd_sort_pass_scatter(const uint32_t *in, uint64_t n, int r, int bit_group, const uint64_t *dst_ptrs,
                                  const uint32_t *dst_seg, void *ws, size_t ws_bytes, lsd_stream_t stream);

/* CUDA IPC plumbing for the above (one process per GPU on one node).  lsd_ipc_export: 64 opaque handle bytes for the
 * allocation that contains dev_ptr plus dev_ptr's offset inside it; send both to the peer process.  lsd_ipc_open:
 * maps the peer's allocation (peer access enabled lazily) and returns the pointer that corresponds to dev_ptr. */
LSD_API int lsd_ipc_export(const void *dev_ptr, void *handle64, uint64_t *offset_out);
LSD_API int lsd_ipc_open(const void *handle64, uint64_t offset, void **peer_ptr_out);
LSD_API int lsd_ipc_close(void *peer_ptr, uint64_t offset);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU sort (one process or thread per GPU, one node).  No reference counterpart: the reference is single-GPU
 * (SURVEY 2.4); this is BASELINE.json's partitioning: every rank histograms the 8-bit digits of its keys
 * (one read), the [4][256] histograms are all-gathered (their sum is the all-reduced histogram of every digit), every
 * rank derives the same contiguous bucket -> rank map ON THE DEVICE for the 8-bit window that ends at the highest BIT that
 * varies over the whole input -- the top digit unless the keys live in a narrow range (all bits above the window are
 * constant, so bucket ranges of the window are still contiguous key ranges: keys below 2^24, in [0, 2^17) or in
 * [2^31, 2^31 + 1000) are balanced over their own 8 most significant varying bits instead of landing on one rank; a
 * window that straddles two digits costs one more read of the keys for its histogram); one pass
 * kernel partitions the local keys by owner on that window and stores them straight into the owners' receive buffers over
 * NVLink peer memory (CUDA IPC; lsd_sort_pass_scatter), and every rank sorts what arrived.  Rank k ends with the k-th
 * slice of the global order in its receive buffer.  If every key of every rank is the same, nothing is exchanged.
 *
 * The two collectives are callbacks, so liblsdsort links no communication library: include/lsdsort_nccl.h fills an
 * lsd_multi_comm from an ncclComm_t, lsdradixsort_b200/multi.py from torch.distributed.  Both must be ordered on the
 * stream they are given (as NCCL calls are) and return 0 on success.
 * ------------------------------------------------------------------------------------- */
typedef struct lsd_multi_comm {
    uint32_t struct_bytes; /* sizeof(lsd_multi_comm) */
    int rank, nranks;      /* nranks <= 64 */
    /* gather `bytes` bytes from every rank's DEVICE buffer `send` into the DEVICE buffer `recv` (rank-major) */
    int (*all_gather)(void *ctx, const void *send, void *recv, size_t bytes, lsd_stream_t stream);
    /* work enqueued on `stream` after the barrier starts on no rank before every rank's work before it has finished */
    int (*barrier)(void *ctx, lsd_stream_t stream);
    void *ctx;
} lsd_multi_comm;

typedef struct lsd_multi_stats {
    uint64_t n_in, n_out;   /* keys this rank brought / owns after the exchange */
    uint64_t n_out_max;     /* the largest share of any rank */
    uint64_t sent_bytes;    /* bytes that left this GPU over NVLink (excludes what it kept) */
    uint32_t first_bucket;  /* buckets of the exchange window this rank owns: [first, last]; first > last when it owns none */
    uint32_t last_bucket;
    float plan_ms;          /* with lsd_multi_set_timing(ctx, 1): device time of histogram + all-gather + plan, */
    float exchange_ms;      /* of barrier + partition/exchange pass + barrier, */
    float sort_ms;          /* and of the local sort; 0 otherwise */
    uint32_t exchange_shift; /* the exchange partitioned on bits [shift, shift + 8) (24 = the top digit); 0xFFFFFFFF: all keys
                                equal, nothing moved */
} lsd_multi_stats;

typedef struct lsd_multi_ctx lsd_multi_ctx;
/* COLLECTIVE: every rank calls it with its own receive buffer (`capacity` keys, from cudaMalloc: it is exported with
 * CUDA IPC and mapped by every other rank).  max_n_local bounds the n_local of later lsd_sort_multi calls on this rank.
 * The context owns a small plan area and one workspace for max(capacity, max_n_local) keys (lsd_sort_workspace_bytes;
 * shared by the exchange pass and the local sort), a side stream and 64 pinned bytes.  r must be 8. */
LSD_API int lsd_multi_ctx_create(const lsd_multi_comm *comm, uint32_t *recv, uint64_t capacity, uint64_t max_n_local, int r,
                                 lsd_multi_ctx **out, lsd_stream_t stream);
LSD_API int lsd_multi_ctx_destroy(lsd_multi_ctx *ctx);
/* COLLECTIVE: sorts the union of every rank's `keys` (n_local <= max_n_local keys each, not modified).  On return the rank's slice
 * (*n_out keys, ascending; every key of rank k <= every key of rank k+1) is being written to the context's receive
 * buffer on `stream`; `scratch` (capacity keys) is ping-pong space for the local sort.  No host synchronisation sits
 * in front of the exchange other than the wait for a 64-byte copy of the plan's result (the digit to partition on and
 * *n_out, which the host needs to enqueue the exchange pass and the local sort).  If any rank's share exceeds its
 * capacity (the balance is only as fine as one of the 256 buckets of the exchange digit: a few very frequent key
 * prefixes), EVERY rank returns LSD_ERR_CAPACITY, *n_out = the largest share (so the caller can retry with bigger
 * buffers) and no key is moved. */
LSD_API int lsd_sort_multi(lsd_multi_ctx *ctx, const uint32_t *keys, uint64_t n_local, uint32_t *scratch, uint64_t *n_out,
                           lsd_stream_t stream);
/* Stats of the most recent lsd_sort_multi on this context.  With timing enabled the call synchronises the stream of
 * that sort first (CUDA events bracket its three stages). */
LSD_API int lsd_multi_last_stats(lsd_multi_ctx *ctx, lsd_multi_stats *out);
LSD_API int lsd_multi_set_timing(lsd_multi_ctx *ctx, int enabled);

/* Same as lsd_sort_ex, but brackets every kernel with CUDA events on `stream`, synchronises,
 * and reports per-stage device times.  stage_ms[0] = digit histogram + plan, stage_ms[1+p] =
 * pass p (0 if skipped), stage_ms[1+passes] = copy-back (0 if none).  Measurement aid for
 * bench.py's roofline leg; replaces the cudaEvent pair at LSDRadixSort.cu:999-1008. */
LSD_API int lsd_sort_timed(uint32_t *keys, uint32_t *scratch, uint64_t n, int r, int block, void *ws,
                           size_t ws_bytes, const lsd_sort_options *opt, lsd_stream_t stream, float *stage_ms,
                           int stage_cap, int *stages_written);

/* After a sort on `stream` has been enqueued: synchronises the stream and reports which
 * passes the device-side plan skipped (bit p set = pass p skipped) and how many kernels
 * of this library the sort launched. */
LSD_API int lsd_sort_read_plan(const void *ws, uint64_t n, int r, uint32_t *skipped_mask, int *launches,
                               lsd_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Host-buffer entry (what TestGPULSDRadixSort does around the call, LSDRadixSort.cu:1001-1005:
 * H2D copy, sort, D2H copy).  The context owns the device buffers so repeated calls do not
 * allocate.  host_keys may be pageable or pinned; the call returns after the sorted keys are
 * back in host_keys.
 * ------------------------------------------------------------------------------------- */
typedef struct lsd_host_ctx lsd_host_ctx;
LSD_API int lsd_host_ctx_create(uint64_t max_n, int r, int block, lsd_host_ctx **out);
LSD_API int lsd_host_ctx_destroy(lsd_host_ctx *ctx);
LSD_API int lsd_sort_host(lsd_host_ctx *ctx, uint32_t *host_keys, uint64_t n);
/* The same three steps enqueued on the context's own stream WITHOUT waiting: the call returns at once (host_keys must be
 * pinned -- lsd_host_alloc -- for the copies to be asynchronous; a pageable buffer makes the call block instead) and
 * lsd_host_ctx_wait returns when the sorted keys are back in host_keys.  One sort per context at a time.  Two contexts
 * used alternately overlap the D2H copy of one array with the H2D copy of the next (PCIe is full duplex), which is where
 * the time of a host-buffer sort goes: 2^28 keys cost 2 x 20 ms of copies around 2.8 ms of sorting
 * (the reference copies and sorts on one stream, synchronously: LSDRadixSort.cu:1001-1005). */
LSD_API int lsd_sort_host_async(lsd_host_ctx *ctx, uint32_t *host_keys, uint64_t n);
LSD_API int lsd_host_ctx_wait(lsd_host_ctx *ctx);
/* Pinned host memory helpers (replace MyCudaHostAlloc, CudaUtils.cpp:3-8). */
LSD_API int lsd_host_alloc(void **ptr, size_t bytes);
LSD_API int lsd_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* LSDSORT_H */
