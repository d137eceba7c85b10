// api.cu -- the extern "C" surface declared in include/lsdsort.h.
#include <algorithm>
#include <new>

#include "onesweep.cuh"
#include <cuda.h>
#include <cstring>

#include "sort.h"

namespace lsd {

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }

struct DeviceFacts {
    int device = -1;
    int sms = 0;
    int smem_optin = 0;
    int cc_major = 0, cc_minor = 0;
};
static thread_local DeviceFacts g_facts;

static const DeviceFacts& facts()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (g_facts.device != dev) {
        DeviceFacts f;
        f.device = dev;
        cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&f.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&f.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&f.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
        if (f.sms <= 0) f.sms = 148;
        g_facts = f;
    }
    return g_facts;
}
int sm_count() { return facts().sms; }
int smem_optin_bytes() { return facts().smem_optin; }

}  // namespace lsd

using namespace lsd;

extern "C" {

LSD_API int lsd_version(void) { return LSD_VERSION; }

LSD_API const char* lsd_status_string(int status)
{
    switch (status) {
        case LSD_OK: return "ok";
        case LSD_ERR_INVALID_VALUE: return "invalid value";
        case LSD_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case LSD_ERR_CUDA: return "CUDA error";
        case LSD_ERR_UNSUPPORTED: return "unsupported size";
        case LSD_ERR_ALIGNMENT: return "misaligned pointer";
        case LSD_ERR_CAPACITY: return "receive buffer too small for this rank's share";
        case LSD_ERR_COMM: return "communication callback failed";
    }
    return "unknown status";
}

LSD_API int lsd_last_cuda_error(void) { return (int)g_last_cuda_error; }

LSD_API int lsd_set_device(int device)
{
    LSD_CUDA_TRY(cudaSetDevice(device));
    return LSD_OK;
}

LSD_API int lsd_device_info(int* sm, int* smem, int* major, int* minor)
{
    int dev = 0;
    LSD_CUDA_TRY(cudaGetDevice(&dev));
    const DeviceFacts& f = facts();
    if (sm) *sm = f.sms;
    if (smem) *smem = f.smem_optin;
    if (major) *major = f.cc_major;
    if (minor) *minor = f.cc_minor;
    return LSD_OK;
}

// ---- build_histogram -----------------------------------------------------------------
LSD_API size_t lsd_build_histogram_bytes(uint64_t n, int r, int block)
{
    if (!valid_radix(r) || block <= 0) return 0;
    const uint64_t tiles = (n + (uint64_t)block - 1) / (uint64_t)block;
    return (size_t)tiles * ((size_t)1 << r) * sizeof(uint32_t);
}

LSD_API int lsd_build_histogram(const uint32_t* keys, uint64_t n, int r, int bit_group, int block, uint32_t* hist,
                                lsd_stream_t stream)
{
    if (!valid_radix(r) || block <= 0 || block > (1 << 20)) return LSD_ERR_INVALID_VALUE;
    if (bit_group < 0 || bit_group >= 32 / r) return LSD_ERR_INVALID_VALUE;  // the reference shifts by >= 32 here (UB)
    if (n == 0) return LSD_OK;
    if (!keys || !hist) return LSD_ERR_INVALID_VALUE;
    return launch_tile_histograms(keys, n, r, bit_group, block, hist, (cudaStream_t)stream);
}

LSD_API int lsd_digit_histograms(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, lsd_stream_t stream)
{
    if (!accepted_radix(r)) return LSD_ERR_INVALID_VALUE;
    if (!hist || (n > 0 && !keys)) return LSD_ERR_INVALID_VALUE;
    if (n > 0 && !aligned_to(keys, 16)) return LSD_ERR_ALIGNMENT;
    if (composite_radix(r)) return launch_digit_histograms_wide(keys, n, r, hist, (cudaStream_t)stream);
    return launch_digit_histograms(keys, n, r, hist, (cudaStream_t)stream);
}

LSD_API int lsd_top_digit_histogram(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, lsd_stream_t stream)
{
    if (!valid_radix(r)) return LSD_ERR_INVALID_VALUE;
    if (!hist || (n > 0 && !keys)) return LSD_ERR_INVALID_VALUE;
    if (n > 0 && !aligned_to(keys, 16)) return LSD_ERR_ALIGNMENT;
    return launch_top_digit_histogram(keys, n, r, hist, (cudaStream_t)stream);
}

// ---- prefix_sum ----------------------------------------------------------------------
LSD_API size_t lsd_prefix_sum_workspace_bytes(uint64_t n, int block) { return scan_workspace_bytes(n, block); }

LSD_API int lsd_prefix_sum(uint32_t* a, uint64_t n, int block, void* ws, size_t ws_bytes, lsd_stream_t stream)
{
    if (block < 0 || block > 1024) return LSD_ERR_INVALID_VALUE;
    if (n == 0) return LSD_OK;
    if (!a || !ws) return LSD_ERR_INVALID_VALUE;
    if (!aligned_to(a, 16) || !aligned_to(ws, 256)) return LSD_ERR_ALIGNMENT;
    return launch_prefix_sum(a, n, block, ws, ws_bytes, (cudaStream_t)stream);
}

// ---- sort ----------------------------------------------------------------------------
LSD_API size_t lsd_sort_workspace_bytes_ex(uint64_t n, int r, int block, const lsd_sort_options* opt)
{
    if (opt && opt->struct_bytes != sizeof(lsd_sort_options)) return 0;  // same check as lsd_sort_ex
    SortLayout L;
    if (make_layout(n, r, block, opt, &L) != LSD_OK) return 0;
    // composite digit widths: lsd_sort_pass runs sub-passes through a temporary key array in the workspace
    if (composite_radix(r)) return std::max(L.total_bytes, wide_pass_workspace_bytes(n));
    return L.total_bytes;
}
LSD_API size_t lsd_sort_workspace_bytes(uint64_t n, int r, int block)
{
    return lsd_sort_workspace_bytes_ex(n, r, block, nullptr);
}

LSD_API int lsd_sort_ex(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                        const lsd_sort_options* opt, lsd_stream_t stream)
{
    if (opt && opt->struct_bytes != sizeof(lsd_sort_options)) return LSD_ERR_INVALID_VALUE;
    return sort_enqueue(keys, scratch, n, r, block, ws, ws_bytes, opt, (cudaStream_t)stream, nullptr, nullptr);
}
LSD_API int lsd_sort(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                     lsd_stream_t stream)
{
    return lsd_sort_ex(keys, scratch, n, r, block, ws, ws_bytes, nullptr, stream);
}

LSD_API size_t lsd_sort_pairs_workspace_bytes(uint64_t n, int r, int block, const lsd_sort_options* opt)
{
    if (opt && opt->struct_bytes != sizeof(lsd_sort_options)) return 0;
    SortLayout L;
    if (make_layout(n, r, block, opt, &L, true) != LSD_OK) return 0;
    return L.total_bytes;
}

LSD_API int lsd_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_scratch, uint32_t* vals_scratch, uint64_t n,
                           int r, int block, void* ws, size_t ws_bytes, const lsd_sort_options* opt, lsd_stream_t stream)
{
    if (opt && opt->struct_bytes != sizeof(lsd_sort_options)) return LSD_ERR_INVALID_VALUE;
    if (n > 0 && (!vals || !vals_scratch)) return LSD_ERR_INVALID_VALUE;
    if (n == 0) return accepted_radix(r) ? LSD_OK : LSD_ERR_INVALID_VALUE;
    return sort_enqueue(keys, keys_scratch, n, r, block, ws, ws_bytes, opt, (cudaStream_t)stream, nullptr, nullptr, vals,
                        vals_scratch);
}

LSD_API size_t lsd_sort64_workspace_bytes(uint64_t n) { return sort64_workspace_bytes(n); }

LSD_API int lsd_sort64(uint64_t* keys, uint64_t* scratch, uint64_t n, uint32_t key_type, void* ws, size_t ws_bytes,
                       lsd_stream_t stream)
{
    return sort64_enqueue(keys, scratch, n, key_type, ws, ws_bytes, (cudaStream_t)stream);
}

LSD_API int lsd_sort_pass(const uint32_t* in, uint32_t* out, uint64_t n, int r, int bit_group, int block, void* ws,
                          size_t ws_bytes, uint64_t* hist_out, lsd_stream_t stream)
{
    if (composite_radix(r)) return pass_enqueue_wide(in, out, n, r, bit_group, ws, ws_bytes, hist_out, (cudaStream_t)stream);
    if (!valid_radix(r)) return LSD_ERR_INVALID_VALUE;
    return pass_enqueue(in, out, n, r, bit_group, block, ws, ws_bytes, hist_out, (cudaStream_t)stream);
}

LSD_API int lsd_sort_pass_scatter(const uint32_t* in, uint64_t n, int r, int bit_group, const uint64_t* dst_ptrs,
                                  const uint32_t* dst_seg, void* ws, size_t ws_bytes, lsd_stream_t stream)
{
    if (!valid_radix(r) || !dst_ptrs) return LSD_ERR_INVALID_VALUE;
    return pass_enqueue(in, nullptr, n, r, bit_group, 0, ws, ws_bytes, nullptr, (cudaStream_t)stream, dst_ptrs, dst_seg);
}

// ---- peer memory (one process per GPU, same node): CUDA IPC handles for the exchange buffers --------------
LSD_API int lsd_ipc_export(const void* dev_ptr, void* handle64, uint64_t* offset_out)
{
    if (!dev_ptr || !handle64 || !offset_out) return LSD_ERR_INVALID_VALUE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle is passed as 64 opaque bytes");
    cudaIpcMemHandle_t h;
    LSD_CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    // the handle names the whole allocation: report where dev_ptr sits inside it
    // (driver entry point fetched through the runtime: no link-time dependency on libcuda)
    typedef CUresult (*get_range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LSD_CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return LSD_ERR_CUDA;
    CUdeviceptr base = 0;
    size_t size = 0;
    if (reinterpret_cast<get_range_fn>(fn)(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) return LSD_ERR_CUDA;
    *offset_out = (uint64_t)((CUdeviceptr)dev_ptr - base);
    memcpy(handle64, &h, 64);
    return LSD_OK;
}

LSD_API int lsd_ipc_open(const void* handle64, uint64_t offset, void** peer_ptr_out)
{
    if (!handle64 || !peer_ptr_out) return LSD_ERR_INVALID_VALUE;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    LSD_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr_out = static_cast<char*>(base) + offset;
    return LSD_OK;
}

LSD_API int lsd_ipc_close(void* peer_ptr, uint64_t offset)
{
    if (!peer_ptr) return LSD_OK;
    LSD_CUDA_TRY(cudaIpcCloseMemHandle(static_cast<char*>(peer_ptr) - offset));
    return LSD_OK;
}

static int sort_timed_impl(uint32_t* keys, uint32_t* vals, uint32_t* scratch, uint32_t* vals_scratch, uint64_t n, int r,
                           int block, void* ws, size_t ws_bytes, const lsd_sort_options* opt, lsd_stream_t stream,
                           float* stage_ms, int stage_cap, int* stages_written)
{
    if (opt && opt->struct_bytes != sizeof(lsd_sort_options)) return LSD_ERR_INVALID_VALUE;
    if (!accepted_radix(r) || !stage_ms) return LSD_ERR_INVALID_VALUE;
    const int passes = 32 / exec_radix(r);  // composite digit widths run (and are reported as) the 8-bit schedule
    const int stages = passes + 2;
    if (stage_cap < stages) return LSD_ERR_INVALID_VALUE;
    cudaEvent_t ev[kMaxPasses + 3];
    for (int i = 0; i < stages + 1; ++i) {
        const cudaError_t e = cudaEventCreate(&ev[i]);
        if (e != cudaSuccess) {  // do not leak the events created so far
            for (int j = 0; j < i; ++j) cudaEventDestroy(ev[j]);
            set_last_cuda_error(e);
            return LSD_ERR_CUDA;
        }
    }
    int rc = LSD_OK;
    if (n == 0) {
        for (int i = 0; i < stages; ++i) stage_ms[i] = 0.f;
    } else {
        rc = sort_enqueue(keys, scratch, n, r, block, ws, ws_bytes, opt, (cudaStream_t)stream, ev, nullptr, vals,
                          vals_scratch);
        if (rc == LSD_OK) {
            cudaError_t e = cudaEventSynchronize(ev[stages]);
            if (e != cudaSuccess) { set_last_cuda_error(e); rc = LSD_ERR_CUDA; }
        }
        if (rc == LSD_OK)
            for (int i = 0; i < stages; ++i) {
                cudaError_t e = cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
                if (e != cudaSuccess) { set_last_cuda_error(e); rc = LSD_ERR_CUDA; break; }
            }
    }
    for (int i = 0; i < stages + 1; ++i) cudaEventDestroy(ev[i]);
    if (stages_written) *stages_written = stages;
    return rc;
}

LSD_API int lsd_sort_timed(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                           const lsd_sort_options* opt, lsd_stream_t stream, float* stage_ms, int stage_cap,
                           int* stages_written)
{
    return sort_timed_impl(keys, nullptr, scratch, nullptr, n, r, block, ws, ws_bytes, opt, stream, stage_ms, stage_cap,
                           stages_written);
}

LSD_API int lsd_sort_pairs_timed(uint32_t* keys, uint32_t* vals, uint32_t* keys_scratch, uint32_t* vals_scratch,
                                 uint64_t n, int r, int block, void* ws, size_t ws_bytes, const lsd_sort_options* opt,
                                 lsd_stream_t stream, float* stage_ms, int stage_cap, int* stages_written)
{
    if (n > 0 && (!vals || !vals_scratch)) return LSD_ERR_INVALID_VALUE;
    return sort_timed_impl(keys, vals, keys_scratch, vals_scratch, n, r, block, ws, ws_bytes, opt, stream, stage_ms,
                           stage_cap, stages_written);
}

LSD_API int lsd_sort_read_plan(const void* ws, uint64_t n, int r, uint32_t* skipped_mask, int* launches,
                               lsd_stream_t stream)
{
    if (!accepted_radix(r)) return LSD_ERR_INVALID_VALUE;
    r = exec_radix(r);  // composite digit widths: the mask names the executed 8-bit passes
    if (skipped_mask) *skipped_mask = 0;
    if (launches) *launches = 0;
    if (n == 0) return LSD_OK;
    if (!ws) return LSD_ERR_INVALID_VALUE;
    SortPlan plan;
    LSD_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    LSD_CUDA_TRY(cudaMemcpy(&plan, ws, sizeof(plan), cudaMemcpyDeviceToHost));  // the plan sits at offset 0
    uint32_t mask = 0;
    for (int p = 0; p < 32 / r; ++p)
        if (plan.skip[p]) mask |= 1u << p;
    if (skipped_mask) *skipped_mask = mask;
    if (launches) {
        SortLayout L;
        if (make_layout(n, r, 0, nullptr, &L) == LSD_OK) *launches = 3 + L.passes * (int)L.portions;
    }
    return LSD_OK;
}

// ---- host-buffer entry ---------------------------------------------------------------
struct lsd_host_ctx {
    uint64_t max_n;
    int r, block;
    uint32_t* d_keys;
    uint32_t* d_scratch;
    void* d_ws;
    size_t ws_bytes;
    cudaStream_t stream;
};

LSD_API int lsd_host_ctx_create(uint64_t max_n, int r, int block, lsd_host_ctx** out)
{
    if (!out || !accepted_radix(r)) return LSD_ERR_INVALID_VALUE;
    *out = nullptr;
    SortLayout L;
    const int st = make_layout(max_n, r, block, nullptr, &L);
    if (st != LSD_OK) return st;
    lsd_host_ctx* c = new (std::nothrow) lsd_host_ctx();
    if (!c) return LSD_ERR_INVALID_VALUE;
    c->max_n = max_n; c->r = r; c->block = block;
    c->d_keys = nullptr; c->d_scratch = nullptr; c->d_ws = nullptr; c->stream = nullptr;
    c->ws_bytes = L.total_bytes;
    const size_t kb = (size_t)(max_n ? max_n : 1) * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&c->d_keys, kb);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_scratch, kb);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ws, c->ws_bytes ? c->ws_bytes : 256);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_last_cuda_error(e);
        lsd_host_ctx_destroy(c);
        return LSD_ERR_CUDA;
    }
    *out = c;
    return LSD_OK;
}

LSD_API int lsd_host_ctx_destroy(lsd_host_ctx* c)
{
    if (!c) return LSD_OK;
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaFree(c->d_ws);
    cudaFree(c->d_scratch);
    cudaFree(c->d_keys);
    delete c;
    return LSD_OK;
}

LSD_API int lsd_sort_host_async(lsd_host_ctx* c, uint32_t* host_keys, uint64_t n)
{
    if (!c) return LSD_ERR_INVALID_VALUE;
    if (n > c->max_n) return LSD_ERR_INVALID_VALUE;
    if (n == 0) return LSD_OK;
    if (!host_keys) return LSD_ERR_INVALID_VALUE;
    const size_t bytes = (size_t)n * sizeof(uint32_t);
    LSD_CUDA_TRY(cudaMemcpyAsync(c->d_keys, host_keys, bytes, cudaMemcpyHostToDevice, c->stream));
    const int rc = sort_enqueue(c->d_keys, c->d_scratch, n, c->r, c->block, c->d_ws, c->ws_bytes, nullptr, c->stream,
                                nullptr, nullptr);
    if (rc != LSD_OK) return rc;
    LSD_CUDA_TRY(cudaMemcpyAsync(host_keys, c->d_keys, bytes, cudaMemcpyDeviceToHost, c->stream));
    return LSD_OK;
}

LSD_API int lsd_host_ctx_wait(lsd_host_ctx* c)
{
    if (!c) return LSD_ERR_INVALID_VALUE;
    LSD_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LSD_OK;
}

LSD_API int lsd_sort_host(lsd_host_ctx* c, uint32_t* host_keys, uint64_t n)
{
    const int rc = lsd_sort_host_async(c, host_keys, n);
    if (rc != LSD_OK || n == 0) return rc;
    return lsd_host_ctx_wait(c);
}

LSD_API int lsd_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return LSD_ERR_INVALID_VALUE;
    LSD_CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return LSD_OK;
}
LSD_API int lsd_host_free(void* ptr)
{
    if (ptr) LSD_CUDA_TRY(cudaFreeHost(ptr));
    return LSD_OK;
}

}  // extern "C"
