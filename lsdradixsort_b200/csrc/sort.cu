// sort.cu -- host orchestration of the LSD sort: workspace layout, device-side plan, pass loop.
//
// Replaces GPULSDRadixSort (LSDRadixSort.cu:839-910).  Where the reference launches ~16 kernels
// plus a D2D copy per pass and creates/destroys two streams per call, this enqueues, on the
// caller's stream and without any host synchronisation:
//     memset(workspace header) -> digit_hist_kernel (zeroes the look-back records on the side) -> plan_kernel -> onesweep_kernel x passes -> copy_back_kernel
// Pass skipping and ping-pong parity are decided ON THE DEVICE by plan_kernel (every pass kernel
// reads the plan and exits at once if its digit is constant), so the sequence is fixed, fully
// asynchronous and CUDA-graph capturable.
#include <algorithm>

#include "onesweep.cuh"
#include "sort.h"

namespace lsd {

// -------------------------------------------------------------------------------------
// plan_kernel: one CTA.  For every pass: skip flag (some bucket holds all n keys), exclusive
// scan of the digit histogram -> absolute bucket bases (uint64), ping-pong parity.
// -------------------------------------------------------------------------------------
constexpr int kPlanThreads = 256;

__global__ void __launch_bounds__(kPlanThreads)
plan_kernel(const uint64_t* __restrict__ hist, uint64_t* __restrict__ bases, SortPlan* __restrict__ plan, uint64_t n,
            int passes, int H, int disable_skip)
{
    __shared__ uint64_t s_warp[kPlanThreads / 32];
    __shared__ uint32_t s_parity;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_parity = 0;
    __syncthreads();
    uint32_t executed = 0, first = kMaxPasses, last = kMaxPasses;
    for (int p = 0; p < passes; ++p) {
        const uint64_t v = (int)tid < H ? hist[p * H + tid] : 0ull;
        const int full = __syncthreads_or(v == n);
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint64_t prefix = 0;
        for (uint32_t w = 0; w < warp; ++w) prefix += s_warp[w];
        if ((int)tid < H) bases[(size_t)(2 * p) * H + tid] = prefix + incl - v;
        if (tid == 0) {
            const uint32_t skip = (full && !disable_skip) ? 1u : 0u;
            plan->skip[p] = skip;
            plan->src_is_scratch[p] = s_parity;
            if (!skip) {
                s_parity ^= 1u;
                ++executed;
                if (first == (uint32_t)kMaxPasses) first = (uint32_t)p;
                last = (uint32_t)p;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        plan->result_in_scratch = s_parity;
        plan->executed_passes = executed;
        plan->first_pass = first;
        plan->last_pass = last;
    }
}

// copy_back_kernel: only does work when an odd number of passes ran (result sits in scratch).
__global__ void __launch_bounds__(256)
copy_back_kernel(uint32_t* __restrict__ keys, const uint32_t* __restrict__ scratch, uint64_t n,
                 const SortPlan* __restrict__ plan)
{
    if (!plan->result_in_scratch) return;
    const uint64_t nvec = n >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)
        reinterpret_cast<uint4*>(keys)[i] = reinterpret_cast<const uint4*>(scratch)[i];
    if (blockIdx.x == 0) {
        const uint64_t t = (nvec << 2) + threadIdx.x;
        if (t < n) keys[t] = scratch[t];
    }
}

// -------------------------------------------------------------------------------------
// shape selection and workspace layout
// -------------------------------------------------------------------------------------
static const OnesweepLauncher* table_for(int r, int* count)
{
    switch (r) {
        case 1: return onesweep_table_r1(count);
        case 2: return onesweep_table_r2(count);
        case 4: return onesweep_table_r4(count);
        case 8: return onesweep_table_r8(count);
    }
    *count = 0;
    return nullptr;
}

// the launch entry a sort of this flavour needs (nullptr: this shape does not have it)
static int (*pass_fn(const OnesweepLauncher& k, bool pairs, bool typed))(const PassArgs&, cudaStream_t)
{
    return typed ? (pairs ? k.launch_pairs_typed : k.launch_typed) : (pairs ? k.launch_pairs : k.launch);
}

const OnesweepLauncher* select_launcher(int r, int block, uint32_t variant, bool pairs, bool typed)
{
    int count = 0;
    const OnesweepLauncher* t = table_for(r, &count);
    if (!t) return nullptr;
    if (variant != 0) return variant < (uint32_t)count ? &t[variant] : nullptr;
    // `block` is the reference's threads-per-block knob (LSDRadixSort.cu:839).  Here it is a HINT: every value gets the
    // tuned default shape -- the first entry that has the requested form (entry 0 has every form for r = 8; for r < 8 the
    // key-value forms live in the warp-multisplit entry behind it) -- so that a drop-in caller passing the reference's
    // B = 128 ... 1024 is not slower than one passing 0.  Exact shapes are selected with lsd_sort_options.variant.
    (void)block;
    for (int i = 0; i < count; ++i)
        if (pass_fn(t[i], pairs, typed)) return &t[i];
    return nullptr;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int make_layout(uint64_t n, int r, int block, const lsd_sort_options* opt, SortLayout* L, bool pairs)
{
    if (!accepted_radix(r)) return LSD_ERR_INVALID_VALUE;
    r = exec_radix(r);  // composite digit widths (11, 16, ...): the full sort runs the 8-bit schedule, same result
    if (block < 0 || block > 1024) return LSD_ERR_INVALID_VALUE;
    if (n > (1ull << 32)) return LSD_ERR_UNSUPPORTED;  // key positions are 32-bit (n = 2^32 included: the last position is 2^32 - 1)
    const uint32_t variant = opt ? opt->variant : 0u;
    const uint32_t key_type = opt ? opt->key_type : 0u;
    if (key_type > LSD_KEY_F32) return LSD_ERR_INVALID_VALUE;
    const OnesweepLauncher* k = select_launcher(r, block, variant, pairs, key_type != 0u);
    if (!k) return LSD_ERR_INVALID_VALUE;
    if (!pass_fn(*k, pairs, key_type != 0u)) return LSD_ERR_UNSUPPORTED;  // this tuning variant lacks the requested form
    L->k = k;
    L->passes = 32 / r;
    L->H = 1 << r;
    uint32_t portion = k->portion_max;
    if (opt && opt->portion_keys) {
        // round the requested portion down to whole tiles (at least one tile)
        uint64_t want = std::max<uint64_t>(k->tile, (uint64_t)opt->portion_keys / k->tile * k->tile);
        portion = (uint32_t)std::min<uint64_t>(want, k->portion_max);
    }
    L->portion_keys = portion;
    L->portions = n ? (n + portion - 1) / portion : 0;
    L->total_tiles = 0;
    for (uint64_t q = 0; q < L->portions; ++q) {
        const uint64_t keys = std::min<uint64_t>(portion, n - q * portion);
        L->total_tiles += (keys + k->tile - 1) / k->tile;
    }
    size_t off = 0;
    L->off_plan = off;     off = align_up(off + sizeof(SortPlan), 256);
    L->off_hist = off;     off = align_up(off + sizeof(uint64_t) * L->passes * L->H, 256);
    L->off_bases = off;    off = align_up(off + sizeof(uint64_t) * 2 * L->passes * L->H, 256);
    L->off_tickets = off;  off = align_up(off + sizeof(uint32_t) * L->passes * std::max<uint64_t>(1, L->portions), 256);
    L->off_lookback = off; off = align_up(off + sizeof(uint32_t) * L->passes * L->total_tiles * L->H, 256);
    L->total_bytes = off;
    return LSD_OK;
}

// -------------------------------------------------------------------------------------
// the sort
// -------------------------------------------------------------------------------------
int sort_enqueue(uint32_t* keys, uint32_t* scratch, uint64_t n, int r, int block, void* ws, size_t ws_bytes,
                 const lsd_sort_options* opt, cudaStream_t s, cudaEvent_t* events, int* launches, uint32_t* vals,
                 uint32_t* vals_scratch)
{
    const bool pairs = vals != nullptr || vals_scratch != nullptr;
    SortLayout L;
    const int st = make_layout(n, r, block, opt, &L, pairs);
    if (st != LSD_OK) return st;
    r = exec_radix(r);
    if (launches) *launches = 0;
    if (n == 0) return LSD_OK;
    if (!keys || !scratch || !ws) return LSD_ERR_INVALID_VALUE;
    if (pairs && (!vals || !vals_scratch)) return LSD_ERR_INVALID_VALUE;
    if (ws_bytes < L.total_bytes) return LSD_ERR_WORKSPACE_TOO_SMALL;
    if (!aligned_to(keys, 16) || !aligned_to(scratch, 16) || !aligned_to(ws, 256)) return LSD_ERR_ALIGNMENT;
    if (pairs && (!aligned_to(vals, 16) || !aligned_to(vals_scratch, 16))) return LSD_ERR_ALIGNMENT;
    const uint32_t key_type = opt ? opt->key_type : 0u;

    char* w = static_cast<char*>(ws);
    SortPlan* plan = reinterpret_cast<SortPlan*>(w + L.off_plan);
    uint64_t* hist = reinterpret_cast<uint64_t*>(w + L.off_hist);
    uint64_t* bases = reinterpret_cast<uint64_t*>(w + L.off_bases);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(w + L.off_tickets);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(w + L.off_lookback);

    int ev = 0, nl = 0;
    if (events) LSD_CUDA_TRY(cudaEventRecord(events[ev++], s));
    // plan, histograms, bases and tickets are zeroed here; the look-back records (~0.5 B per key: 131 MB at 2^28 keys) by the
    // histogram kernel, which is bound by its shared atomics and has the HBM bandwidth to spare
    LSD_CUDA_TRY(cudaMemsetAsync(ws, 0, L.off_lookback, s));
    int rc = launch_digit_histograms(keys, n, r, hist, s, key_type, w + L.off_lookback, L.total_bytes - L.off_lookback);
    if (rc != LSD_OK) return rc;
    ++nl;
    plan_kernel<<<1, kPlanThreads, 0, s>>>(hist, bases, plan, n, L.passes, L.H, opt ? (int)opt->disable_skip : 0);
    LSD_LAUNCH_CHECK();
    ++nl;
    if (events) LSD_CUDA_TRY(cudaEventRecord(events[ev++], s));

    for (int p = 0; p < L.passes; ++p) {
        uint32_t* lb = lookback + (size_t)p * L.total_tiles * L.H;
        for (uint64_t q = 0; q < L.portions; ++q) {
            const uint64_t pbase = q * (uint64_t)L.portion_keys;
            const uint32_t pkeys = (uint32_t)std::min<uint64_t>(L.portion_keys, n - pbase);
            PassArgs a;
            a.keys = keys;
            a.scratch = scratch;
            a.plan = plan;
            a.bases_in = bases + (size_t)(2 * p + (q & 1)) * L.H;
            a.bases_out = (q + 1 < L.portions) ? bases + (size_t)(2 * p + ((q + 1) & 1)) * L.H : nullptr;
            a.lookback = lb;
            a.ticket = tickets + (size_t)p * L.portions + q;
            a.portion_base = pbase;
            a.portion_keys = pkeys;
            a.tiles = (pkeys + L.k->tile - 1) / L.k->tile;
            a.pass = p;
            a.shift = p * r;
            a.trace = (opt && opt->debug_trace && p == L.passes - 1 && q == 0)
                          ? reinterpret_cast<unsigned long long*>(opt->debug_trace) : nullptr;
            a.dst_ptrs = nullptr;
            a.dst_seg = nullptr;
            a.vals = vals;
            a.vals_scratch = vals_scratch;
            a.key_type = key_type;
            a.digit_mask = (uint32_t)L.H - 1u;
            rc = pass_fn(*L.k, pairs, key_type != 0u)(a, s);
            if (rc != LSD_OK) return rc;
            ++nl;
            lb += (size_t)a.tiles * L.H;
        }
        if (events) LSD_CUDA_TRY(cudaEventRecord(events[ev++], s));
    }

    const int copy_grid = (int)std::min<uint64_t>((uint64_t)sm_count() * 8, ((n >> 2) + 255) / 256 + 1);
    copy_back_kernel<<<copy_grid, 256, 0, s>>>(keys, scratch, n, plan);
    LSD_LAUNCH_CHECK();
    ++nl;
    if (pairs) {
        copy_back_kernel<<<copy_grid, 256, 0, s>>>(vals, vals_scratch, n, plan);
        LSD_LAUNCH_CHECK();
        ++nl;
    }
    if (events) LSD_CUDA_TRY(cudaEventRecord(events[ev++], s));
    if (launches) *launches = nl;
    return LSD_OK;
}


// -------------------------------------------------------------------------------------
// single pass (lsd_sort_pass): digit histogram of the whole input -> plan with skipping disabled
// (pass `bit_group` reads `in`, writes `out`) -> one onesweep launch per portion.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPlanThreads)
single_pass_plan_kernel(const uint64_t* __restrict__ hist, uint64_t* __restrict__ bases, SortPlan* __restrict__ plan,
                        uint64_t* __restrict__ hist_out, int pass, int H, int zero_bases, const uint32_t* __restrict__ abort_flag)
{
    __shared__ uint64_t s_warp[kPlanThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t v = (int)tid < H ? hist[pass * H + tid] : 0ull;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t prefix = 0;
    for (uint32_t w = 0; w < warp; ++w) prefix += s_warp[w];
    if ((int)tid < H) {
        // peer-scatter mode: every bucket has its own destination pointer, offsets start at 0
        bases[(size_t)(2 * pass) * H + tid] = zero_bases ? 0ull : prefix + incl - v;
        if (hist_out) hist_out[tid] = prefix + incl - v;
    }
    if (tid == 0) {
        plan->skip[pass] = (abort_flag != nullptr && *abort_flag != 0u) ? 1u : 0u;
        plan->src_is_scratch[pass] = 0;  // PassArgs.keys = in, PassArgs.scratch = out
    }
}

int pass_enqueue(const uint32_t* in, uint32_t* out, uint64_t n, int r, int bit_group, int block, void* ws,
                 size_t ws_bytes, uint64_t* hist_out, cudaStream_t s, const uint64_t* dst_ptrs, const uint32_t* dst_seg,
                 const uint32_t* abort_flag, int peer_shift)
{
    SortLayout L;
    const int st = make_layout(n, r, block, nullptr, &L);
    if (st != LSD_OK) return st;
    if (bit_group < 0 || bit_group >= L.passes) return LSD_ERR_INVALID_VALUE;
    const bool peer = dst_ptrs != nullptr;
    if (peer_shift >= 0 && (!peer || peer_shift + r > 32)) return LSD_ERR_INVALID_VALUE;
    if (peer && !L.k->launch_peer) return LSD_ERR_UNSUPPORTED;
    if (n == 0) {
        if (hist_out) LSD_CUDA_TRY(cudaMemsetAsync(hist_out, 0, sizeof(uint64_t) * L.H, s));
        return LSD_OK;
    }
    if (!in || (!out && !peer) || !ws) return LSD_ERR_INVALID_VALUE;
    if (ws_bytes < L.total_bytes) return LSD_ERR_WORKSPACE_TOO_SMALL;
    if (!aligned_to(in, 16) || (!peer && !aligned_to(out, 16)) || !aligned_to(ws, 256)) return LSD_ERR_ALIGNMENT;

    char* w = static_cast<char*>(ws);
    SortPlan* plan = reinterpret_cast<SortPlan*>(w + L.off_plan);
    uint64_t* hist = reinterpret_cast<uint64_t*>(w + L.off_hist);
    uint64_t* bases = reinterpret_cast<uint64_t*>(w + L.off_bases);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(w + L.off_tickets);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(w + L.off_lookback);
    // only this pass's slice of tickets / look-back words is used, but zeroing up to its end is simplest
    const size_t lb_words_per_pass = (size_t)L.total_tiles * L.H;
    LSD_CUDA_TRY(cudaMemsetAsync(ws, 0, L.off_lookback + sizeof(uint32_t) * lb_words_per_pass, s));
    int rc = LSD_OK;
    if (!peer) {  // peer-scatter offsets start at 0 in every destination: no global histogram needed
        rc = launch_one_digit_histogram(in, n, r, bit_group, hist, s);  // the digit of this pass only: one shared atomic per key
        if (rc != LSD_OK) return rc;
    }
    single_pass_plan_kernel<<<1, kPlanThreads, 0, s>>>(hist, bases, plan, peer ? nullptr : hist_out, bit_group, L.H, peer ? 1 : 0, abort_flag);
    LSD_LAUNCH_CHECK();
    uint32_t* lb = lookback;
    for (uint64_t q = 0; q < L.portions; ++q) {
        const uint64_t pbase = q * (uint64_t)L.portion_keys;
        const uint32_t pkeys = (uint32_t)std::min<uint64_t>(L.portion_keys, n - pbase);
        PassArgs a;
        a.keys = const_cast<uint32_t*>(in);
        a.scratch = out;
        a.plan = plan;
        a.bases_in = bases + (size_t)(2 * bit_group + (q & 1)) * L.H;
        a.bases_out = (q + 1 < L.portions) ? bases + (size_t)(2 * bit_group + ((q + 1) & 1)) * L.H : nullptr;
        a.lookback = lb;
        a.ticket = tickets + q;
        a.portion_base = pbase;
        a.portion_keys = pkeys;
        a.tiles = (pkeys + L.k->tile - 1) / L.k->tile;
        a.pass = bit_group;
        a.shift = peer_shift >= 0 ? peer_shift : bit_group * r;  // peer-scatter: the digit may sit at any bit position
        a.trace = nullptr;
        a.dst_ptrs = dst_ptrs;
        a.dst_seg = dst_seg;
        a.vals = nullptr;
        a.vals_scratch = nullptr;
        a.key_type = 0;
        a.digit_mask = (uint32_t)L.H - 1u;
        rc = peer ? L.k->launch_peer(a, s) : L.k->launch(a, s);
        if (rc != LSD_OK) return rc;
        lb += (size_t)a.tiles * L.H;
    }
    return LSD_OK;
}

// -------------------------------------------------------------------------------------
// single pass on a COMPOSITE digit width (r in 3..16 other than 4 and 8): digit `bit_group` is bits
// [bit_group*r, min(32, (bit_group+1)*r)).  A stable counting-sort pass on a w-bit digit equals a stable pass on its
// low 8 bits followed by a stable pass on the remaining w-8 bits (the LSD argument applied inside the digit), so it
// runs as one or two sub-passes of the 8-bit kernel in its run-time (shift, mask) form; two sub-passes go through a
// temporary key array in the workspace.  hist_out gets the 2^r bucket starts of the whole digit.
// Workspace: [r = 8 sort layout][2^16 uint64 field histogram][n keys of temporary space].
// -------------------------------------------------------------------------------------
struct WideLayout {
    SortLayout L;
    size_t off_field, off_tmp, total_bytes;
};

static int make_wide_layout(uint64_t n, WideLayout* W)
{
    const int st = make_layout(n, 8, 0, nullptr, &W->L);
    if (st != LSD_OK) return st;
    size_t off = align_up(W->L.total_bytes, 256);
    W->off_field = off;  off = align_up(off + (sizeof(uint64_t) << 16), 256);
    W->off_tmp = off;    off = align_up(off + sizeof(uint32_t) * (size_t)n, 256);
    W->total_bytes = off;
    return LSD_OK;
}

size_t wide_pass_workspace_bytes(uint64_t n)
{
    WideLayout W;
    return make_wide_layout(n, &W) == LSD_OK ? W.total_bytes : 0;
}

static int sub_pass_enqueue(const uint32_t* in, uint32_t* out, uint64_t n, int shift, int bits, const SortLayout& L, char* w,
                            cudaStream_t s)
{
    SortPlan* plan = reinterpret_cast<SortPlan*>(w + L.off_plan);
    uint64_t* hist = reinterpret_cast<uint64_t*>(w + L.off_hist);
    uint64_t* bases = reinterpret_cast<uint64_t*>(w + L.off_bases);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(w + L.off_tickets);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(w + L.off_lookback);
    LSD_CUDA_TRY(cudaMemsetAsync(w, 0, L.off_lookback + sizeof(uint32_t) * (size_t)L.total_tiles * L.H, s));
    int rc = launch_field_histogram(in, n, shift, bits, hist, s);  // row 0 of the [4][256] area; bins >= 2^bits stay zero
    if (rc != LSD_OK) return rc;
    single_pass_plan_kernel<<<1, kPlanThreads, 0, s>>>(hist, bases, plan, nullptr, 0, L.H, 0, nullptr);
    LSD_LAUNCH_CHECK();
    uint32_t* lb = lookback;
    for (uint64_t q = 0; q < L.portions; ++q) {
        const uint64_t pbase = q * (uint64_t)L.portion_keys;
        const uint32_t pkeys = (uint32_t)std::min<uint64_t>(L.portion_keys, n - pbase);
        PassArgs a;
        a.keys = const_cast<uint32_t*>(in);
        a.scratch = out;
        a.plan = plan;
        a.bases_in = bases + (size_t)(q & 1) * L.H;
        a.bases_out = (q + 1 < L.portions) ? bases + (size_t)((q + 1) & 1) * L.H : nullptr;
        a.lookback = lb;
        a.ticket = tickets + q;
        a.portion_base = pbase;
        a.portion_keys = pkeys;
        a.tiles = (pkeys + L.k->tile - 1) / L.k->tile;
        a.pass = 0;
        a.shift = shift;
        a.trace = nullptr;
        a.dst_ptrs = nullptr;
        a.dst_seg = nullptr;
        a.vals = nullptr;
        a.vals_scratch = nullptr;
        a.key_type = 0;
        a.digit_mask = (1u << bits) - 1u;
        rc = L.k->launch(a, s);
        if (rc != LSD_OK) return rc;
        lb += (size_t)a.tiles * L.H;
    }
    return LSD_OK;
}

int pass_enqueue_wide(const uint32_t* in, uint32_t* out, uint64_t n, int r, int bit_group, void* ws, size_t ws_bytes,
                      uint64_t* hist_out, cudaStream_t s)
{
    if (!composite_radix(r)) return LSD_ERR_INVALID_VALUE;
    if (bit_group < 0 || bit_group >= digit_count(r)) return LSD_ERR_INVALID_VALUE;
    WideLayout W;
    const int st = make_wide_layout(n, &W);
    if (st != LSD_OK) return st;
    if (n == 0) {
        if (hist_out) LSD_CUDA_TRY(cudaMemsetAsync(hist_out, 0, sizeof(uint64_t) << r, s));
        return LSD_OK;
    }
    if (!in || !out || !ws) return LSD_ERR_INVALID_VALUE;
    if (ws_bytes < W.total_bytes) return LSD_ERR_WORKSPACE_TOO_SMALL;
    if (!aligned_to(in, 16) || !aligned_to(out, 16) || !aligned_to(ws, 256)) return LSD_ERR_ALIGNMENT;
    char* w = static_cast<char*>(ws);
    const int shift = bit_group * r;
    const int width = std::min(r, 32 - shift);
    int rc = LSD_OK;
    if (hist_out) {  // bucket starts of the whole digit; entries above 2^width (a narrower top digit) equal n
        uint64_t* field = reinterpret_cast<uint64_t*>(w + W.off_field);
        rc = launch_field_histogram(in, n, shift, width, field, s);
        if (rc != LSD_OK) return rc;
        if (width < r) LSD_CUDA_TRY(cudaMemsetAsync(field + ((size_t)1 << width), 0, (sizeof(uint64_t) << r) - (sizeof(uint64_t) << width), s));
        rc = launch_field_scan(field, r, s);
        if (rc != LSD_OK) return rc;
        LSD_CUDA_TRY(cudaMemcpyAsync(hist_out, field, sizeof(uint64_t) << r, cudaMemcpyDeviceToDevice, s));
    }
    if (width <= 8) return sub_pass_enqueue(in, out, n, shift, width, W.L, w, s);
    uint32_t* tmp = reinterpret_cast<uint32_t*>(w + W.off_tmp);
    rc = sub_pass_enqueue(in, tmp, n, shift, 8, W.L, w, s);
    if (rc != LSD_OK) return rc;
    return sub_pass_enqueue(tmp, out, n, shift + 8, width - 8, W.L, w, s);
}

// [digit_count(r)][2^r] uint64 histograms for a composite digit width: one read of the keys per digit
int launch_digit_histograms_wide(const uint32_t* keys, uint64_t n, int r, uint64_t* hist, cudaStream_t s)
{
    if (!composite_radix(r)) return LSD_ERR_INVALID_VALUE;
    const int digits = digit_count(r);
    LSD_CUDA_TRY(cudaMemsetAsync(hist, 0, (sizeof(uint64_t) << r) * digits, s));
    for (int i = 0; i < digits; ++i) {
        const int shift = i * r;
        const int rc = launch_field_histogram(keys, n, shift, std::min(r, 32 - shift), hist + ((size_t)i << r), s);
        if (rc != LSD_OK) return rc;
    }
    return LSD_OK;
}

}  // namespace lsd
