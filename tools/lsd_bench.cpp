// lsd_bench -- command-line harness over liblsdsort's C ABI that reproduces the reference's Test*/Benchmark* drivers
// (LSDRadixSort/LSDRadixSort.cu:1029-1185) and prints the same stdout blocks as BenchmarkLSDRadixSort.md,
// BenchmarkPrefixSum.md and BenchmarkBuildHistogram.md, so the outputs can be diffed field by field:
//
//     lsd_bench sort            [--elems N[,N..]] [--blocks B[,B..]] [--rs R[,R..]] [--seed S]   (BenchmarkGPULSDRadixSort, :1138)
//     lsd_bench prefix_sum      [--elems ..] [--blocks ..]                                      (BenchmarkGPUPrefixSum,   :1083)
//     lsd_bench build_histogram [--elems ..] [--blocks ..] [--rs ..]                            (BenchmarkBuildHistogram, :1124)
//     lsd_bench pairs           [--elems ..] [--rs ..]                                          (key-value sort; no reference twin)
//     lsd_bench sort64          [--elems ..]                                                    (64-bit keys; no reference twin)
//
// Defaults are the reference's sweep axes (elems 32 Mi, blocks 32..1024, rs 1,2,4,8; :1029-1062).  Differences, all on
// purpose: the "CPU" line times this harness's own host checker (std::sort, a running sum, a counting loop -- what the
// reference checks against at :97, :128, :643), because the library has no CPU path; the reference's SKIP rules
// (:940, :953, :727) do not apply (the workspace is a few MiB, so nothing is skipped); a failed check prints
// "CHECK FAILED" and exits 1 instead of crashing.  Harness only: nothing here is on the product path.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "lsdsort.h"

#define CK(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            std::fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)
#define LSD(expr)                                                                             \
    do {                                                                                      \
        int s_ = (expr);                                                                      \
        if (s_ != LSD_OK) {                                                                   \
            std::fprintf(stderr, "%s:%d lsd status %d (%s)\n", __FILE__, __LINE__, s_, lsd_status_string(s_)); \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }
double gib(double bytes) { return bytes / (1024.0 * 1024.0 * 1024.0); }

std::vector<long long> parse_list(const char* s)
{
    std::vector<long long> v;
    const char* p = s;
    while (*p) {
        char* end = nullptr;
        long long x = std::strtoll(p, &end, 0);
        if (end == p) break;
        // allow 2^k written as e.g. 1<<25 is not parsed; accept "32Mi"-style suffixes
        if (*end == 'K' || *end == 'k') { x <<= 10; ++end; }
        else if (*end == 'M' || *end == 'm') { x <<= 20; ++end; }
        else if (*end == 'G' || *end == 'g') { x <<= 30; ++end; }
        if (*end == 'i') ++end;
        v.push_back(x);
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',' && *end != 0) break;
    }
    return v;
}

struct Args {
    std::vector<long long> elems{1024LL * 1024 * 32};         // .cu:1029-1039
    std::vector<long long> blocks{32, 64, 128, 256, 512, 1024};  // .cu:1041-1050
    std::vector<long long> rs{1, 2, 4, 8};                     // .cu:1052-1059
    unsigned seed = 0;
    int reps = 1;
};

std::vector<uint32_t> make_input(size_t n, unsigned seed)
{
    // the reference draws from std::default_random_engine(0) over [0, UINT32_MAX] (Utils.h RNG); a named engine keeps
    // the input reproducible across standard libraries
    std::mt19937 eng(seed);
    std::vector<uint32_t> a(n);
    for (auto& x : a) x = eng();
    return a;
}

struct DeviceBuf {
    void* p = nullptr;
    explicit DeviceBuf(size_t bytes) { CK(cudaMalloc(&p, bytes ? bytes : 256)); }
    ~DeviceBuf() { cudaFree(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

float timed(cudaStream_t s, int reps, const std::function<void()>& restore, const std::function<void()>& run)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 0.f;
    for (int i = 0; i < reps; ++i) {
        restore();
        CK(cudaEventRecord(e0, s));
        run();
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i == 0 || ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

bool report_check(bool ok)
{
    if (!ok) std::cout << "CHECK FAILED" << std::endl;
    return ok;
}

// ---- TestGPULSDRadixSort (.cu:912-1027) ---------------------------------------------------------------------------
bool test_sort(size_t count, int block, int r, const Args& a, bool pairs)
{
    std::cout << (pairs ? "-- Test GPU LSD Radix Sort (key-value) --" : "-- Test GPU LSD Radix Sort --") << std::endl;
    const size_t size = count * sizeof(uint32_t);
    const size_t ws_bytes = pairs ? lsd_sort_pairs_workspace_bytes(count, r, block, nullptr) : lsd_sort_workspace_bytes(count, r, block);
    std::cout << "Elements: " << gib((double)size) << " GB" << std::endl;
    std::cout << "Histograms: " << gib((double)ws_bytes) << " GB" << std::endl;  // the whole workspace: there is no 3*G*H array
    std::cout << "Block Sums: " << 0 << " GB" << std::endl;
    std::cout << "Block Size: " << block << std::endl;
    std::cout << "R: " << r << std::endl;
    if (count > 0 && ws_bytes == 0) {
        std::cout << "SKIP: unsupported configuration" << std::endl;
        return true;
    }
    std::vector<uint32_t> h_in = make_input(count, a.seed), want = h_in, got(count), got_v;
    std::vector<uint32_t> idx;
    auto t0 = Clock::now();
    if (pairs) {
        idx.resize(count);
        for (size_t i = 0; i < count; ++i) idx[i] = (uint32_t)i;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return h_in[x] < h_in[y]; });
        for (size_t i = 0; i < count; ++i) want[i] = h_in[idx[i]];
    } else {
        std::sort(want.begin(), want.end());
    }
    const double cpu_ms = ms_since(t0);
    std::cout << "CPU " << cpu_ms << " ms" << std::endl;

    DeviceBuf d_a(size), d_b(size), d_src(size), d_ws(ws_bytes), d_v(pairs ? size : 0), d_vb(pairs ? size : 0), d_iota(pairs ? size : 0);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    CK(cudaMemcpy(d_src.p, h_in.data(), size, cudaMemcpyHostToDevice));
    if (pairs) {
        std::vector<uint32_t> iota(count);
        for (size_t i = 0; i < count; ++i) iota[i] = (uint32_t)i;
        CK(cudaMemcpy(d_iota.p, iota.data(), size, cudaMemcpyHostToDevice));
    }
    const float gpu_ms = timed(s, a.reps,
        [&] {
            CK(cudaMemcpyAsync(d_a.p, d_src.p, size, cudaMemcpyDeviceToDevice, s));
            if (pairs) CK(cudaMemcpyAsync(d_v.p, d_iota.p, size, cudaMemcpyDeviceToDevice, s));
        },
        [&] {
            if (pairs)
                LSD(lsd_sort_pairs(d_a.as<uint32_t>(), d_v.as<uint32_t>(), d_b.as<uint32_t>(), d_vb.as<uint32_t>(), count, r, block,
                                   d_ws.p, ws_bytes, nullptr, (lsd_stream_t)s));
            else
                LSD(lsd_sort(d_a.as<uint32_t>(), d_b.as<uint32_t>(), count, r, block, d_ws.p, ws_bytes, (lsd_stream_t)s));
        });
    CK(cudaMemcpy(got.data(), d_a.p, size, cudaMemcpyDeviceToHost));
    bool ok = got == want;
    if (pairs) {
        got_v.resize(count);
        CK(cudaMemcpy(got_v.data(), d_v.p, size, cudaMemcpyDeviceToHost));
        ok = ok && got_v == idx;
    }
    std::cout << "GPU " << gpu_ms << " ms" << std::endl;
    std::cout << "Speedup: x" << cpu_ms / gpu_ms << std::endl;
    cudaStreamDestroy(s);
    return report_check(ok);
}

// ---- 64-bit keys (lsd_sort64; no reference twin: the reference sorts uint32 only, .cu:62, :839) ---------------------
bool test_sort64(size_t count, const Args& a)
{
    std::cout << "-- Test GPU LSD Radix Sort (64-bit keys) --" << std::endl;
    const size_t size = count * sizeof(uint64_t);
    const size_t ws_bytes = lsd_sort64_workspace_bytes(count);
    std::cout << "Elements: " << gib((double)size) << " GB" << std::endl;
    std::cout << "Histograms: " << gib((double)ws_bytes) << " GB" << std::endl;
    std::cout << "Block Sums: " << 0 << " GB" << std::endl;
    std::cout << "Block Size: " << 0 << std::endl;
    std::cout << "R: " << 8 << std::endl;
    std::mt19937_64 eng(a.seed);
    std::vector<uint64_t> h_in(count), got(count);
    for (auto& k : h_in) k = eng();
    std::vector<uint64_t> want = h_in;
    auto t0 = Clock::now();
    std::sort(want.begin(), want.end());
    const double cpu_ms = ms_since(t0);
    std::cout << "CPU " << cpu_ms << " ms" << std::endl;
    DeviceBuf d_a(size), d_b(size), d_src(size), d_ws(ws_bytes);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    CK(cudaMemcpy(d_src.p, h_in.data(), size, cudaMemcpyHostToDevice));
    const float gpu_ms = timed(s, a.reps, [&] { CK(cudaMemcpyAsync(d_a.p, d_src.p, size, cudaMemcpyDeviceToDevice, s)); },
                               [&] { LSD(lsd_sort64(d_a.as<uint64_t>(), d_b.as<uint64_t>(), count, LSD_KEY_U64, d_ws.p, ws_bytes, (lsd_stream_t)s)); });
    CK(cudaMemcpy(got.data(), d_a.p, size, cudaMemcpyDeviceToHost));
    std::cout << "GPU " << gpu_ms << " ms" << std::endl;
    std::cout << "Speedup: x" << cpu_ms / gpu_ms << std::endl;
    cudaStreamDestroy(s);
    return report_check(got == want);
}

// ---- TestGPUPrefixSum (.cu:304-371) -------------------------------------------------------------------------------
bool test_prefix_sum(size_t count, int block, const Args& a)
{
    std::cout << "-- Test exclusive prefix sum --" << std::endl;
    const size_t size = count * sizeof(uint32_t);
    std::vector<uint32_t> h_in = make_input(count, a.seed), want(count), got(count);
    auto t0 = Clock::now();
    uint32_t run = 0;
    for (size_t i = 0; i < count; ++i) { want[i] = run; run += h_in[i]; }
    const double cpu_ms = ms_since(t0);
    const size_t ws_bytes = lsd_prefix_sum_workspace_bytes(count, block);
    DeviceBuf d_a(size), d_src(size), d_ws(ws_bytes);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    CK(cudaMemcpy(d_src.p, h_in.data(), size, cudaMemcpyHostToDevice));
    const float gpu_ms = timed(s, a.reps,
        [&] { CK(cudaMemcpyAsync(d_a.p, d_src.p, size, cudaMemcpyDeviceToDevice, s)); },
        [&] { LSD(lsd_prefix_sum(d_a.as<uint32_t>(), count, block, d_ws.p, ws_bytes, (lsd_stream_t)s)); });
    CK(cudaMemcpy(got.data(), d_a.p, size, cudaMemcpyDeviceToHost));
    std::cout << "Prefix sum of " << gib((double)size) << " GB of data" << std::endl;
    std::cout << "Threads per block: " << block << std::endl;
    std::cout << "Prefix Sum Sequential: " << cpu_ms << " ms" << std::endl;
    std::cout << "GPU Prefix Sum: " << gpu_ms << " ms" << std::endl;
    std::cout << "Speedup: x" << cpu_ms / gpu_ms << std::endl;
    cudaStreamDestroy(s);
    return report_check(got == want);
}

// ---- TestBuildHistogram (.cu:704-793) -----------------------------------------------------------------------------
bool test_build_histogram(size_t count, int block, int r, int bit_group, const Args& a)
{
    std::cout << "-- Test Build Histogram --" << std::endl;
    const size_t size = count * sizeof(uint32_t);
    const size_t h_bytes = lsd_build_histogram_bytes(count, r, block);
    std::cout << "Elements: " << gib((double)size) << " GB" << std::endl;
    std::cout << "Histograms: " << gib((double)h_bytes) << " GB" << std::endl;
    std::cout << "Block Size: " << block << std::endl;
    std::cout << "R: " << r << std::endl;
    std::cout << "Bit Group: " << bit_group << std::endl;
    std::vector<uint32_t> h_in = make_input(count, a.seed);
    const size_t H = (size_t)1 << r, G = (count + block - 1) / block;
    std::vector<uint32_t> want(G * H, 0), got(G * H);
    auto t0 = Clock::now();
    for (size_t i = 0; i < count; ++i) want[(i / block) * H + ((h_in[i] >> (bit_group * r)) & (H - 1))] += 1;
    const double cpu_ms = ms_since(t0);
    std::cout << "CPU " << cpu_ms << " ms" << std::endl;
    DeviceBuf d_a(size), d_h(h_bytes);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    CK(cudaMemcpy(d_a.p, h_in.data(), size, cudaMemcpyHostToDevice));
    const float gpu_ms = timed(s, a.reps, [] {},
        [&] { LSD(lsd_build_histogram(d_a.as<uint32_t>(), count, r, bit_group, block, d_h.as<uint32_t>(), (lsd_stream_t)s)); });
    CK(cudaMemcpy(got.data(), d_h.p, h_bytes, cudaMemcpyDeviceToHost));
    std::cout << "GPU " << gpu_ms << " ms" << std::endl;
    std::cout << "Speedup: x" << cpu_ms / gpu_ms << std::endl;
    cudaStreamDestroy(s);
    return report_check(got == want);
}

void usage()
{
    std::cerr << "usage: lsd_bench {sort|pairs|sort64|prefix_sum|build_histogram} [--elems N,..] [--blocks B,..] [--rs R,..] "
                 "[--seed S] [--reps K]\n";
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 2) { usage(); return 64; }
    const std::string mode = argv[1];
    Args a;
    for (int i = 2; i < argc; ++i) {
        const std::string f = argv[i];
        if (i + 1 >= argc) { usage(); return 64; }
        const char* v = argv[++i];
        if (f == "--elems") a.elems = parse_list(v);
        else if (f == "--blocks") a.blocks = parse_list(v);
        else if (f == "--rs") a.rs = parse_list(v);
        else if (f == "--seed") a.seed = (unsigned)std::strtoul(v, nullptr, 0);
        else if (f == "--reps") a.reps = std::max(1, std::atoi(v));
        else { usage(); return 64; }
    }
    bool ok = true;
    if (mode == "sort" || mode == "pairs") {
        if (mode == "pairs" && argc == 2) a.blocks = {0};
        for (long long n : a.elems)
            for (long long b : a.blocks)
                for (long long r : a.rs) ok = test_sort((size_t)n, (int)b, (int)r, a, mode == "pairs") && ok;
    } else if (mode == "sort64") {
        for (long long n : a.elems) ok = test_sort64((size_t)n, a) && ok;
    } else if (mode == "prefix_sum") {
        for (long long n : a.elems)
            for (long long b : a.blocks) ok = test_prefix_sum((size_t)n, (int)b, a) && ok;
    } else if (mode == "build_histogram") {
        std::mt19937 eng(a.seed);
        for (long long n : a.elems)
            for (long long b : a.blocks)
                for (long long r : a.rs) {
                    const int bit_group = (int)(eng() % (32 / r));  // the reference draws RNG(0, 0, 32/r).Get() (.cu:1132)
                    ok = test_build_histogram((size_t)n, (int)b, (int)r, bit_group, a) && ok;
                }
    } else {
        usage();
        return 64;
    }
    return ok ? 0 : 1;
}
