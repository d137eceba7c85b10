// lsd_multi_check -- the multi-GPU sort called the way a C/C++ user of liblsdsort would call it: one host thread per GPU
// in ONE process, an ncclComm_t per GPU (ncclCommInitAll), include/lsdsort_nccl.h for the two collectives, lsd_sort_multi
// for everything else.  Checks the result against std::sort of the union of all ranks' keys and prints per-stage device
// times.  Harness only: nothing here is on the product path.
//
//     lsd_multi_check [--gpus N] [--log2n K (keys per rank)] [--kind uniform|sorted|equal|nibble|range17|offset] [--reps R]
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "lsdsort.h"
#include "lsdsort_nccl.h"

#define CK(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            std::fprintf(stderr, "%s:%d CUDA error %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));    \
            std::exit(2);                                                                                 \
        }                                                                                                 \
    } while (0)

static uint32_t hash32(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}

static void make_keys(std::vector<uint32_t>& k, const std::string& kind, int rank)
{
    for (size_t i = 0; i < k.size(); ++i) {
        const uint32_t u = hash32(((uint64_t)rank << 40) + i);
        if (kind == "sorted") k[i] = (uint32_t)(i * 7);
        else if (kind == "equal") k[i] = 0xDEADBEEFu;
        else if (kind == "nibble") k[i] = u & 0xFu;                  // BASELINE config 4: only the low nibble varies
        else if (kind == "range17") k[i] = u & 0x1FFFFu;             // keys below 2^17: the exchange window straddles two digits
        else if (kind == "offset") k[i] = 0x80000000u + u % 1000u;   // a thousand values next to 2^31
        else k[i] = u;
    }
}

int main(int argc, char** argv)
{
    int gpus = 0, log2n = 22, reps = 3;
    std::string kind = "uniform";
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--gpus")) gpus = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--log2n")) log2n = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--reps")) reps = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--kind")) kind = argv[i + 1];
    }
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (gpus <= 0) gpus = have;
    if (gpus > have || gpus < 1) {
        std::fprintf(stderr, "lsd_multi_check: %d GPUs requested, %d present\n", gpus, have);
        return 2;
    }
    const uint64_t n_local = 1ull << log2n;
    // the ordinary 25 % slack for everything the exchange window balances; "sorted" ramps differ per rank count
    const uint64_t capacity = kind == "sorted" ? n_local * gpus + 64 : n_local + n_local / 4 + 65536;
    std::vector<ncclComm_t> comms(gpus);
    std::vector<int> devs(gpus);
    for (int i = 0; i < gpus; ++i) devs[i] = i;
    if (ncclCommInitAll(comms.data(), gpus, devs.data()) != ncclSuccess) {
        std::fprintf(stderr, "ncclCommInitAll failed\n");
        return 2;
    }
    std::vector<std::vector<uint32_t>> in(gpus), out(gpus);
    std::vector<lsd_multi_stats> stats(gpus);
    std::atomic<int> failures{0};
    std::vector<std::thread> threads;
    for (int rank = 0; rank < gpus; ++rank) {
        threads.emplace_back([&, rank] {
            CK(cudaSetDevice(rank));
            cudaStream_t s;
            CK(cudaStreamCreate(&s));
            in[rank].resize(n_local);
            make_keys(in[rank], kind, rank);
            uint32_t *keys, *recv, *scratch;
            int* token;
            CK(cudaMalloc(&keys, n_local * 4));
            CK(cudaMalloc(&recv, capacity * 4));
            CK(cudaMalloc(&scratch, capacity * 4));
            CK(cudaMalloc(&token, sizeof(int)));
            CK(cudaMemset(token, 0, sizeof(int)));
            CK(cudaMemcpy(keys, in[rank].data(), n_local * 4, cudaMemcpyHostToDevice));
            lsd_nccl_comm state;
            lsd_multi_comm comm;
            lsd_multi_comm_from_nccl(&state, comms[rank], rank, gpus, token, &comm);
            lsd_multi_ctx* ctx = nullptr;
            int st = lsd_multi_ctx_create(&comm, recv, capacity, n_local, 8, &ctx, (lsd_stream_t)s);
            if (st != LSD_OK) {
                std::fprintf(stderr, "rank %d: lsd_multi_ctx_create: %s (cuda %d)\n", rank, lsd_status_string(st), lsd_last_cuda_error());
                ++failures;
                return;
            }
            lsd_multi_set_timing(ctx, 1);
            uint64_t n_out = 0;
            for (int rep = 0; rep < reps; ++rep) {
                CK(cudaMemsetAsync(recv, 0x5A, capacity * 4, s));  // poison: nothing of the previous repetition may survive
                st = lsd_sort_multi(ctx, keys, n_local, scratch, &n_out, (lsd_stream_t)s);
                if (st != LSD_OK) {
                    std::fprintf(stderr, "rank %d: lsd_sort_multi: %s\n", rank, lsd_status_string(st));
                    ++failures;
                    return;
                }
            }
            lsd_multi_last_stats(ctx, &stats[rank]);
            out[rank].resize(n_out);
            CK(cudaMemcpyAsync(out[rank].data(), recv, n_out * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            // barrier through NCCL before tearing the peer mappings down
            comm.barrier(comm.ctx, (lsd_stream_t)s);
            CK(cudaStreamSynchronize(s));
            lsd_multi_ctx_destroy(ctx);
            cudaFree(keys); cudaFree(recv); cudaFree(scratch); cudaFree(token);
            cudaStreamDestroy(s);
        });
    }
    for (auto& t : threads) t.join();
    for (auto c : comms) ncclCommDestroy(c);
    if (failures.load()) return 1;
    std::vector<uint32_t> want, got;
    for (int r = 0; r < gpus; ++r) {
        want.insert(want.end(), in[r].begin(), in[r].end());
        got.insert(got.end(), out[r].begin(), out[r].end());
    }
    std::sort(want.begin(), want.end());
    const bool ok = want == got;
    std::printf("-- lsd_sort_multi over the C ABI (threads + ncclCommInitAll) --\nGPUs: %d\nKeys per rank: %llu\nKind: %s\n", gpus,
                (unsigned long long)n_local, kind.c_str());
    for (int r = 0; r < gpus; ++r)
    {
        char where[64];
        if (stats[r].exchange_shift == 0xFFFFFFFFu) std::snprintf(where, sizeof(where), "all keys equal: nothing exchanged");
        else std::snprintf(where, sizeof(where), "buckets %u..%u of bits %u..%u", stats[r].first_bucket, stats[r].last_bucket,
                           stats[r].exchange_shift, stats[r].exchange_shift + 7);
        std::printf("rank %d: owns %llu keys (%s), sent %.1f MB | plan %.3f ms, exchange %.3f ms, local sort %.3f ms\n", r,
                    (unsigned long long)stats[r].n_out, where, stats[r].sent_bytes / 1e6, stats[r].plan_ms, stats[r].exchange_ms,
                    stats[r].sort_ms);
    }
    std::printf("%s\n", ok ? "CHECK PASSED: concatenated rank slices == std::sort of all keys" : "CHECK FAILED");
    return ok ? 0 : 1;
}
