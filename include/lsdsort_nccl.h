/*
 * lsdsort_nccl.h -- fills an lsd_multi_comm (include/lsdsort.h) from an ncclComm_t.
 *
 * Header-only on purpose: the two callbacks are compiled into the CALLER's translation unit, so liblsdsort itself
 * links no NCCL.  Usage (one process or thread per GPU):
 *
 *     lsd_nccl_comm state;
 *     lsd_multi_comm comm;
 *     lsd_multi_comm_from_nccl(&state, nccl_comm, rank, nranks, token_dev, &comm);   // token_dev: one int of device memory
 *     lsd_multi_ctx_create(&comm, recv, capacity, n_local, 8, &ctx, stream);
 *     lsd_sort_multi(ctx, keys, n_local, scratch, &n_out, stream);
 */
#ifndef LSDSORT_NCCL_H
#define LSDSORT_NCCL_H

#include <cuda_runtime.h>
#include <nccl.h>

#include "lsdsort.h"

typedef struct lsd_nccl_comm {
    ncclComm_t comm;
    int *token; /* one int of device memory: the barrier is a 1-element all-reduce, ordered on the stream */
} lsd_nccl_comm;

static int lsd_nccl_all_gather(void *ctx, const void *send, void *recv, size_t bytes, lsd_stream_t stream)
{
    lsd_nccl_comm *c = (lsd_nccl_comm *)ctx;
    return ncclAllGather(send, recv, bytes, ncclChar, c->comm, (cudaStream_t)stream) == ncclSuccess ? 0 : 1;
}

static int lsd_nccl_barrier(void *ctx, lsd_stream_t stream)
{
    lsd_nccl_comm *c = (lsd_nccl_comm *)ctx;
    return ncclAllReduce(c->token, c->token, 1, ncclInt, ncclSum, c->comm, (cudaStream_t)stream) == ncclSuccess ? 0 : 1;
}

static inline void lsd_multi_comm_from_nccl(lsd_nccl_comm *state, ncclComm_t comm, int rank, int nranks, int *token_dev,
                                            lsd_multi_comm *out)
{
    state->comm = comm;
    state->token = token_dev;
    out->struct_bytes = (uint32_t)sizeof(lsd_multi_comm);
    out->rank = rank;
    out->nranks = nranks;
    out->all_gather = lsd_nccl_all_gather;
    out->barrier = lsd_nccl_barrier;
    out->ctx = state;
}

#endif /* LSDSORT_NCCL_H */
