// NCCL plumbing for item sharding (one process per GPU).  libnccl is opened lazily with dlopen so that a single-GPU
// run has no NCCL dependency at all; only the logP all-reduce of the theta step goes through it.
#pragma once
#include "common.cuh"

namespace gpirt {

struct Comm {
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
};

int comm_unique_id(void* out128);
int comm_init(Comm& c, int rank, int world, const void* unique_id128);
int comm_allreduce_sum_f64(Comm& c, double* buf, size_t count, cudaStream_t stream);
// in-place all-gather: rank r contributes buf[r * count_per_rank, (r+1) * count_per_rank)
int comm_allgather_f64(Comm& c, double* buf, size_t count_per_rank, cudaStream_t stream);
void comm_destroy(Comm& c);
int comm_shutdown();    // destroys the cached communicator (refused while a live sampler still holds it)

}  // namespace gpirt
