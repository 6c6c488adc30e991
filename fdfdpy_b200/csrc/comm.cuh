// One-process-per-GPU communicator for the sharded paths: NCCL over NVLink / NVSwitch, resolved at run
// time (dlopen) so that single-GPU use never needs the NCCL library.  All calls are stream ordered.
#pragma once
#include "common.cuh"

struct FdfdComm {
    void* nccl;         // ncclComm_t
    int rank, world;
};

int comm_load(const char* libnccl_path);                  // idempotent; nullptr = default search order
int comm_unique_id(void* id128);
int comm_create(FdfdComm** out, const void* id128, int rank, int world);
void comm_destroy(FdfdComm* c);
// counts are in doubles (a complex number is two)
int comm_send(FdfdComm* c, const void* buf, size_t count, int peer, cudaStream_t st);
int comm_recv(FdfdComm* c, void* buf, size_t count, int peer, cudaStream_t st);
// both directions of one neighbour exchange fused in a single NCCL group (no ordering deadlock)
int comm_sendrecv(FdfdComm* c, const void* sbuf, int send_peer, void* rbuf, int recv_peer, size_t count,
                  cudaStream_t st);
// both boundary rows of a slab to its two neighbours and both halo rows back, in a single NCCL group
int comm_halo_exchange(FdfdComm* c, const void* first, const void* last, void* halo_lo, void* halo_hi, int lower,
                       int upper, size_t count, cudaStream_t st);
int comm_allreduce_sum(FdfdComm* c, void* buf, size_t count, cudaStream_t st);     // in place, float64
int comm_allreduce_max_i32(FdfdComm* c, int* buf, size_t count, cudaStream_t st);  // in place
