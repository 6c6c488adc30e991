// One-process-per-GPU communicator for the sharded paths: NCCL over NVLink / NVSwitch, resolved at run
// time (dlopen) so that single-GPU use never needs the NCCL library.  All calls are stream ordered.
#pragma once
#include "common.cuh"

struct LocalHub;        // in-process transport (comm.cu): several ranks as threads of one process, for tests
struct FdfdComm {
    void* nccl;         // ncclComm_t (null for an in-process communicator)
    int rank, world;
    LocalHub* hub;      // non-null: in-process transport
    int group_depth;    // comm_group_begin nesting
    void* pending;      // in-process transport: operations queued inside a group
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
// several point-to-point operations fused into one NCCL group (no ordering deadlock between peers); the
// sub-group collectives of the distributed fronts (broadcast, all-gather over 2/4/8 ranks) are built from these
int comm_group_begin(FdfdComm* c);
int comm_group_end(FdfdComm* c);
// In-process communicators: `world` ranks that live in ONE process (one thread each, same or different devices).
// Same call surface as the NCCL one; transfers are device-to-device copies behind a rendezvous.  This is how the
// multi-rank code paths are exercised on a single-GPU box.
int comm_create_local(FdfdComm** out, int world);     // out[world]; destroy every entry with comm_destroy
void comm_abort(FdfdComm* c);                          // wake every rank blocked in the hub with an error
