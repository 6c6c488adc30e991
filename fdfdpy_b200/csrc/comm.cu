// NCCL binding of the sharded paths (see comm.cuh).  The prototypes come from <nccl.h>; the symbols are
// looked up with dlsym so the library neither links against nor requires NCCL unless a communicator is made.
#include <dlfcn.h>
#include <nccl.h>
#include "comm.cuh"

namespace {
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
} g_nccl;

#define NCCL_CHECK(call)                                                                          \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            snprintf(g_fdfd_err, sizeof(g_fdfd_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");          \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

template <class F>
bool sym(F& fn, const char* name) {
    fn = reinterpret_cast<F>(dlsym(g_nccl.handle, name));
    return fn != nullptr;
}
}  // namespace

int comm_load(const char* path) {
    if (g_nccl.handle) return 0;
    void* h = nullptr;
    if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // e.g. already loaded by torch
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) FDFD_FAIL("cannot load libnccl.so.2: %s", dlerror());
    g_nccl.handle = h;
    bool ok = sym(g_nccl.GetUniqueId, "ncclGetUniqueId") && sym(g_nccl.CommInitRank, "ncclCommInitRank") &&
              sym(g_nccl.CommDestroy, "ncclCommDestroy") && sym(g_nccl.GetErrorString, "ncclGetErrorString") &&
              sym(g_nccl.Send, "ncclSend") && sym(g_nccl.Recv, "ncclRecv") && sym(g_nccl.AllReduce, "ncclAllReduce") &&
              sym(g_nccl.GroupStart, "ncclGroupStart") && sym(g_nccl.GroupEnd, "ncclGroupEnd");
    if (!ok) {
        g_nccl.handle = nullptr;
        FDFD_FAIL("libnccl.so.2 lacks a required symbol");
    }
    return 0;
}

int comm_unique_id(void* id128) {
    if (comm_load(nullptr)) return -1;
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int comm_create(FdfdComm** out, const void* id128, int rank, int world) {
    if (comm_load(nullptr)) return -1;
    if (world < 1 || rank < 0 || rank >= world) FDFD_FAIL("bad rank %d of %d", rank, world);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    NCCL_CHECK(g_nccl.CommInitRank(&comm, world, id, rank));
    FdfdComm* c = new FdfdComm();
    c->nccl = comm; c->rank = rank; c->world = world;
    *out = c;
    return 0;
}

void comm_destroy(FdfdComm* c) {
    if (!c) return;
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl);
    delete c;
}

int comm_send(FdfdComm* c, const void* buf, size_t count, int peer, cudaStream_t st) {
    NCCL_CHECK(g_nccl.Send(buf, count, ncclDouble, peer, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_recv(FdfdComm* c, void* buf, size_t count, int peer, cudaStream_t st) {
    NCCL_CHECK(g_nccl.Recv(buf, count, ncclDouble, peer, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_sendrecv(FdfdComm* c, const void* sbuf, int send_peer, void* rbuf, int recv_peer, size_t count,
                  cudaStream_t st) {
    NCCL_CHECK(g_nccl.GroupStart());
    ncclResult_t r1 = g_nccl.Send(sbuf, count, ncclDouble, send_peer, (ncclComm_t)c->nccl, st);
    ncclResult_t r2 = g_nccl.Recv(rbuf, count, ncclDouble, recv_peer, (ncclComm_t)c->nccl, st);
    NCCL_CHECK(g_nccl.GroupEnd());
    NCCL_CHECK(r1);
    NCCL_CHECK(r2);
    return 0;
}
int comm_halo_exchange(FdfdComm* c, const void* first, const void* last, void* halo_lo, void* halo_hi, int lower,
                       int upper, size_t count, cudaStream_t st) {
    // one NCCL group = one fused kernel for all four transfers.  Per peer the posting order pairs the
    // messages: with two ranks (lower == upper) the peer's first row meets the halo_hi receive, its last row
    // the halo_lo receive.
    NCCL_CHECK(g_nccl.GroupStart());
    ncclResult_t r1 = g_nccl.Send(first, count, ncclDouble, lower, (ncclComm_t)c->nccl, st);
    ncclResult_t r2 = g_nccl.Send(last, count, ncclDouble, upper, (ncclComm_t)c->nccl, st);
    ncclResult_t r3 = g_nccl.Recv(halo_hi, count, ncclDouble, upper, (ncclComm_t)c->nccl, st);
    ncclResult_t r4 = g_nccl.Recv(halo_lo, count, ncclDouble, lower, (ncclComm_t)c->nccl, st);
    NCCL_CHECK(g_nccl.GroupEnd());
    NCCL_CHECK(r1); NCCL_CHECK(r2); NCCL_CHECK(r3); NCCL_CHECK(r4);
    return 0;
}
int comm_allreduce_sum(FdfdComm* c, void* buf, size_t count, cudaStream_t st) {
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_allreduce_max_i32(FdfdComm* c, int* buf, size_t count, cudaStream_t st) {
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclInt32, ncclMax, (ncclComm_t)c->nccl, st));
    return 0;
}
