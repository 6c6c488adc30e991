// NCCL binding of the sharded paths (see comm.cuh).  The prototypes come from <nccl.h>; the symbols are
// looked up with dlsym so the library neither links against nor requires NCCL unless a communicator is made.
#include <dlfcn.h>
#include <nccl.h>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
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

// ------------------------------------------------------------------------------------------
// in-process transport
// ------------------------------------------------------------------------------------------
struct LocalMsg { const void* ptr; size_t count; };
struct LocalOp { int kind; const void* sbuf; void* rbuf; size_t count; int peer; cudaStream_t st; };   // kind 0 send, 1 recv
struct LocalHub {
    int world, refs;
    bool aborted;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::deque<LocalMsg>> box;      // box[src * world + dst]: posted, not yet consumed
    std::vector<unsigned long long> posted, consumed;   // per (src, dst) counters
};
static int local_timeout_s() {
    const char* e = getenv("FDFD_LOCAL_TIMEOUT_S");
    int v = e ? atoi(e) : 0;
    return v > 0 ? v : 120;
}

static int local_run(FdfdComm* c, std::vector<LocalOp>& ops) {
    LocalHub* h = c->hub;
    // 1. the data of every send must be complete before a peer may copy it
    for (auto& o : ops)
        if (o.kind == 0) FDFD_CHECK(cudaStreamSynchronize(o.st));
    std::vector<std::pair<int, unsigned long long>> mine;       // (slot, ticket) of my posts
    {
        std::lock_guard<std::mutex> lk(h->mu);
        for (auto& o : ops)
            if (o.kind == 0) {
                const int slot = c->rank * h->world + o.peer;
                h->box[slot].push_back({o.sbuf, o.count});
                mine.push_back({slot, ++h->posted[slot]});
            }
    }
    h->cv.notify_all();
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(local_timeout_s());
    // 2. receives, in the order they were listed (messages between one pair of ranks match in posting order)
    for (auto& o : ops) {
        if (o.kind != 1) continue;
        const int slot = o.peer * h->world + c->rank;
        LocalMsg m;
        {
            std::unique_lock<std::mutex> lk(h->mu);
            if (!h->cv.wait_until(lk, deadline, [&] { return h->aborted || !h->box[slot].empty(); }))
                FDFD_FAIL("in-process communicator: rank %d timed out waiting for rank %d", c->rank, o.peer);
            if (h->aborted) FDFD_FAIL("in-process communicator aborted (another rank failed)");
            m = h->box[slot].front();
            h->box[slot].pop_front();
        }
        if (m.count != o.count) {
            comm_abort(c);
            FDFD_FAIL("in-process communicator: rank %d expected %zu doubles from rank %d, got %zu", c->rank, o.count,
                      o.peer, m.count);
        }
        FDFD_CHECK(cudaMemcpyAsync(o.rbuf, m.ptr, sizeof(double) * o.count, cudaMemcpyDefault, o.st));
        FDFD_CHECK(cudaStreamSynchronize(o.st));
        {
            std::lock_guard<std::mutex> lk(h->mu);
            ++h->consumed[slot];
        }
        h->cv.notify_all();
    }
    // 3. my send buffers are free again once every post has been consumed
    for (auto& t : mine) {
        std::unique_lock<std::mutex> lk(h->mu);
        if (!h->cv.wait_until(lk, deadline, [&] { return h->aborted || h->consumed[t.first] >= t.second; }))
            FDFD_FAIL("in-process communicator: rank %d timed out waiting for a send to complete", c->rank);
        if (h->aborted) FDFD_FAIL("in-process communicator aborted (another rank failed)");
    }
    return 0;
}
static int local_post(FdfdComm* c, const LocalOp& op) {
    if (op.peer < 0 || op.peer >= c->world || op.peer == c->rank) FDFD_FAIL("in-process communicator: bad peer %d", op.peer);
    auto* q = static_cast<std::vector<LocalOp>*>(c->pending);
    q->push_back(op);
    if (c->group_depth > 0) return 0;
    std::vector<LocalOp> ops;
    ops.swap(*q);
    return local_run(c, ops);
}
template <class T, class F>
static int local_allreduce(FdfdComm* c, T* buf, size_t count, cudaStream_t st, F combine) {
    // everybody sends to everybody; summed on the host in rank order (test transport: clarity over speed)
    const int w = c->world;
    if (w == 1) return 0;
    const size_t bytes = sizeof(T) * count, dbl = (bytes + 7) / 8;
    char* tmp = nullptr;
    FDFD_CHECK(cudaMalloc(&tmp, dbl * 8 * w));
    FDFD_CHECK(cudaMemcpyAsync(tmp + dbl * 8 * c->rank, buf, bytes, cudaMemcpyDeviceToDevice, st));
    std::vector<LocalOp> ops;
    for (int p = 0; p < w; ++p)
        if (p != c->rank) {
            ops.push_back({0, tmp + dbl * 8 * c->rank, nullptr, dbl, p, st});
            ops.push_back({1, nullptr, tmp + dbl * 8 * p, dbl, p, st});
        }
    int rc = local_run(c, ops);
    if (!rc) {
        std::vector<T> host((dbl * 8 / sizeof(T)) * w), acc(count);
        if (cudaMemcpy(host.data(), tmp, dbl * 8 * w, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
        const size_t stride = dbl * 8 / sizeof(T);
        for (size_t i = 0; i < count; ++i) {
            T v = host[i];
            for (int p = 1; p < w; ++p) v = combine(v, host[p * stride + i]);
            acc[i] = v;
        }
        if (!rc && cudaMemcpy(buf, acc.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = -1;
    }
    cudaFree(tmp);
    if (rc && !g_fdfd_err[0]) snprintf(g_fdfd_err, sizeof(g_fdfd_err), "in-process all-reduce failed");
    return rc;
}

int comm_create_local(FdfdComm** out, int world) {
    if (world < 1 || world > 64) FDFD_FAIL("in-process communicator: world size 1..64");
    LocalHub* h = new LocalHub();
    h->world = world; h->refs = world; h->aborted = false;
    h->box.resize((size_t)world * world);
    h->posted.assign((size_t)world * world, 0);
    h->consumed.assign((size_t)world * world, 0);
    for (int r = 0; r < world; ++r) {
        FdfdComm* c = new FdfdComm();
        c->nccl = nullptr; c->rank = r; c->world = world; c->hub = h; c->group_depth = 0;
        c->pending = new std::vector<LocalOp>();
        out[r] = c;
    }
    return 0;
}
void comm_abort(FdfdComm* c) {
    if (!c || !c->hub) return;
    {
        std::lock_guard<std::mutex> lk(c->hub->mu);
        c->hub->aborted = true;
    }
    c->hub->cv.notify_all();
}

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
    c->hub = nullptr; c->group_depth = 0; c->pending = nullptr;
    *out = c;
    return 0;
}

void comm_destroy(FdfdComm* c) {
    if (!c) return;
    if (c->hub) {
        bool last;
        {
            std::lock_guard<std::mutex> lk(c->hub->mu);
            last = --c->hub->refs == 0;
        }
        if (last) delete c->hub;
        delete static_cast<std::vector<LocalOp>*>(c->pending);
    } else if (c->nccl && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)c->nccl);
    }
    delete c;
}

int comm_group_begin(FdfdComm* c) {
    if (c->group_depth++ == 0 && !c->hub) NCCL_CHECK(g_nccl.GroupStart());
    return 0;
}
int comm_group_end(FdfdComm* c) {
    if (c->group_depth <= 0) FDFD_FAIL("comm_group_end without comm_group_begin");
    if (--c->group_depth > 0) return 0;
    if (c->hub) {
        std::vector<LocalOp> ops;
        ops.swap(*static_cast<std::vector<LocalOp>*>(c->pending));
        return local_run(c, ops);
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    return 0;
}

int comm_send(FdfdComm* c, const void* buf, size_t count, int peer, cudaStream_t st) {
    if (c->hub) return local_post(c, {0, buf, nullptr, count, peer, st});
    NCCL_CHECK(g_nccl.Send(buf, count, ncclDouble, peer, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_recv(FdfdComm* c, void* buf, size_t count, int peer, cudaStream_t st) {
    if (c->hub) return local_post(c, {1, nullptr, buf, count, peer, st});
    NCCL_CHECK(g_nccl.Recv(buf, count, ncclDouble, peer, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_sendrecv(FdfdComm* c, const void* sbuf, int send_peer, void* rbuf, int recv_peer, size_t count,
                  cudaStream_t st) {
    if (c->hub) {
        if (comm_group_begin(c) || comm_send(c, sbuf, count, send_peer, st) || comm_recv(c, rbuf, count, recv_peer, st)) return -1;
        return comm_group_end(c);
    }
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
    if (c->hub) {
        if (comm_group_begin(c) || comm_send(c, first, count, lower, st) || comm_send(c, last, count, upper, st) ||
            comm_recv(c, halo_hi, count, upper, st) || comm_recv(c, halo_lo, count, lower, st))
            return -1;
        return comm_group_end(c);
    }
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
    if (c->hub) return local_allreduce<double>(c, static_cast<double*>(buf), count, st, [](double a, double b) { return a + b; });
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c->nccl, st));
    return 0;
}
int comm_allreduce_max_i32(FdfdComm* c, int* buf, size_t count, cudaStream_t st) {
    if (c->hub) return local_allreduce<int>(c, buf, count, st, [](int a, int b) { return a > b ? a : b; });
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclInt32, ncclMax, (ncclComm_t)c->nccl, st));
    return 0;
}
