// Krylov drivers on top of the matrix-free stencil: BiCGSTAB (optionally right-preconditioned by
// the cached direct factorisation), COCG on the symmetrised operator, and iterative refinement.
//
// * All iteration scalars live on the device; the host only enqueues kernels and looks at the residual
//   norm every `check_every` iterations.
// * Vector updates and the inner products that follow them are FUSED: x/r update + (r0.r, r.r) in one
//   pass, (t.s, t.t) in one pass, scalar recurrences folded into the last reduction stage.  Reductions are
//   warp shuffles -> one partial per CTA -> one finishing CTA (fixed order: deterministic).
// * Vectors are stored as complex128 or complex64 (template parameter V); arithmetic, inner products and
//   the iteration scalars are fp64 in both cases.
// * Slab operators (one grid over several GPUs): the partial sums are all-reduced on the device between the
//   two reduction stages, halo rows are exchanged inside the stencil application.
#include <algorithm>
#include <type_traits>
#include <vector>
#include "krylov.cuh"

#define RED_BLOCKS 1184   // 148 SMs x 8
#define RED_THREADS 256

// scalar slots
enum { S_RHO = 0, S_RHO_OLD, S_ALPHA, S_OMEGA, S_BETA, S_R0V, S_TS, S_TT, S_RR, S_PQ, S_TMP0, S_TMP1, S_COUNT };
// what the finishing stage of a reduction does with its sums (a, b)
enum { POST_STORE = 0,      // sc[o0] = a (; sc[o1] = b)
       POST_BICG_ALPHA,     // R0V = a; ALPHA = RHO / R0V
       POST_BICG_OMEGA,     // TS = a, TT = b; OMEGA = TS / TT
       POST_BICG_RHO,       // RHO_OLD = RHO; RHO = a; RR = b; BETA = (RHO / RHO_OLD) (ALPHA / OMEGA)
       POST_COCG_ALPHA,     // PQ = a; ALPHA = RHO / PQ
       POST_COCG_BETA };    // BETA = a / RHO; RHO = a; RR = b

__device__ __forceinline__ void block_reduce2(cplx& a, cplx& b) {
    __shared__ double sh[4][RED_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_down_sync(0xffffffffu, a.x, o);
        a.y += __shfl_down_sync(0xffffffffu, a.y, o);
        b.x += __shfl_down_sync(0xffffffffu, b.x, o);
        b.y += __shfl_down_sync(0xffffffffu, b.y, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a.x; sh[1][w] = a.y; sh[2][w] = b.x; sh[3][w] = b.y; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        for (int i = 0; i < RED_THREADS / 32; ++i) { t0 += sh[0][i]; t1 += sh[1][i]; t2 += sh[2][i]; t3 += sh[3][i]; }
        a = make_double2(t0, t1);
        b = make_double2(t2, t3);
    }
}

__device__ __forceinline__ void dot_acc(cplx& acc, cplx u, cplx v, bool conj_u) {
    if (conj_u) u.y = -u.y;
    acc.x += u.x * v.x - u.y * v.y;
    acc.y += u.x * v.y + u.y * v.x;
}

// partial[2 * blk] = sum op(a1) b1,  partial[2 * blk + 1] = sum op(a2) b2  (second pair optional)
template <class V>
__global__ void __launch_bounds__(RED_THREADS)
dot2_partial_kernel(const V* __restrict__ a1, const V* __restrict__ b1, const V* __restrict__ a2,
                    const V* __restrict__ b2, size_t n, int conj1, int conj2, cplx* __restrict__ partial) {
    cplx s1 = make_double2(0.0, 0.0), s2 = make_double2(0.0, 0.0);
#pragma unroll 4
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx u = vload(a1 + i), v = (b1 == a1) ? u : vload(b1 + i);
        dot_acc(s1, u, v, conj1);
        if (a2) {
            cplx p = (a2 == a1) ? u : ((a2 == b1) ? v : vload(a2 + i));
            cplx q = (b2 == a1) ? u : ((b2 == b1) ? v : vload(b2 + i));
            dot_acc(s2, p, q, conj2);
        }
    }
    block_reduce2(s1, s2);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s1; partial[2 * blockIdx.x + 1] = s2; }
}

__device__ __forceinline__ void post_op(cplx* sc, int post, int o0, int o1, cplx a, cplx b) {
    switch (post) {
    case POST_STORE: sc[o0] = a; if (o1 >= 0) sc[o1] = b; break;
    case POST_BICG_ALPHA: sc[S_R0V] = a; sc[S_ALPHA] = cdiv(sc[S_RHO], a); break;
    case POST_BICG_OMEGA: sc[S_TS] = a; sc[S_TT] = b; sc[S_OMEGA] = cdiv(a, b); break;
    case POST_BICG_RHO: {
        cplx rho_old = sc[S_RHO];
        sc[S_RHO_OLD] = rho_old; sc[S_RHO] = a; sc[S_RR] = b;
        sc[S_BETA] = cmul(cdiv(a, rho_old), cdiv(sc[S_ALPHA], sc[S_OMEGA]));
        break;
    }
    case POST_COCG_ALPHA: sc[S_PQ] = a; sc[S_ALPHA] = cdiv(sc[S_RHO], a); break;
    case POST_COCG_BETA: sc[S_BETA] = cdiv(a, sc[S_RHO]); sc[S_RHO] = a; sc[S_RR] = b; break;
    }
}

// finishing stage: sums the per-CTA partials; post >= 0 applies the scalar recurrence right away,
// post < 0 leaves the two sums in sc[S_TMP0..1] (they are all-reduced over the ranks first)
__global__ void __launch_bounds__(RED_THREADS)
dot2_final_kernel(const cplx* __restrict__ partial, int count, cplx* __restrict__ sc, int post, int o0, int o1,
                  int real1, int real2) {
    cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < count; i += blockDim.x) { a = cadd(a, partial[2 * i]); b = cadd(b, partial[2 * i + 1]); }
    block_reduce2(a, b);
    if (threadIdx.x == 0) {
        if (real1) a.y = 0.0;
        if (real2) b.y = 0.0;
        if (post < 0) { sc[S_TMP0] = a; sc[S_TMP1] = b; }
        else post_op(sc, post, o0, o1, a, b);
    }
}
__global__ void post_kernel(cplx* sc, int post, int o0, int o1) { post_op(sc, post, o0, o1, sc[S_TMP0], sc[S_TMP1]); }

// second stage (+ all-reduce over the ranks of a slab operator) of a reduction whose partials are in place
static int finish_dots(cudaStream_t st, cplx* partial, cplx* sc, int post, int o0, int o1, int real1, int real2,
                       FdfdComm* comm) {
    const bool dist = comm && comm->world > 1;
    dot2_final_kernel<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, sc, dist ? -1 : post, o0, o1, real1, real2);
    ++g_fdfd_launches;
    if (dist) {
        // the scalars are summed over the ranks on the same stream; they never visit the host
        if (comm_allreduce_sum(comm, sc + S_TMP0, 4, st)) return -1;
        post_kernel<<<1, 1, 0, st>>>(sc, post, o0, o1);
        ++g_fdfd_launches;
    }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}

template <class V>
static int dots(cudaStream_t st, const V* a1, const V* b1, int conj1, const V* a2, const V* b2, int conj2, size_t n,
                cplx* partial, cplx* sc, int post, int o0, int o1, int real_only, FdfdComm* comm) {
    dot2_partial_kernel<V><<<RED_BLOCKS, RED_THREADS, 0, st>>>(a1, b1, a2, b2, n, conj1, conj2, partial);
    ++g_fdfd_launches;
    return finish_dots(st, partial, sc, post, o0, o1, real_only, real_only, comm);
}

// one inner product into sc[0]; `partial` holds 2 * RED_BLOCKS entries, `sc` S_COUNT entries
int dev_dot(cudaStream_t st, const cplx* a, const cplx* b, size_t n, bool conj_a, cplx* partial, cplx* sc,
            int real_only, FdfdComm* comm) {
    return dots<cplx>(st, a, b, conj_a, nullptr, nullptr, 0, n, partial, sc, POST_STORE, 0, -1, real_only, comm);
}

// p = r + beta (p - omega v)
template <class V>
__global__ void bicg_p_kernel(V* __restrict__ p, const V* __restrict__ r, const V* __restrict__ v,
                              const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx beta = sc[S_BETA], om = sc[S_OMEGA];
    cplx t = csub(vload(p + i), cmul(om, vload(v + i)));
    vstore(p + i, cadd(vload(r + i), cmul(beta, t)));
}
// s = r - alpha v
template <class V>
__global__ void bicg_s_kernel(V* __restrict__ s, const V* __restrict__ r, const V* __restrict__ v,
                              const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vstore(s + i, csub(vload(r + i), cmul(sc[S_ALPHA], vload(v + i))));
}
__device__ __forceinline__ cplx as_stored(cplx v, const cplx*) { return v; }
__device__ __forceinline__ cplx as_stored(cplx v, const cplx32*) {
    return make_double2((double)(float)v.x, (double)(float)v.y);
}
// x += alpha ph + omega sh ; r = s - omega t ; partial sums of conj(r0).r and conj(r).r of the NEW r
template <class V>
__global__ void __launch_bounds__(RED_THREADS)
bicg_xr_dots_kernel(V* __restrict__ x, V* __restrict__ r, const V* __restrict__ ph, const V* __restrict__ sh,
                    const V* __restrict__ s, const V* __restrict__ t, const V* __restrict__ r0,
                    const cplx* __restrict__ sc, size_t n, cplx* __restrict__ partial) {
    const cplx al = sc[S_ALPHA], om = sc[S_OMEGA];
    cplx d1 = make_double2(0.0, 0.0), d2 = make_double2(0.0, 0.0);
#pragma unroll 2
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx sv = vload(s + i);
        cplx xv = vload(x + i);
        cfma(xv, al, vload(ph + i));
        cfma(xv, om, (sh == s) ? sv : vload(sh + i));
        vstore(x + i, xv);
        cplx rn = as_stored(csub(sv, cmul(om, vload(t + i))), r);   // inner products of the value as stored
        vstore(r + i, rn);
        dot_acc(d1, vload(r0 + i), rn, true);
        dot_acc(d2, rn, rn, true);
    }
    block_reduce2(d1, d2);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = d1; partial[2 * blockIdx.x + 1] = d2; }
}
// x += alpha p ; r -= alpha q ; partial sums of r.r (unconjugated: COCG) and conj(r).r (norm) of the NEW r
template <class V>
__global__ void __launch_bounds__(RED_THREADS)
cocg_xr_dots_kernel(V* __restrict__ x, V* __restrict__ r, const V* __restrict__ p, const V* __restrict__ q,
                    const cplx* __restrict__ sc, size_t n, cplx* __restrict__ partial) {
    const cplx al = sc[S_ALPHA];
    cplx d1 = make_double2(0.0, 0.0), d2 = make_double2(0.0, 0.0);
#pragma unroll 2
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx xv = vload(x + i);
        cfma(xv, al, vload(p + i));
        vstore(x + i, xv);
        cplx rn = as_stored(csub(vload(r + i), cmul(al, vload(q + i))), r);
        vstore(r + i, rn);
        dot_acc(d1, rn, rn, false);
        dot_acc(d2, rn, rn, true);
    }
    block_reduce2(d1, d2);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = d1; partial[2 * blockIdx.x + 1] = d2; }
}
// p = r + beta p
template <class V>
__global__ void cocg_p_kernel(V* __restrict__ p, const V* __restrict__ r, const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vstore(p + i, cadd(vload(r + i), cmul(sc[S_BETA], vload(p + i))));
}
// v[i] *= sxf[ix] * syf[iy]  = v / (isxf isyf): left scaling that makes A complex symmetric
template <class V>
__global__ void sym_scale_kernel(V* __restrict__ v, const cplx* __restrict__ isxf, const cplx* __restrict__ isyf,
                                 int nx, int ny) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nx * ny) return;
    int ix = (int)(i / ny), iy = (int)(i % ny);
    vstore(v + i, cdiv(vload(v + i), cmul(isxf[ix], isyf[iy])));
}
__global__ void axpy_one_kernel(cplx* __restrict__ x, const cplx* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = cadd(x[i], d[i]);
}

struct Scratch {
    void* base = nullptr;
    ~Scratch() { if (base) cudaFree(base); }
};

static int host_scalar(cudaStream_t st, const cplx* d, cplx* h) {
    FDFD_CHECK(cudaMemcpyAsync(h, d, sizeof(cplx), cudaMemcpyDeviceToHost, st));
    FDFD_CHECK(cudaStreamSynchronize(st));
    return 0;
}

// y (+)= c12 .* conj(x): the anti-linear part of the Newton Jacobian (nonlinear_solvers.py:134-135, Jac12)
__global__ void conj_couple_kernel(cplx* __restrict__ y, const cplx* __restrict__ c12, const cplx* __restrict__ x,
                                   size_t n, int subtract) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx t = cmul(c12[i], cconj(x[i]));
    y[i] = subtract ? csub(y[i], t) : cadd(y[i], t);
}

// Slab operators (operator.cuh): vectors carry one halo row before and after the rows this rank owns.
// The solvers below work on the owned rows (pointer = extended base + pad, n = owned cells) and hand the
// extended base to the stencil, which fills the halos from the neighbouring ranks.
static inline size_t kpad(const FdfdOp* op) { return op->halo ? (size_t)op->ny : 0; }
static inline size_t kn(const FdfdOp* op) { return (size_t)(op->nx - 2 * op->halo) * op->ny; }

static int couple(const FdfdOp* op, cplx* y, const cplx* c12, const cplx* x, int subtract) {
    { conj_couple_kernel<<<ceil_div(op->n(), 256), 256, 0, op->stream>>>(y, c12, x, op->n(), subtract); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}
static int couple(const FdfdOp*, cplx32*, const cplx*, const cplx32*, int) { return 0; }   // rejected at entry

template <class V>
static int apply_A(const FdfdOp* op, const V* x, V* y, int fused, const cplx* c12) {
    const size_t pad = kpad(op);
    if (fused ? op_apply_fused_t<V>(op, x - pad, y - pad, 1) : op_apply_planes_t<V>(op, x - pad, y - pad, 1)) return -1;
    return c12 ? couple(op, y, c12, x, 0) : 0;
}

template <class V>
static int residual_A(const FdfdOp* op, const V* b, const V* x, V* r, const cplx* c12) {
    const size_t pad = kpad(op);
    if (op_residual_t<V>(op, b - pad, x - pad, r - pad, 1)) return -1;
    return c12 ? couple(op, r, c12, x, 1) : 0;
}

// ------------------------------------------------------------------------------------------
// Restricted additive Schwarz preconditioner of the slab path (one subdomain per rank).
// Subdomain = the rank's rows + `ov` overlap rows + `npml_s` artificial PML rows on both sides, a local torus
// factorised by the direct solver (op_create_schwarz_sub).  z = M^-1 r:
//   1. r on the owned rows is copied into the subdomain right-hand side, the overlap rows arrive from the two
//      neighbouring ranks (one grouped send/recv of ov rows each way), the PML rows stay zero;
//   2. one substitution pass with the local factors;
//   3. only the owned rows of the local solution are kept (the "restricted" part: no second exchange, no double
//      counting in the overlap).
// Information crosses one subdomain boundary per application, so the iteration count of the preconditioned
// BiCGSTAB grows with the number of slabs (measured: DESIGN.md section 5).
// ------------------------------------------------------------------------------------------
struct SchwarzPre {
    FdfdOp* sub;          // not owned (the host keeps the handle it assembled and factorised)
    NdSolver* nd;         // not owned
    int ov, npml_s, ext, nxl;
    cplx *rin, *zout;     // subdomain right-hand side / solution, (nxl + 2 ext) x ny
    cudaEvent_t ev_in, ev_out;
};

int schwarz_attach(FdfdOp* slab, FdfdOp* sub, NdSolver* nd, int overlap, int npml_sub) {
    if (!slab->halo) FDFD_FAIL("the Schwarz preconditioner belongs to a slab operator");
    if (slab->schwarz) { schwarz_destroy(slab->schwarz); slab->schwarz = nullptr; }
    if (!sub) return 0;                                   // detach
    const int nxl = slab->nx - 2, ext = overlap + npml_sub;
    if (!nd || sub->nx != nxl + 2 * ext || sub->ny != slab->ny || nd->nx != sub->nx || nd->ny != sub->ny)
        FDFD_FAIL("Schwarz subdomain / factor shape does not match the slab (%d + 2 x %d rows)", nxl, ext);
    if (overlap > nxl) FDFD_FAIL("Schwarz overlap exceeds the slab");
    SchwarzPre* s = new SchwarzPre();
    s->sub = sub; s->nd = nd; s->ov = overlap; s->npml_s = npml_sub; s->ext = ext; s->nxl = nxl;
    s->rin = s->zout = nullptr; s->ev_in = s->ev_out = nullptr;
    slab->schwarz = s;
    const size_t ne = sub->n();
    FDFD_CHECK(cudaMalloc(&s->rin, sizeof(cplx) * ne));
    FDFD_CHECK(cudaMalloc(&s->zout, sizeof(cplx) * ne));
    FDFD_CHECK(cudaMemset(s->rin, 0, sizeof(cplx) * ne));          // the PML rows of the right-hand side are never written
    FDFD_CHECK(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
    FDFD_CHECK(cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
    return 0;
}
void schwarz_destroy(SchwarzPre* s) {
    if (!s) return;
    cudaFree(s->rin); cudaFree(s->zout);
    if (s->ev_in) cudaEventDestroy(s->ev_in);
    if (s->ev_out) cudaEventDestroy(s->ev_out);
    delete s;
}
static int schwarz_apply(const FdfdOp* op, const cplx* in, cplx* out) {
    SchwarzPre* s = op->schwarz;
    cudaStream_t st = op->stream;
    const size_t ny = (size_t)op->ny, n = (size_t)s->nxl * ny;
    cplx* owned = s->rin + (size_t)s->ext * ny;
    FDFD_CHECK(cudaMemcpyAsync(owned, in, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
    if (s->ov > 0) {
        const size_t cnt = (size_t)s->ov * ny;
        cplx *first = owned, *last = owned + (size_t)(s->nxl - s->ov) * ny;
        cplx *lo = s->rin + (size_t)s->npml_s * ny, *hi = owned + n;
        if (!op->comm || op->comm->world == 1) {          // one slab: its neighbours are itself
            FDFD_CHECK(cudaMemcpyAsync(lo, last, sizeof(cplx) * cnt, cudaMemcpyDeviceToDevice, st));
            FDFD_CHECK(cudaMemcpyAsync(hi, first, sizeof(cplx) * cnt, cudaMemcpyDeviceToDevice, st));
        } else {
            const int w = op->comm->world, r = op->comm->rank;
            if (comm_halo_exchange(op->comm, first, last, lo, hi, (r + w - 1) % w, (r + 1) % w, 2 * cnt, st)) return -1;
        }
    }
    // the substitution runs on the subdomain operator's stream
    FDFD_CHECK(cudaEventRecord(s->ev_in, st));
    FDFD_CHECK(cudaStreamWaitEvent(s->sub->stream, s->ev_in, 0));
    if (nd_solve(s->nd, s->sub, s->rin, s->zout, 1)) return -1;
    FDFD_CHECK(cudaEventRecord(s->ev_out, s->sub->stream));
    FDFD_CHECK(cudaStreamWaitEvent(st, s->ev_out, 0));
    FDFD_CHECK(cudaMemcpyAsync(out, s->zout + (size_t)s->ext * ny, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int precond_solve(NdSolver* nd, const FdfdOp* op, const cplx* in, cplx* out) {
    return nd ? nd_solve(nd, op, in, out, 1) : schwarz_apply(op, in, out);
}
static int precond_solve(NdSolver*, const FdfdOp*, const cplx32*, cplx32*) { return 0; }   // rejected at entry

template <class V>
int krylov_bicgstab_t(const FdfdOp* op, NdSolver* precond_nd, const V* d_b, V* d_x, double tol, int maxiter, int fused,
                      int check_every, const cplx* c12, int real_inner, KrylovResult* res) {
    const int RI = (real_inner || c12) ? 1 : 0;   // an R-linear operator needs the real inner product
    const size_t n = kn(op), pad = kpad(op), vs = n + 2 * pad;
    if (op->halo && (precond_nd || c12))
        FDFD_FAIL("slab operators take no whole-grid factors and no anti-linear term (their preconditioner is the attached Schwarz one)");
    if (!std::is_same<V, cplx>::value && (precond_nd || c12))
        FDFD_FAIL("the complex64 solver takes neither a preconditioner nor an anti-linear term");
    // right preconditioner: whole-grid factors handed in, or the Schwarz preconditioner attached to a slab operator
    const bool precond = precond_nd != nullptr || (std::is_same<V, cplx>::value && op->halo && op->schwarz != nullptr);
    FdfdComm* comm = op->comm;
    d_b += pad; d_x += pad;
    cudaStream_t st = op->stream;
    const int nvec = (precond ? 8 : 6) + 1;          // + the last checked iterate (breakdown restarts)
    int restarts = 0;
    Scratch ws, wsc;
    FDFD_CHECK(cudaMalloc(&ws.base, sizeof(V) * vs * nvec));
    FDFD_CHECK(cudaMalloc(&wsc.base, sizeof(cplx) * (2 * RED_BLOCKS + S_COUNT)));
    if (pad) FDFD_CHECK(cudaMemsetAsync(ws.base, 0, sizeof(V) * vs * nvec, st));
    V *r = static_cast<V*>(ws.base) + pad, *r0 = r + vs, *p = r0 + vs, *v = p + vs, *s = v + vs, *t = s + vs;
    V *ph = precond ? t + vs : p, *sh = precond ? ph + vs : s;
    V* xbak = (precond ? sh : t) + vs;
    cplx* partial = static_cast<cplx*>(wsc.base);
    cplx* sc = partial + 2 * RED_BLOCKS;
    const int nblk = ceil_div(n, 256);
    cplx h;
    if (check_every < 1) check_every = 1;
    // r = b - A x
    FDFD_CHECK(cudaMemcpyAsync(xbak, d_x, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    if (residual_A<V>(op, d_b, d_x, r, c12)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(r0, r, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    FDFD_CHECK(cudaMemsetAsync(p, 0, sizeof(V) * n, st));
    FDFD_CHECK(cudaMemsetAsync(v, 0, sizeof(V) * n, st));
    cplx init[S_COUNT];
    for (int i = 0; i < S_COUNT; ++i) init[i] = make_double2(1.0, 0.0);
    FDFD_CHECK(cudaMemcpyAsync(sc, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (dots<V>(st, d_b, d_b, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    const double bnorm = sqrt(h.x);
    res->iters = 0; res->converged = 0; res->relres = 1.0;
    if (bnorm == 0.0) {
        FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(V) * n, st));
        res->converged = 1; res->relres = 0.0;
        return 0;
    }
    // rho = r0.r and ||r||^2 of the starting residual; beta of the first step is (rho / 1) (1 / 1)
    if (dots<V>(st, r0, r, 1, r, r, 1, n, partial, sc, POST_BICG_RHO, 0, 0, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    res->relres = sqrt(h.x) / bnorm;
    if (res->relres <= tol) { res->converged = 1; return 0; }
    for (int it = 1; it <= maxiter; ++it) {
        { bicg_p_kernel<V><<<nblk, 256, 0, st>>>(p, r, v, sc, n); ++g_fdfd_launches; }
        if (precond && precond_solve(precond_nd, op, p, ph)) return -1;
        if (apply_A<V>(op, ph, v, fused, c12)) return -1;
        if (dots<V>(st, r0, v, 1, nullptr, nullptr, 0, n, partial, sc, POST_BICG_ALPHA, 0, 0, RI, comm)) return -1;
        { bicg_s_kernel<V><<<nblk, 256, 0, st>>>(s, r, v, sc, n); ++g_fdfd_launches; }
        if (precond && precond_solve(precond_nd, op, s, sh)) return -1;
        if (apply_A<V>(op, sh, t, fused, c12)) return -1;
        if (dots<V>(st, t, s, 1, t, t, 1, n, partial, sc, POST_BICG_OMEGA, 0, 0, RI, comm)) return -1;
        { bicg_xr_dots_kernel<V><<<RED_BLOCKS, RED_THREADS, 0, st>>>(d_x, r, ph, sh, s, t, r0, sc, n, partial); ++g_fdfd_launches; }
        if (finish_dots(st, partial, sc, POST_BICG_RHO, 0, 0, RI, 1, comm)) return -1;
        FDFD_CHECK(cudaGetLastError());
        res->iters = it;
        if (it % check_every == 0 || it == maxiter) {
            if (host_scalar(st, sc + S_RR, &h)) return -1;       // ||r||^2 came with the fused update
            res->relres = sqrt(h.x) / bnorm;
            if (!(res->relres == res->relres) || res->relres > 1e150) {
                // breakdown (rho or omega ~ 0 poisoned the recurrences): go back to the last checked iterate and
                // restart the recurrences from its true residual; a few restarts are allowed
                if (restarts++ >= 5) break;
                FDFD_CHECK(cudaMemcpyAsync(d_x, xbak, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
                if (residual_A<V>(op, d_b, d_x, r, c12)) return -1;
                FDFD_CHECK(cudaMemcpyAsync(r0, r, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
                FDFD_CHECK(cudaMemsetAsync(p, 0, sizeof(V) * n, st));
                FDFD_CHECK(cudaMemsetAsync(v, 0, sizeof(V) * n, st));
                FDFD_CHECK(cudaMemcpyAsync(sc, init, sizeof(init), cudaMemcpyHostToDevice, st));
                if (dots<V>(st, r0, r, 1, r, r, 1, n, partial, sc, POST_BICG_RHO, 0, 0, RI, comm)) return -1;
                continue;
            }
            if (res->relres <= tol) { res->converged = 1; break; }
            FDFD_CHECK(cudaMemcpyAsync(xbak, d_x, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (!(res->relres == res->relres) || res->relres > 1e150)
        FDFD_CHECK(cudaMemcpyAsync(d_x, xbak, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));   // never hand back NaNs
    // report the TRUE residual of the returned iterate
    if (residual_A<V>(op, d_b, d_x, r, c12)) return -1;
    if (dots<V>(st, r, r, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    res->relres = sqrt(h.x) / bnorm;
    res->converged = res->relres <= tol * 10 ? res->converged : 0;
    return 0;
}

template <class V>
int krylov_cocg_t(const FdfdOp* op, const V* d_b, V* d_x, double tol, int maxiter, int fused, int check_every,
                  KrylovResult* res) {
    const size_t n = kn(op), pad = kpad(op), vs = n + 2 * pad;
    FdfdComm* comm = op->comm;
    d_b += pad; d_x += pad;
    const cplx *sxf = op->isxf + op->halo, *syf = op->isyf;
    const int nxo = op->nx - 2 * op->halo;
    cudaStream_t st = op->stream;
    Scratch ws, wsc;
    int restarts = 0;
    FDFD_CHECK(cudaMalloc(&ws.base, sizeof(V) * vs * 5));
    FDFD_CHECK(cudaMalloc(&wsc.base, sizeof(cplx) * (2 * RED_BLOCKS + S_COUNT)));
    if (pad) FDFD_CHECK(cudaMemsetAsync(ws.base, 0, sizeof(V) * vs * 5, st));
    V *r = static_cast<V*>(ws.base) + pad, *p = r + vs, *q = p + vs, *bs = q + vs, *xbak = bs + vs;
    cplx* partial = static_cast<cplx*>(wsc.base);
    cplx* sc = partial + 2 * RED_BLOCKS;
    const int nblk = ceil_div(n, 256);
    cplx h;
    if (check_every < 1) check_every = 1;
    // symmetrised system  D A x = D b,  D = diag(sxf[ix] syf[iy])
    FDFD_CHECK(cudaMemcpyAsync(bs, d_b, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    { sym_scale_kernel<V><<<nblk, 256, 0, st>>>(bs, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
    if (residual_A<V>(op, d_b, d_x, r, nullptr)) return -1;
    { sym_scale_kernel<V><<<nblk, 256, 0, st>>>(r, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
    FDFD_CHECK(cudaMemcpyAsync(p, r, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    FDFD_CHECK(cudaMemcpyAsync(xbak, d_x, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    if (dots<V>(st, bs, bs, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    const double bnorm = sqrt(h.x);
    res->iters = 0; res->converged = 0; res->relres = 1.0;
    if (bnorm == 0.0) {
        FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(V) * n, st));
        res->converged = 1; res->relres = 0.0;
        return 0;
    }
    if (dots<V>(st, r, r, 0, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RHO, -1, 0, comm)) return -1;
    for (int it = 1; it <= maxiter; ++it) {
        if (apply_A<V>(op, p, q, fused, nullptr)) return -1;
        { sym_scale_kernel<V><<<nblk, 256, 0, st>>>(q, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
        if (dots<V>(st, p, q, 0, nullptr, nullptr, 0, n, partial, sc, POST_COCG_ALPHA, 0, 0, 0, comm)) return -1;
        { cocg_xr_dots_kernel<V><<<RED_BLOCKS, RED_THREADS, 0, st>>>(d_x, r, p, q, sc, n, partial); ++g_fdfd_launches; }
        if (finish_dots(st, partial, sc, POST_COCG_BETA, 0, 0, 0, 1, comm)) return -1;
        { cocg_p_kernel<V><<<nblk, 256, 0, st>>>(p, r, sc, n); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
        res->iters = it;
        if (it % check_every == 0 || it == maxiter) {
            if (host_scalar(st, sc + S_RR, &h)) return -1;
            res->relres = sqrt(h.x) / bnorm;
            if (!(res->relres == res->relres) || res->relres > 1e150) {
                // COCG breakdown (p^T A p ~ 0 is possible for a complex symmetric indefinite operator): restart from
                // the last checked iterate with p = r
                if (restarts++ >= 5) break;
                FDFD_CHECK(cudaMemcpyAsync(d_x, xbak, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
                if (residual_A<V>(op, d_b, d_x, r, nullptr)) return -1;
                { sym_scale_kernel<V><<<nblk, 256, 0, st>>>(r, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
                FDFD_CHECK(cudaMemcpyAsync(p, r, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
                if (dots<V>(st, r, r, 0, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RHO, -1, 0, comm)) return -1;
                continue;
            }
            if (res->relres <= tol) { res->converged = 1; break; }
            FDFD_CHECK(cudaMemcpyAsync(xbak, d_x, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (!(res->relres == res->relres) || res->relres > 1e150)
        FDFD_CHECK(cudaMemcpyAsync(d_x, xbak, sizeof(V) * n, cudaMemcpyDeviceToDevice, st));
    // true residual in the ORIGINAL (unscaled) system
    if (residual_A<V>(op, d_b, d_x, r, nullptr)) return -1;
    if (dots<V>(st, r, r, 1, d_b, d_b, 1, n, partial, sc, POST_STORE, S_RR, S_TT, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    double rn = sqrt(h.x);
    if (host_scalar(st, sc + S_TT, &h)) return -1;
    res->relres = rn / sqrt(h.x);
    return 0;
}

// ------------------------------------------------------------------------------------------
// Restarted GMRES, right-preconditioned (whole-grid factors of a nearby operator, or the Schwarz preconditioner of a
// slab operator): ONE preconditioner application per iteration against BiCGSTAB's two, a monotone residual and no
// breakdowns -- on the Schwarz-preconditioned slab systems BiCGSTAB needs 2-5x the applications and occasionally
// breaks down (tests/schwarz_model.py).  Arnoldi with classical Gram-Schmidt applied twice (CGS2): the j + 1 inner
// products of a pass come from one kernel (four basis vectors per CTA pass, so w is re-read j / 4 times, not j),
// are summed over the ranks with one all-reduce and cross to the host once (the Hessenberg least-squares problem is
// updated there with Givens rotations: O(j) flops).  Basis vectors keep the slab's extended layout.
// ------------------------------------------------------------------------------------------
#define GM_NV 4
__global__ void __launch_bounds__(RED_THREADS)
gmres_dots_kernel(const cplx* __restrict__ V, size_t vs, const cplx* __restrict__ w, int k, size_t n,
                  cplx* __restrict__ partial, int kpad) {
    // partial[blk][i] = sum over the CTA's elements of conj(V_i) w, i in [4 y, 4 y + 4); slot k (y == 0) = |w|^2
    const int i0 = blockIdx.y * GM_NV;
    cplx acc[GM_NV], ww = make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < GM_NV; ++q) acc[q] = make_double2(0.0, 0.0);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const cplx wv = w[e];
        if (blockIdx.y == 0) ww.x += wv.x * wv.x + wv.y * wv.y;
#pragma unroll
        for (int q = 0; q < GM_NV; ++q)
            if (i0 + q < k) dot_acc(acc[q], __ldg(V + (size_t)(i0 + q) * vs + e), wv, true);
    }
    block_reduce2(acc[0], acc[1]);
    __syncthreads();
    block_reduce2(acc[2], acc[3]);
    __syncthreads();
    cplx dummy = make_double2(0.0, 0.0);
    if (blockIdx.y == 0) block_reduce2(ww, dummy);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < GM_NV; ++q)
            if (i0 + q < k) partial[(size_t)blockIdx.x * kpad + i0 + q] = acc[q];
        if (blockIdx.y == 0) partial[(size_t)blockIdx.x * kpad + k] = ww;
    }
}
__global__ void __launch_bounds__(RED_THREADS)
gmres_dots_final_kernel(const cplx* __restrict__ partial, int nblk, int kpad, cplx* __restrict__ out) {
    cplx a = make_double2(0.0, 0.0), dummy = make_double2(0.0, 0.0);
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) a = cadd(a, partial[(size_t)b * kpad + blockIdx.x]);
    block_reduce2(a, dummy);
    if (threadIdx.x == 0) out[blockIdx.x] = a;
}
// w = beta w - sum_i c[i] V_i   (beta = 1: Gram-Schmidt update; beta = 0 with negated c: a combination of the basis)
__global__ void __launch_bounds__(256)
gmres_axpy_kernel(cplx* __restrict__ w, const cplx* __restrict__ V, size_t vs, const cplx* __restrict__ c, int k,
                  size_t n, double beta) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    cplx acc = beta != 0.0 ? w[e] : make_double2(0.0, 0.0);
    for (int i = 0; i < k; ++i) {
        const cplx h = __ldg(c + i), v = __ldg(V + (size_t)i * vs + e);
        acc.x -= h.x * v.x - h.y * v.y;
        acc.y -= h.x * v.y + h.y * v.x;
    }
    w[e] = acc;
}
__global__ void __launch_bounds__(256) gmres_scale_kernel(cplx* __restrict__ w, size_t n, double s) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) w[e] = make_double2(w[e].x * s, w[e].y * s);
}
__global__ void __launch_bounds__(256) gmres_add_kernel(cplx* __restrict__ x, const cplx* __restrict__ z, size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) x[e] = cadd(x[e], z[e]);
}

int krylov_gmres(const FdfdOp* op, NdSolver* precond_nd, const cplx* d_b, cplx* d_x, double tol, int maxiter, int restart,
                 int fused, KrylovResult* res) {
    if (op->halo && precond_nd) FDFD_FAIL("slab operators take no whole-grid factors (their preconditioner is the attached Schwarz one)");
    const bool precond = precond_nd != nullptr || (op->halo && op->schwarz != nullptr);
    const size_t n = kn(op), pad = kpad(op), vs = n + 2 * pad;
    FdfdComm* comm = op->comm;
    const bool dist = comm && comm->world > 1;
    cudaStream_t st = op->stream;
    if (restart < 2) restart = 50;
    if (restart > maxiter) restart = maxiter > 1 ? maxiter : 2;
    const int m = restart, kpad_ = m + 2;
    constexpr int GM_BLOCKS = 592;
    d_b += pad; d_x += pad;
    Scratch wsV, wsS;
    FDFD_CHECK(cudaMalloc(&wsV.base, sizeof(cplx) * vs * (size_t)(m + 3)));
    FDFD_CHECK(cudaMalloc(&wsS.base, sizeof(cplx) * ((size_t)GM_BLOCKS * kpad_ + 2 * kpad_ + 2 * RED_BLOCKS + S_COUNT)));
    if (pad) FDFD_CHECK(cudaMemsetAsync(wsV.base, 0, sizeof(cplx) * vs * (size_t)(m + 3), st));
    cplx* V = static_cast<cplx*>(wsV.base) + pad;           // basis vector i at V + i vs
    cplx* z = V + (size_t)(m + 1) * vs;                      // M^-1 v_j
    cplx* u = z + vs;                                        // residual / basis combination
    cplx* gpart = static_cast<cplx*>(wsS.base);
    cplx* dh = gpart + (size_t)GM_BLOCKS * kpad_;            // device copy of one pass's inner products
    cplx* dc = dh + kpad_;                                   // device copy of the coefficients of an update
    cplx* partial = dc + kpad_;
    cplx* sc = partial + 2 * RED_BLOCKS;
    const unsigned nblk = (unsigned)ceil_div(n, 256);
    std::vector<cplx> hh(kpad_), h1(kpad_);
    std::vector<cplx> H((size_t)(m + 1) * m), cs(m), g(m + 1), y(m);
    std::vector<double> cc(m);
    cplx hs;
    if (dots<cplx>(st, d_b, d_b, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &hs)) return -1;
    const double bnorm = sqrt(hs.x);
    res->iters = 0; res->converged = 0; res->relres = 1.0;
    if (bnorm == 0.0) {
        FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(cplx) * n, st));
        res->converged = 1; res->relres = 0.0;
        return 0;
    }
    // one Gram-Schmidt pass: hh[0..k) = V^H w, hh[k] = |w|^2 (before the update), then w -= V hh
    auto gs_pass = [&](cplx* w, int k) -> int {
        dim3 grid(GM_BLOCKS, (unsigned)ceil_div(k, GM_NV));
        gmres_dots_kernel<<<grid, RED_THREADS, 0, st>>>(V, vs, w, k, n, gpart, kpad_);
        gmres_dots_final_kernel<<<k + 1, RED_THREADS, 0, st>>>(gpart, GM_BLOCKS, kpad_, dh);
        g_fdfd_launches += 2;
        if (dist && comm_allreduce_sum(comm, dh, 2 * (size_t)(k + 1), st)) return -1;
        FDFD_CHECK(cudaMemcpyAsync(hh.data(), dh, sizeof(cplx) * (k + 1), cudaMemcpyDeviceToHost, st));
        gmres_axpy_kernel<<<nblk, 256, 0, st>>>(w, V, vs, dh, k, n, 1.0);
        ++g_fdfd_launches;
        FDFD_CHECK(cudaStreamSynchronize(st));
        FDFD_CHECK(cudaGetLastError());
        return 0;
    };
    int total = 0;
    while (total < maxiter) {
        // r = b - A x ; v_0 = r / |r|
        if (residual_A<cplx>(op, d_b, d_x, V, nullptr)) return -1;
        if (dots<cplx>(st, V, V, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, 0, comm)) return -1;
        if (host_scalar(st, sc + S_RR, &hs)) return -1;
        const double beta = sqrt(hs.x);
        res->relres = beta / bnorm;
        if (!(res->relres == res->relres)) FDFD_FAIL("GMRES: the residual is not finite");
        if (res->relres <= tol) { res->converged = 1; break; }
        gmres_scale_kernel<<<nblk, 256, 0, st>>>(V, n, 1.0 / beta);
        ++g_fdfd_launches;
        for (auto& e : g) e = make_double2(0.0, 0.0);
        g[0] = make_double2(beta, 0.0);
        int j = 0;
        bool done = false;
        for (; j < m && total < maxiter && !done; ++j, ++total) {
            cplx* vj = V + (size_t)j * vs;
            cplx* w = V + (size_t)(j + 1) * vs;
            if (precond) {
                if (precond_solve(precond_nd, op, vj, z)) return -1;
                if (apply_A<cplx>(op, z, w, fused, nullptr)) return -1;
            } else if (apply_A<cplx>(op, vj, w, fused, nullptr)) return -1;
            const int k = j + 1;
            if (gs_pass(w, k)) return -1;
            for (int i = 0; i < k; ++i) h1[i] = hh[i];
            if (gs_pass(w, k)) return -1;                 // second pass: hh = corrections, hh[k] = |w'|^2 before it
            double corr = 0.0;
            for (int i = 0; i < k; ++i) {
                corr += hh[i].x * hh[i].x + hh[i].y * hh[i].y;
                H[(size_t)i * m + j] = cadd(h1[i], hh[i]);
            }
            const double hn2 = hh[k].x - corr;
            const double hnext = hn2 > 0.0 ? sqrt(hn2) : 0.0;
            H[(size_t)k * m + j] = make_double2(hnext, 0.0);
            // Givens rotations of the new column, then the one that annihilates its sub-diagonal entry
            for (int i = 0; i < j; ++i) {
                const cplx a = H[(size_t)i * m + j], b2 = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = cadd(cscale(a, cc[i]), cmul(cs[i], b2));
                H[(size_t)(i + 1) * m + j] = csub(cscale(b2, cc[i]), cmul(cconj(cs[i]), a));
            }
            {
                const cplx a = H[(size_t)j * m + j];
                const double an = sqrt(a.x * a.x + a.y * a.y), t = sqrt(an * an + hnext * hnext);
                if (t == 0.0) FDFD_FAIL("GMRES: breakdown with a zero Hessenberg column");
                cc[j] = an / t;
                // G = [c s; -conj(s) c] with c = |a| / t, s = (a / |a|) hnext / t maps (a, hnext) to (a t / |a|, 0); a = 0: swap
                cs[j] = an > 0.0 ? cscale(a, hnext / (an * t)) : make_double2(1.0, 0.0);
                H[(size_t)j * m + j] = an > 0.0 ? cscale(a, t / an) : make_double2(t, 0.0);
                H[(size_t)(j + 1) * m + j] = make_double2(0.0, 0.0);
                const cplx gj = g[j];
                g[j] = cscale(gj, cc[j]);
                g[j + 1] = cneg(cmul(cconj(cs[j]), gj));
            }
            res->iters = total + 1;
            res->relres = sqrt(g[j + 1].x * g[j + 1].x + g[j + 1].y * g[j + 1].y) / bnorm;
            if (res->relres <= tol || hnext == 0.0) done = true;
            else {
                gmres_scale_kernel<<<nblk, 256, 0, st>>>(w, n, 1.0 / hnext);
                ++g_fdfd_launches;
            }
        }
        // y = R^-1 g (host), x += M^-1 (V y)
        for (int i = j - 1; i >= 0; --i) {
            cplx t = g[i];
            for (int l = i + 1; l < j; ++l) t = csub(t, cmul(H[(size_t)i * m + l], y[l]));
            y[i] = cdiv(t, H[(size_t)i * m + i]);
        }
        for (int i = 0; i < j; ++i) hh[i] = cneg(y[i]);
        FDFD_CHECK(cudaMemcpyAsync(dc, hh.data(), sizeof(cplx) * j, cudaMemcpyHostToDevice, st));
        gmres_axpy_kernel<<<nblk, 256, 0, st>>>(u, V, vs, dc, j, n, 0.0);
        ++g_fdfd_launches;
        if (precond) {
            if (precond_solve(precond_nd, op, u, z)) return -1;
            gmres_add_kernel<<<nblk, 256, 0, st>>>(d_x, z, n);
        } else {
            gmres_add_kernel<<<nblk, 256, 0, st>>>(d_x, u, n);
        }
        ++g_fdfd_launches;
        FDFD_CHECK(cudaStreamSynchronize(st));          // hh is reused by the next cycle
        if (done) break;
    }
    // the TRUE residual of the returned iterate
    if (residual_A<cplx>(op, d_b, d_x, u, nullptr)) return -1;
    if (dots<cplx>(st, u, u, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &hs)) return -1;
    res->relres = sqrt(hs.x) / bnorm;
    res->converged = res->relres <= tol * 10 ? 1 : 0;
    return 0;
}

int krylov_bicgstab(const FdfdOp* op, NdSolver* precond, const cplx* d_b, cplx* d_x, double tol, int maxiter,
                    int fused, int check_every, const cplx* c12, int real_inner, KrylovResult* res) {
    return krylov_bicgstab_t<cplx>(op, precond, d_b, d_x, tol, maxiter, fused, check_every, c12, real_inner, res);
}
int krylov_cocg(const FdfdOp* op, const cplx* d_b, cplx* d_x, double tol, int maxiter, int fused, int check_every,
                KrylovResult* res) {
    return krylov_cocg_t<cplx>(op, d_b, d_x, tol, maxiter, fused, check_every, res);
}
int krylov_bicgstab_c64(const FdfdOp* op, const cplx32* d_b, cplx32* d_x, double tol, int maxiter, int fused,
                        int check_every, KrylovResult* res) {
    return krylov_bicgstab_t<cplx32>(op, nullptr, d_b, d_x, tol, maxiter, fused, check_every, nullptr, 0, res);
}
int krylov_cocg_c64(const FdfdOp* op, const cplx32* d_b, cplx32* d_x, double tol, int maxiter, int fused,
                    int check_every, KrylovResult* res) {
    return krylov_cocg_t<cplx32>(op, d_b, d_x, tol, maxiter, fused, check_every, res);
}

// The factorisation pivots only inside 64-wide tiles, so its accuracy is not guaranteed the way a pivoted sparse LU's
// is (linalg.py:139-146).  What IS guaranteed is the residual contract: the fp64 stencil residual is evaluated after
// every substitution, refinement steps run while it is above `tol`, and if it still exceeds FDFD_RESIDUAL_LIMIT after
// max_refine >= 1 steps the system is handed to BiCGSTAB right-preconditioned by the same factors; a solve that
// misses the limit even then FAILS instead of returning a wrong field.  max_refine == 0 is the raw substitution
// (residual reported, no guard).  steps_out: refinement steps, + 1000 + Krylov iterations if the fallback ran.
#define FDFD_RESIDUAL_LIMIT 1e-10
#define FDFD_FALLBACK_ITERS 60

static int residual_norms(NdSolver* nd, const FdfdOp* op, const cplx* d_b, const cplx* d_x, cplx* r, int nrhs,
                          const std::vector<double>& bn, cplx* partial, cplx* sc, std::vector<double>& rel) {
    const size_t n = op->n();
    cplx h;
    if (op_residual(op, d_b, d_x, r, nrhs)) return -1;
    for (int j = 0; j < nrhs; ++j) {
        if (dev_dot(op->stream, r + j * n, r + j * n, n, true, partial, sc, 0, nullptr)) return -1;
        if (host_scalar(op->stream, sc, &h)) return -1;
        double v = bn[j] > 0 ? sqrt(h.x) / bn[j] : 0.0;
        rel[j] = (v == v) ? v : 1e300;        // NaN -> "infinitely bad"
    }
    return 0;
}

int refine_solve(NdSolver* nd, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs, int max_refine, double tol,
                 double* relres_out, int* steps_out) {
    const size_t n = op->n();
    cudaStream_t st = op->stream;
    // residual / correction vectors live in the solver handle (grown on demand, never freed per call)
    const size_t need = 2 * n * nrhs + 2 * RED_BLOCKS + S_COUNT;
    if (need > nd->ws_refine_cap) {
        if (nd->ws_refine) cudaFree(nd->ws_refine);
        nd->ws_refine = nullptr;
        nd->ws_refine_cap = 0;
        FDFD_CHECK(cudaMalloc(&nd->ws_refine, sizeof(cplx) * need));
        nd->ws_refine_cap = need;
    }
    cplx *r = nd->ws_refine, *d = r + n * nrhs, *partial = d + n * nrhs, *sc = partial + 2 * RED_BLOCKS;
    std::vector<double> bn(nrhs), rel(nrhs, 0.0);
    cplx h;
    for (int j = 0; j < nrhs; ++j) {
        if (dev_dot(st, d_b + j * n, d_b + j * n, n, true, partial, sc, 0, nullptr)) return -1;
        if (host_scalar(st, sc, &h)) return -1;
        bn[j] = sqrt(h.x);
        if (!(bn[j] == bn[j]) || bn[j] > 1e300) FDFD_FAIL("direct solve: right-hand side %d is not finite", j);
    }
    if (nd_solve(nd, op, d_b, d_x, nrhs)) return -1;
    double worst = 0.0;
    int step = 0;
    for (;; ++step) {
        if (residual_norms(nd, op, d_b, d_x, r, nrhs, bn, partial, sc, rel)) return -1;
        worst = *std::max_element(rel.begin(), rel.end());
        if (worst <= tol || step >= max_refine || worst >= 1e300) break;
        if (nd_solve(nd, op, r, d, nrhs)) return -1;
        { axpy_one_kernel<<<ceil_div(n * nrhs, 256), 256, 0, st>>>(d_x, d, n * nrhs); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    int fallback_iters = 0;
    const double limit = std::max(tol, FDFD_RESIDUAL_LIMIT);
    if (max_refine >= 1 && worst > limit) {
        // refinement stalled: Krylov on the same factors (they are a preconditioner even when they are not a solver)
        for (int j = 0; j < nrhs; ++j) {
            if (rel[j] <= limit) continue;
            if (rel[j] >= 1e300) FDFD_CHECK(cudaMemsetAsync(d_x + j * n, 0, sizeof(cplx) * n, st));
            KrylovResult kr;
            if (krylov_bicgstab(op, nd, d_b + j * n, d_x + j * n, std::max(tol, 1e-12), FDFD_FALLBACK_ITERS, 0, 1, nullptr, 0,
                                &kr))
                return -1;
            fallback_iters += kr.iters;
        }
        if (residual_norms(nd, op, d_b, d_x, r, nrhs, bn, partial, sc, rel)) return -1;
        worst = *std::max_element(rel.begin(), rel.end());
        if (worst > limit)
            FDFD_FAIL("direct solve: relative residual %.3e is above %.1e after %d refinement steps and %d BiCGSTAB "
                      "iterations on the factors (structure too ill-conditioned for the unpivoted block factorisation)",
                      worst, limit, step, fallback_iters);
    }
    *relres_out = worst;
    *steps_out = step + (fallback_iters ? 1000 + fallback_iters : 0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// Device-resident Born / Newton iterations for the Kerr problem (nonlinear_solvers.py:13-110, simulation.py:57-68).
// Every Kerr term of the reference is  eps_nl = 3 chi region |E|^2 w(eps_r): a sum of them is K(x) |E|^2 with one
// complex coefficient plane K, and  d eps_nl / dE = K conj(E).  The permittivity update, the Newton right-hand side
// f(E) = (A + Anl(E)) E - b, the Jacobian's anti-linear diagonal and the convergence norm are kernels; the host sees
// one scalar per iteration (plus the residual norms the inner Krylov solve checks).
// ------------------------------------------------------------------------------------------
// eps_out = scale * K |E|^2 ; c12 (optional) = kfac * conj(K conj(E)) * E = kfac * conj(K) * E * E
__global__ void kerr_terms_kernel(const cplx* __restrict__ K, const cplx* __restrict__ E, cplx* __restrict__ eps_out,
                                  double scale, cplx* __restrict__ c12, double kfac, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cplx k = K[i], e = E[i];
    const double e2 = e.x * e.x + e.y * e.y;
    eps_out[i] = make_double2(scale * k.x * e2, scale * k.y * e2);
    if (c12) c12[i] = cscale(cmul(cconj(k), cmul(e, e)), kfac);
}
// x = a - b (out may alias a); partial sums of |x|^2-type norms are taken by dots() afterwards
__global__ void sub_kernel(cplx* __restrict__ out, const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = csub(a[i], b[i]);
}
__global__ void neg_kernel(cplx* __restrict__ v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = cneg(v[i]);
}

// one solve with the work operator (planes = A + diagonal perturbation), optionally with the anti-linear term c12:
// strategy 0 ("reuse"): BiCGSTAB preconditioned by the LINEAR operator's cached factors, falling back to an exact
// factorisation of the work operator when that stalls; strategy 1 ("refactor"): always the exact factorisation.
static int nl_inner_solve(FdfdOp* op_nl, NdSolver* lin, NdSolver* work, const cplx* b, cplx* x, const cplx* c12,
                          int strategy, int* inner_iters) {
    KrylovResult kr;
    kr.iters = 0; kr.relres = 1.0; kr.converged = 0;
    if (strategy == 0) {
        if (krylov_bicgstab(op_nl, lin, b, x, 1e-13, 40, 0, 1, c12, 0, &kr)) return -1;
        *inner_iters += kr.iters;
        if (kr.relres <= 1e-11) return 0;
    }
    if (nd_factor(work, op_nl)) return -1;
    if (!c12) {
        double rr;
        int steps;
        return refine_solve(work, op_nl, b, x, 1, 3, 1e-12, &rr, &steps);
    }
    if (krylov_bicgstab(op_nl, work, b, x, 1e-13, 200, 0, 1, c12, 0, &kr)) return -1;
    *inner_iters += kr.iters;
    if (kr.relres > 1e-9) FDFD_FAIL("Newton Jacobian solve did not converge (relres %.2e after %d iterations)", kr.relres, kr.iters);
    return 0;
}

int nl_solve(FdfdOp* op_nl, NdSolver* lin, NdSolver* work, const cplx* d_K, const cplx* d_b, cplx* d_E, int method,
             int strategy, double thr, int max_iter, double* conv, int* iters_out, int* inner_iters_out) {
    const size_t n = op_nl->n();
    cudaStream_t st = op_nl->stream;
    if (op_nl->pol != 0) FDFD_FAIL("the nonlinear solvers are defined for Ez (nonlinear_solvers.py:18)");
    if (strategy == 0 && (!lin || !lin->factored)) FDFD_FAIL("nonlinear solve ('reuse'): the linear operator is not factorised");
    Scratch ws;
    FDFD_CHECK(cudaMalloc(&ws.base, sizeof(cplx) * (4 * n + 2 * RED_BLOCKS + S_COUNT)));
    cplx *x = static_cast<cplx*>(ws.base), *f = x + n, *c12 = f + n, *dl = c12 + n, *partial = dl + n, *sc = partial + 2 * RED_BLOCKS;
    const double kfac = op_nl->omega * op_nl->omega * FDFD_EPS0 * op_nl->L0;
    const int nblk = ceil_div(n, 256);
    cplx h;
    int it = 0, inner = 0;
    for (int i = 0; i < max_iter; ++i) conv[i] = 0.0;
    for (it = 0; it < max_iter; ++it) {
        double cv;
        if (method == 0) {
            // Born: E <- (A + Anl(E))^-1 b
            { kerr_terms_kernel<<<nblk, 256, 0, st>>>(d_K, d_E, op_nl->eps_nl, 1.0, nullptr, 0.0, n); ++g_fdfd_launches; }
            if (op_assemble_dev(op_nl, op_nl->eps_r, op_nl->eps_nl, op_nl->averaging)) return -1;
            FDFD_CHECK(cudaMemcpyAsync(x, d_E, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));   // initial guess
            if (nl_inner_solve(op_nl, lin, work, d_b, x, nullptr, strategy, &inner)) return -1;
            { sub_kernel<<<nblk, 256, 0, st>>>(dl, x, d_E, n); ++g_fdfd_launches; }
            if (dots<cplx>(st, dl, dl, 1, x, x, 1, n, partial, sc, POST_STORE, S_RR, S_TT, 1, nullptr)) return -1;
            FDFD_CHECK(cudaMemcpyAsync(d_E, x, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
        } else {
            // Newton: f = (A + Anl(E)) E - b ;  J11 = A + diag(k (eps_nl + dnl_de E)) = A(eps_nl -> 2 K |E|^2),
            // J12 = diag(k conj(K) E^2) acting on conj(dE);  E <- E - dE
            { kerr_terms_kernel<<<nblk, 256, 0, st>>>(d_K, d_E, op_nl->eps_nl, 1.0, c12, kfac, n); ++g_fdfd_launches; }
            if (op_assemble_dev(op_nl, op_nl->eps_r, op_nl->eps_nl, op_nl->averaging)) return -1;
            if (op_residual(op_nl, d_b, d_E, f, 1)) return -1;                  // b - (A + Anl) E
            { neg_kernel<<<nblk, 256, 0, st>>>(f, n); ++g_fdfd_launches; }
            if (dots<cplx>(st, f, f, 1, nullptr, nullptr, 0, n, partial, sc, POST_STORE, S_RR, -1, 1, nullptr)) return -1;
            if (host_scalar(st, sc + S_RR, &h)) return -1;
            if (h.x > 0.0) {
                { kerr_terms_kernel<<<nblk, 256, 0, st>>>(d_K, d_E, op_nl->eps_nl, 2.0, nullptr, 0.0, n); ++g_fdfd_launches; }
                if (op_assemble_dev(op_nl, op_nl->eps_r, op_nl->eps_nl, op_nl->averaging)) return -1;
                FDFD_CHECK(cudaMemsetAsync(dl, 0, sizeof(cplx) * n, st));
                if (nl_inner_solve(op_nl, lin, work, f, dl, c12, strategy, &inner)) return -1;
            } else {
                FDFD_CHECK(cudaMemsetAsync(dl, 0, sizeof(cplx) * n, st));
            }
            { sub_kernel<<<nblk, 256, 0, st>>>(d_E, d_E, dl, n); ++g_fdfd_launches; }
            if (dots<cplx>(st, dl, dl, 1, d_E, d_E, 1, n, partial, sc, POST_STORE, S_RR, S_TT, 1, nullptr)) return -1;
        }
        FDFD_CHECK(cudaGetLastError());
        cplx nrm[2];
        FDFD_CHECK(cudaMemcpyAsync(&nrm[0], sc + S_RR, sizeof(cplx), cudaMemcpyDeviceToHost, st));
        FDFD_CHECK(cudaMemcpyAsync(&nrm[1], sc + S_TT, sizeof(cplx), cudaMemcpyDeviceToHost, st));
        FDFD_CHECK(cudaStreamSynchronize(st));
        cv = sqrt(nrm[0].x) / sqrt(nrm[1].x);
        conv[it] = cv;
        if (cv < thr) { ++it; break; }
    }
    *iters_out = it;
    *inner_iters_out = inner;
    return 0;
}
