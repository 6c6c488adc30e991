// Krylov drivers on top of the matrix-free stencil: BiCGSTAB (optionally right-preconditioned by
// the cached direct factorisation), COCG on the symmetrised operator, and iterative refinement.
// All iteration scalars live on the device; the host only enqueues kernels and looks at the
// residual norm every `check_every` iterations.
#include <algorithm>
#include "krylov.cuh"

#define RED_BLOCKS 592   // 148 SMs x 4
#define RED_THREADS 256

template <bool CONJ_A>
__global__ void __launch_bounds__(RED_THREADS)
dot_partial_kernel(const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n, cplx* __restrict__ partial) {
    __shared__ double sx[RED_THREADS / 32], sy[RED_THREADS / 32];
    double ax = 0.0, ay = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx u = a[i], v = b[i];
        if (CONJ_A) u.y = -u.y;
        ax += u.x * v.x - u.y * v.y;
        ay += u.x * v.y + u.y * v.x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_down_sync(0xffffffffu, ax, o);
        ay += __shfl_down_sync(0xffffffffu, ay, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sx[w] = ax; sy[w] = ay; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tx = 0, ty = 0;
        for (int i = 0; i < RED_THREADS / 32; ++i) { tx += sx[i]; ty += sy[i]; }
        partial[blockIdx.x] = make_double2(tx, ty);
    }
}

__global__ void __launch_bounds__(RED_THREADS)
dot_final_kernel(const cplx* __restrict__ partial, int count, cplx* __restrict__ out, int real_only) {
    __shared__ double sx[RED_THREADS], sy[RED_THREADS];
    double ax = 0, ay = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) { ax += partial[i].x; ay += partial[i].y; }
    sx[threadIdx.x] = ax; sy[threadIdx.x] = ay;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sx[threadIdx.x] += sx[threadIdx.x + s]; sy[threadIdx.x] += sy[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = make_double2(sx[0], real_only ? 0.0 : sy[0]);
}

int dev_dot(cudaStream_t st, const cplx* a, const cplx* b, size_t n, bool conj_a, cplx* partial, cplx* out,
            int real_only, FdfdComm* comm) {
    if (conj_a) { dot_partial_kernel<true><<<RED_BLOCKS, RED_THREADS, 0, st>>>(a, b, n, partial); ++g_fdfd_launches; }
    else { dot_partial_kernel<false><<<RED_BLOCKS, RED_THREADS, 0, st>>>(a, b, n, partial); ++g_fdfd_launches; }
    { dot_final_kernel<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, out, real_only); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    // slabs: the scalar is summed over the ranks on the same stream, it never visits the host
    if (comm && comm->world > 1 && comm_allreduce_sum(comm, out, 2, st)) return -1;
    return 0;
}

// scalar slots
enum { S_RHO = 0, S_RHO_OLD, S_ALPHA, S_OMEGA, S_BETA, S_R0V, S_TS, S_TT, S_RR, S_PQ, S_COUNT };

__global__ void bicg_beta_kernel(cplx* sc) {   // beta = (rho/rho_old) * (alpha/omega); rho_old = rho
    cplx beta = cmul(cdiv(sc[S_RHO], sc[S_RHO_OLD]), cdiv(sc[S_ALPHA], sc[S_OMEGA]));
    sc[S_BETA] = beta;
    sc[S_RHO_OLD] = sc[S_RHO];
}
__global__ void bicg_alpha_kernel(cplx* sc) { sc[S_ALPHA] = cdiv(sc[S_RHO], sc[S_R0V]); }
__global__ void bicg_omega_kernel(cplx* sc) { sc[S_OMEGA] = cdiv(sc[S_TS], sc[S_TT]); }
__global__ void cocg_alpha_kernel(cplx* sc) { sc[S_ALPHA] = cdiv(sc[S_RHO], sc[S_PQ]); }
__global__ void cocg_beta_kernel(cplx* sc) {   // on entry S_RR holds the new r^T r
    sc[S_BETA] = cdiv(sc[S_RR], sc[S_RHO]);
    sc[S_RHO] = sc[S_RR];
}

// p = r + beta (p - omega v)
__global__ void bicg_p_kernel(cplx* __restrict__ p, const cplx* __restrict__ r, const cplx* __restrict__ v,
                              const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx beta = sc[S_BETA], om = sc[S_OMEGA];
    cplx t = csub(p[i], cmul(om, v[i]));
    p[i] = cadd(r[i], cmul(beta, t));
}
// s = r - alpha v
__global__ void bicg_s_kernel(cplx* __restrict__ s, const cplx* __restrict__ r, const cplx* __restrict__ v,
                              const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    s[i] = csub(r[i], cmul(sc[S_ALPHA], v[i]));
}
// x += alpha ph + omega sh ; r = s - omega t
__global__ void bicg_xr_kernel(cplx* __restrict__ x, cplx* __restrict__ r, const cplx* __restrict__ ph,
                               const cplx* __restrict__ sh, const cplx* __restrict__ s, const cplx* __restrict__ t,
                               const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx al = sc[S_ALPHA], om = sc[S_OMEGA];
    cplx xv = x[i];
    cfma(xv, al, ph[i]);
    cfma(xv, om, sh[i]);
    x[i] = xv;
    r[i] = csub(s[i], cmul(om, t[i]));
}
// x += alpha p ; r -= alpha q
__global__ void cocg_xr_kernel(cplx* __restrict__ x, cplx* __restrict__ r, const cplx* __restrict__ p,
                               const cplx* __restrict__ q, const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx al = sc[S_ALPHA];
    cplx xv = x[i];
    cfma(xv, al, p[i]);
    x[i] = xv;
    r[i] = csub(r[i], cmul(al, q[i]));
}
// p = r + beta p
__global__ void cocg_p_kernel(cplx* __restrict__ p, const cplx* __restrict__ r, const cplx* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = cadd(r[i], cmul(sc[S_BETA], p[i]));
}
// v[i] *= sxf[ix] * syf[iy]  = v / (isxf isyf): left scaling that makes A complex symmetric
__global__ void sym_scale_kernel(cplx* __restrict__ v, const cplx* __restrict__ isxf, const cplx* __restrict__ isyf,
                                 int nx, int ny) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nx * ny) return;
    int ix = (int)(i / ny), iy = (int)(i % ny);
    v[i] = cdiv(v[i], cmul(isxf[ix], isyf[iy]));
}
__global__ void axpy_one_kernel(cplx* __restrict__ x, const cplx* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = cadd(x[i], d[i]);
}

struct Scratch {
    cplx* base = nullptr;
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

static int apply_A(const FdfdOp* op, const cplx* x, cplx* y, int fused, const cplx* c12) {
    const size_t pad = kpad(op);
    if (fused ? op_apply_fused(op, x - pad, y - pad, 1) : op_apply_planes(op, x - pad, y - pad, 1)) return -1;
    if (c12) {
        { conj_couple_kernel<<<ceil_div(op->n(), 256), 256, 0, op->stream>>>(y, c12, x, op->n(), 0); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    return 0;
}

static int residual_A(const FdfdOp* op, const cplx* b, const cplx* x, cplx* r, const cplx* c12) {
    const size_t pad = kpad(op);
    if (op_residual(op, b - pad, x - pad, r - pad, 1)) return -1;
    if (c12) {
        { conj_couple_kernel<<<ceil_div(op->n(), 256), 256, 0, op->stream>>>(r, c12, x, op->n(), 1); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    return 0;
}

int krylov_bicgstab(const FdfdOp* op, NdSolver* precond, const cplx* d_b, cplx* d_x, double tol, int maxiter,
                    int fused, int check_every, const cplx* c12, int real_inner, KrylovResult* res) {
    const int RI = (real_inner || c12) ? 1 : 0;   // an R-linear operator needs the real inner product
    const size_t n = kn(op), pad = kpad(op), vs = n + 2 * pad;
    if (op->halo && (precond || c12)) FDFD_FAIL("slab operators take neither a preconditioner nor an anti-linear term");
    FdfdComm* comm = op->comm;
    d_b += pad; d_x += pad;
    cudaStream_t st = op->stream;
    const int nvec = precond ? 8 : 6;
    Scratch ws;
    FDFD_CHECK(cudaMalloc(&ws.base, sizeof(cplx) * (vs * nvec + RED_BLOCKS + S_COUNT)));
    if (pad) FDFD_CHECK(cudaMemsetAsync(ws.base, 0, sizeof(cplx) * vs * nvec, st));
    cplx *r = ws.base + pad, *r0 = r + vs, *p = r0 + vs, *v = p + vs, *s = v + vs, *t = s + vs;
    cplx *ph = precond ? t + vs : p, *sh = precond ? ph + vs : s;
    cplx* partial = ws.base + vs * nvec;
    cplx* sc = partial + RED_BLOCKS;
    const int nblk = ceil_div(n, 256);
    cplx h;
    if (check_every < 1) check_every = 1;
    // r = b - A x
    if (residual_A(op, d_b, d_x, r, c12)) return -1;
    FDFD_CHECK(cudaMemcpyAsync(r0, r, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
    FDFD_CHECK(cudaMemsetAsync(p, 0, sizeof(cplx) * n, st));
    FDFD_CHECK(cudaMemsetAsync(v, 0, sizeof(cplx) * n, st));
    cplx init[S_COUNT];
    for (int i = 0; i < S_COUNT; ++i) init[i] = make_double2(1.0, 0.0);
    FDFD_CHECK(cudaMemcpyAsync(sc, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (dev_dot(st, d_b, d_b, n, true, partial, sc + S_RR, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    const double bnorm = sqrt(h.x);
    res->iters = 0; res->converged = 0; res->relres = 1.0;
    if (bnorm == 0.0) {
        FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(cplx) * n, st));
        res->converged = 1; res->relres = 0.0;
        return 0;
    }
    if (dev_dot(st, r, r, n, true, partial, sc + S_RR, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    res->relres = sqrt(h.x) / bnorm;
    if (res->relres <= tol) { res->converged = 1; return 0; }
    for (int it = 1; it <= maxiter; ++it) {
        if (dev_dot(st, r0, r, n, true, partial, sc + S_RHO, RI, comm)) return -1;
        { bicg_beta_kernel<<<1, 1, 0, st>>>(sc); ++g_fdfd_launches; }
        { bicg_p_kernel<<<nblk, 256, 0, st>>>(p, r, v, sc, n); ++g_fdfd_launches; }
        if (precond) { if (nd_solve(precond, op, p, ph, 1)) return -1; }
        if (apply_A(op, ph, v, fused, c12)) return -1;
        if (dev_dot(st, r0, v, n, true, partial, sc + S_R0V, RI, comm)) return -1;
        { bicg_alpha_kernel<<<1, 1, 0, st>>>(sc); ++g_fdfd_launches; }
        { bicg_s_kernel<<<nblk, 256, 0, st>>>(s, r, v, sc, n); ++g_fdfd_launches; }
        if (precond) { if (nd_solve(precond, op, s, sh, 1)) return -1; }
        if (apply_A(op, sh, t, fused, c12)) return -1;
        if (dev_dot(st, t, s, n, true, partial, sc + S_TS, RI, comm)) return -1;
        if (dev_dot(st, t, t, n, true, partial, sc + S_TT, RI, comm)) return -1;
        { bicg_omega_kernel<<<1, 1, 0, st>>>(sc); ++g_fdfd_launches; }
        { bicg_xr_kernel<<<nblk, 256, 0, st>>>(d_x, r, ph, sh, s, t, sc, n); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
        res->iters = it;
        if (it % check_every == 0 || it == maxiter) {
            if (dev_dot(st, r, r, n, true, partial, sc + S_RR, RI, comm)) return -1;
            if (host_scalar(st, sc + S_RR, &h)) return -1;
            res->relres = sqrt(h.x) / bnorm;
            if (!(res->relres == res->relres)) break;           // NaN: breakdown
            if (res->relres <= tol) { res->converged = 1; break; }
        }
    }
    // report the TRUE residual of the returned iterate
    if (residual_A(op, d_b, d_x, r, c12)) return -1;
    if (dev_dot(st, r, r, n, true, partial, sc + S_RR, RI, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    res->relres = sqrt(h.x) / bnorm;
    res->converged = res->relres <= tol * 10 ? res->converged : 0;
    return 0;
}

int krylov_cocg(const FdfdOp* op, const cplx* d_b, cplx* d_x, double tol, int maxiter, int fused, int check_every,
                KrylovResult* res) {
    const size_t n = kn(op), pad = kpad(op), vs = n + 2 * pad;
    FdfdComm* comm = op->comm;
    d_b += pad; d_x += pad;
    const cplx *sxf = op->isxf + op->halo, *syf = op->isyf;
    const int nxo = op->nx - 2 * op->halo;
    cudaStream_t st = op->stream;
    Scratch ws;
    FDFD_CHECK(cudaMalloc(&ws.base, sizeof(cplx) * (vs * 4 + RED_BLOCKS + S_COUNT)));
    if (pad) FDFD_CHECK(cudaMemsetAsync(ws.base, 0, sizeof(cplx) * vs * 4, st));
    cplx *r = ws.base + pad, *p = r + vs, *q = p + vs, *bs = q + vs;
    cplx* partial = ws.base + vs * 4;
    cplx* sc = partial + RED_BLOCKS;
    const int nblk = ceil_div(n, 256);
    cplx h;
    if (check_every < 1) check_every = 1;
    // symmetrised system  D A x = D b,  D = diag(sxf[ix] syf[iy])
    FDFD_CHECK(cudaMemcpyAsync(bs, d_b, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
    { sym_scale_kernel<<<nblk, 256, 0, st>>>(bs, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
    if (residual_A(op, d_b, d_x, r, nullptr)) return -1;
    { sym_scale_kernel<<<nblk, 256, 0, st>>>(r, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
    FDFD_CHECK(cudaMemcpyAsync(p, r, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, st));
    if (dev_dot(st, bs, bs, n, true, partial, sc + S_RR, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    const double bnorm = sqrt(h.x);
    res->iters = 0; res->converged = 0; res->relres = 1.0;
    if (bnorm == 0.0) {
        FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(cplx) * n, st));
        res->converged = 1; res->relres = 0.0;
        return 0;
    }
    if (dev_dot(st, r, r, n, false, partial, sc + S_RHO, 0, comm)) return -1;
    for (int it = 1; it <= maxiter; ++it) {
        if (apply_A(op, p, q, fused, nullptr)) return -1;
        { sym_scale_kernel<<<nblk, 256, 0, st>>>(q, sxf, syf, nxo, op->ny); ++g_fdfd_launches; }
        if (dev_dot(st, p, q, n, false, partial, sc + S_PQ, 0, comm)) return -1;
        { cocg_alpha_kernel<<<1, 1, 0, st>>>(sc); ++g_fdfd_launches; }
        { cocg_xr_kernel<<<nblk, 256, 0, st>>>(d_x, r, p, q, sc, n); ++g_fdfd_launches; }
        if (dev_dot(st, r, r, n, false, partial, sc + S_RR, 0, comm)) return -1;
        { cocg_beta_kernel<<<1, 1, 0, st>>>(sc); ++g_fdfd_launches; }
        { cocg_p_kernel<<<nblk, 256, 0, st>>>(p, r, sc, n); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
        res->iters = it;
        if (it % check_every == 0 || it == maxiter) {
            if (dev_dot(st, r, r, n, true, partial, sc + S_TT, 0, comm)) return -1;
            if (host_scalar(st, sc + S_TT, &h)) return -1;
            res->relres = sqrt(h.x) / bnorm;
            if (!(res->relres == res->relres)) break;
            if (res->relres <= tol) { res->converged = 1; break; }
        }
    }
    // true residual in the ORIGINAL (unscaled) system
    if (residual_A(op, d_b, d_x, r, nullptr)) return -1;
    if (dev_dot(st, r, r, n, true, partial, sc + S_RR, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    double rn = sqrt(h.x);
    if (dev_dot(st, d_b, d_b, n, true, partial, sc + S_RR, 0, comm)) return -1;
    if (host_scalar(st, sc + S_RR, &h)) return -1;
    res->relres = rn / sqrt(h.x);
    return 0;
}

int refine_solve(NdSolver* nd, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs, int max_refine, double tol,
                 double* relres_out, int* steps_out) {
    const size_t n = op->n();
    cudaStream_t st = op->stream;
    // residual / correction vectors live in the solver handle (grown on demand, never freed per call)
    const size_t need = 2 * n * nrhs + RED_BLOCKS + 2;
    if (need > nd->ws_refine_cap) {
        if (nd->ws_refine) cudaFree(nd->ws_refine);
        nd->ws_refine = nullptr;
        nd->ws_refine_cap = 0;
        FDFD_CHECK(cudaMalloc(&nd->ws_refine, sizeof(cplx) * need));
        nd->ws_refine_cap = need;
    }
    cplx *r = nd->ws_refine, *d = r + n * nrhs, *partial = d + n * nrhs, *sc = partial + RED_BLOCKS;
    std::vector<double> bn(nrhs);
    cplx h;
    for (int j = 0; j < nrhs; ++j) {
        if (dev_dot(st, d_b + j * n, d_b + j * n, n, true, partial, sc, 0, nullptr)) return -1;
        if (host_scalar(st, sc, &h)) return -1;
        bn[j] = sqrt(h.x);
    }
    if (nd_solve(nd, op, d_b, d_x, nrhs)) return -1;
    double worst = 0.0;
    int step = 0;
    for (;; ++step) {
        if (op_residual(op, d_b, d_x, r, nrhs)) return -1;
        worst = 0.0;
        for (int j = 0; j < nrhs; ++j) {
            if (dev_dot(st, r + j * n, r + j * n, n, true, partial, sc, 0, nullptr)) return -1;
            if (host_scalar(st, sc, &h)) return -1;
            double rel = bn[j] > 0 ? sqrt(h.x) / bn[j] : 0.0;
            if (!(rel == rel)) rel = 1e300;
            worst = std::max(worst, rel);
        }
        if (worst <= tol || step >= max_refine) break;
        if (nd_solve(nd, op, r, d, nrhs)) return -1;
        { axpy_one_kernel<<<ceil_div(n * nrhs, 256), 256, 0, st>>>(d_x, d, n * nrhs); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    *relres_out = worst;
    *steps_out = step;
    return 0;
}
