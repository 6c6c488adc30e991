// Krylov / refinement drivers (see krylov.cu)
#pragma once
#include <vector>
#include "direct.cuh"

struct KrylovResult {
    int iters;
    int converged;
    double relres;     // ||b - A x|| / ||b|| of the returned iterate (true residual)
};

int dev_dot(cudaStream_t st, const cplx* a, const cplx* b, size_t n, bool conj_a, cplx* partial, cplx* sc,
            int real_only, FdfdComm* comm = nullptr);
// complex64 vectors (fp64 arithmetic and scalars): no preconditioner, no anti-linear term
int krylov_bicgstab_c64(const FdfdOp* op, const cplx32* d_b, cplx32* d_x, double tol, int maxiter, int fused,
                        int check_every, KrylovResult* res);
int krylov_cocg_c64(const FdfdOp* op, const cplx32* d_b, cplx32* d_x, double tol, int maxiter, int fused,
                    int check_every, KrylovResult* res);
int krylov_bicgstab(const FdfdOp* op, NdSolver* precond, const cplx* d_b, cplx* d_x, double tol, int maxiter,
                    int fused, int check_every, const cplx* c12, int real_inner, KrylovResult* res);
int krylov_cocg(const FdfdOp* op, const cplx* d_b, cplx* d_x, double tol, int maxiter, int fused, int check_every,
                KrylovResult* res);
// restarted GMRES, right-preconditioned by `precond` (whole-grid factors) or by a slab operator's Schwarz preconditioner
int krylov_gmres(const FdfdOp* op, NdSolver* precond, const cplx* d_b, cplx* d_x, double tol, int maxiter, int restart,
                 int fused, KrylovResult* res);
// restricted additive Schwarz preconditioner of a slab operator (sub == nullptr detaches); see krylov.cu
int schwarz_attach(FdfdOp* slab, FdfdOp* sub, NdSolver* nd, int overlap, int npml_sub);
int refine_solve(NdSolver* nd, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs, int max_refine, double tol,
                 double* relres_out, int* steps_out);
// device-resident Born (method 0) / Newton (method 1) iteration of the Kerr problem; see krylov.cu
int nl_solve(FdfdOp* op_nl, NdSolver* lin, NdSolver* work, const cplx* d_K, const cplx* d_b, cplx* d_E, int method,
             int strategy, double thr, int max_iter, double* conv, int* iters_out, int* inner_iters_out);
