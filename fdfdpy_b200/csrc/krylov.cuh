// Krylov / refinement drivers (see krylov.cu)
#pragma once
#include <vector>
#include "direct.cuh"

struct KrylovResult {
    int iters;
    int converged;
    double relres;     // ||b - A x|| / ||b|| of the returned iterate (true residual)
};

int dev_dot(cudaStream_t st, const cplx* a, const cplx* b, size_t n, bool conj_a, cplx* partial, cplx* out,
            int real_only, FdfdComm* comm = nullptr);
int krylov_bicgstab(const FdfdOp* op, NdSolver* precond, const cplx* d_b, cplx* d_x, double tol, int maxiter,
                    int fused, int check_every, const cplx* c12, int real_inner, KrylovResult* res);
int krylov_cocg(const FdfdOp* op, const cplx* d_b, cplx* d_x, double tol, int maxiter, int fused, int check_every,
                KrylovResult* res);
int refine_solve(NdSolver* nd, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs, int max_refine, double tol,
                 double* relres_out, int* steps_out);
