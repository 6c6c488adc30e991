/* fdfd_b200.h -- C ABI of the B200-native fdfdpy hot path (libfdfd_b200.so).
 *
 * Plain pointers and sizes only; complex numbers are interleaved (re, im) doubles, i.e. numpy
 * complex128 / C99 double _Complex.  Every function returns 0 on success and -1 on failure;
 * fdfd_last_error() then holds a message (thread local).  "host" pointers are ordinary CPU
 * memory (copied in/out inside the call); "dev" pointers are CUDA device pointers of the current
 * device (zero-copy; e.g. torch.Tensor.data_ptr()).  Fields are (Nx, Ny) C-ordered, y fastest.
 *
 * Each entry point names the reference interface it replaces (fancompute/fdfdpy, file:line).
 */
#ifndef FDFD_B200_H
#define FDFD_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct FdfdOp fdfd_op;         /* the Maxwell operator A on the device               */
typedef struct NdSolver fdfd_direct;   /* structured direct solver: plan + cached factorisation */
typedef struct FdfdComm fdfd_comm;     /* one-process-per-GPU communicator (NCCL over NVLink/NVSwitch) */

int fdfd_version(void);
const char* fdfd_last_error(void);
/* device bookkeeping helpers used by the Python host and the benchmark */
int fdfd_device_count(int* count);
int fdfd_set_device(int device);
int fdfd_mem_info(double* free_bytes, double* total_bytes);
int fdfd_malloc(void** dev_ptr, double bytes);
int fdfd_free(void* dev_ptr);
int fdfd_memcpy_h2d(void* dev, const void* host, double bytes);
int fdfd_memcpy_d2h(void* host, const void* dev, double bytes);
int fdfd_op_sync(fdfd_op* op);
/* number of CUDA kernels this library has launched so far (reset != 0 zeroes the counter) */
double fdfd_launch_count(int reset);
/* measurement helpers (bench.py): CUDA-event timer on the operator's stream; live per-launch timing
 * of the GEMM kernels (totals: ms, real flops, launches for the 64x64- and the 32x32-tile kernel);
 * and a register-resident DMMA loop that measures the FP64 tensor-pipe ceiling of this board. */
int fdfd_timer_start(fdfd_op* op);
int fdfd_timer_stop(fdfd_op* op, double* ms);
int fdfd_gemm_timing(int enable);
int fdfd_gemm_timing_read(double* out6);
int fdfd_gemm_timing_exec_flops(double* flops);   /* flops the tensor pipe executed (3M: 6 per complex MAC) */
int fdfd_dmma_peak(double* tflops);
int fdfd_phase_timing(int enable);
int fdfd_phase_timing_read(double* out13);   /* ms: assemble,pivot,panel,rowgemm,copy,update,expand,solve_fwd,solve_bwd,stencil,ggemm,schur,small */
int fdfd_phase_timing_read_levels(double* out, int max_levels);   /* out[level * 13 + phase] */
int fdfd_dmma_probe(int warps_per_sm, int independent_accumulators, double* tflops);
/* the same loop run for ~50 ms with the SM clock measured inside the kernel:
 * out4 = { TFLOP/s, SM MHz during the probe, ms, real flops } */
int fdfd_dmma_probe_clocked(int warps_per_sm, int independent_accumulators, double* out4);
/* issue-order probe of the 3M inner loop (24 accumulators, register operands); pattern 0..3, see capi.cu.
 * out4 = { TFLOP/s executed, SM MHz, ms, fraction of the pipe at that clock } */
int fdfd_dmma_pattern_probe(int pattern, int warps_per_sm, double* out4);
/* the GEMM's k-step from a resident shared-memory tile: mode 0 no barrier, 1 barrier per k-tile, 2 barrier + cp.async */
int fdfd_dmma_smem_probe(int mode, double* out4);
/* page-lock / unlock an existing host buffer so the *_host entry points copy at full PCIe rate */
int fdfd_host_register(void* host, double bytes);
int fdfd_host_unregister(void* host);
/* page-locked host buffers for the result arrays the Python host hands back (recycled by a pool) */
int fdfd_host_alloc(void** host_ptr, double bytes);
int fdfd_host_free(void* host_ptr);

/* ---- operator: replaces linalg.py:39 construct_A, pml.py:44 S_create, derivatives.py:7 createDws.
 * pol: 0 = 'Ez', 1 = 'Hz'.  The sc-PML inverse stretch factors are computed on the device at
 * creation; fdfd_op_assemble_* builds the five stencil planes of
 *   A = Dxf mu^-1 Dxb + Dyf mu^-1 Dyb + w^2 eps            (Ez, linalg.py:61-63)
 *   A = Dxf ex^-1 Dxb + Dyf ey^-1 Dyb + w^2 mu             (Hz, linalg.py:96-98)
 * eps_nl (may be NULL) adds the Kerr diagonal Anl of simulation.py:68-70.                         */
int fdfd_op_create(fdfd_op** out, int nx, int ny, double omega, double dl, int npml_x, int npml_y,
                   int pol, double L0);
void fdfd_op_destroy(fdfd_op* op);
int fdfd_op_assemble_host(fdfd_op* op, const double* eps_r_c128, const double* eps_nl_c128, int averaging);
int fdfd_op_assemble_dev(fdfd_op* op, const void* d_eps_r, const void* d_eps_nl, int averaging);
/* the same for a real (float64) eps_r without nonlinearity: half the host->device bytes */
int fdfd_op_assemble_host_f64(fdfd_op* op, const double* eps_r_f64, int averaging);
/* inverse stretch factors 1/s as four 1-D complex arrays (lengths nx, nx, ny, ny): pml.py:63-76 */
int fdfd_op_get_sfactors_host(fdfd_op* op, double* isxf, double* isxb, double* isyf, double* isyb);
/* the five planes c0,cxm,cxp,cym,cyp (5*nx*ny complex) for export of A as a sparse matrix */
int fdfd_op_get_planes_host(fdfd_op* op, double* planes_c128);
/* y = A x (replaces the scipy A.dot(x) calls, e.g. nonlinear_solvers.py:127). fused != 0 selects the
 * matrix-free Ez kernel that rebuilds the coefficients from eps and the PML factors.            */
int fdfd_op_apply_host(fdfd_op* op, const double* x_c128, double* y_c128, int nvec, int fused);
int fdfd_op_apply_dev(fdfd_op* op, const void* d_x, void* d_y, int nvec, int fused);
/* in-plane fields from the transverse one: simulation.py:138-176 (Hx,Hy | Ex,Ey).
 * averaging: 1/0 = Hz edge averaging on/off (solve_fields' argument), -1 = as assembled.          */
int fdfd_op_derive_fields_host(fdfd_op* op, const double* x_c128, double* f1_c128, double* f2_c128,
                               int averaging);
int fdfd_op_derive_fields_dev(fdfd_op* op, const void* d_x, void* d_f1, void* d_f2, int averaging);

/* ---- direct solver: replaces linalg.py:123 solver_direct (pyMKL pardisoSolver factor/solve,
 * scipy spsolve).  The elimination plan comes from fdfdpy_b200/ndplan.py, one call per level. */
typedef struct {
    int kind;              /* 0 leaf, 1 merge */
    int nb, kmax, mmax, ncls, child_mmax;
    const int* cls;        /* [nb]   shape class of every front */
    const int* k_cls;      /* [ncls] number of real pivots per class */
    const int* ch1;        /* [nb]   merge: first child front  */
    const int* ch2;        /* [nb]   merge: second child front */
    const int* c1map;      /* [ncls][child_mmax] merge: child ring position -> front slot (-1 pad) */
    const int* c2map;
    const int* x0;         /* [nb]   leaf: box origin */
    const int* y0;
    const int* slot_lx;    /* [ncls][kmax+mmax] leaf: slot -> local coordinates (-1 pad) */
    const int* slot_ly;
    const int* slot_right; /* leaf: slot of the +x neighbour if this slot owns its entries, else -1 */
    const int* slot_up;
    int send_to, recv_from; /* sharded tree (ndplan.shard_plan): peer ranks of this level's exchange, -1 = none */
} fdfd_level_desc;

int fdfd_direct_create(fdfd_direct** out, int nx, int ny, int tile);
int fdfd_direct_add_level(fdfd_direct* s, const fdfd_level_desc* level);
void fdfd_direct_destroy(fdfd_direct* s);
/* numeric factorisation of the operator's current planes; cached inside the handle and reused by
 * every later solve (the README "save the factorization of A" to-do).                           */
int fdfd_direct_factor(fdfd_direct* s, fdfd_op* op);
int fdfd_direct_stats(fdfd_direct* s, double* factor_bytes, double* factor_flops);
/* x = A^-1 b for nrhs right-hand sides [nrhs][nx*ny], followed by up to max_refine steps of
 * iterative refinement with the fp64 stencil residual; relres = max_j ||b_j - A x_j|| / ||b_j||.
 * Residual contract (the reference's pivoted LU is accurate by construction, linalg.py:139-146; the
 * block factorisation here is not): with max_refine >= 1, a residual still above max(tol, 1e-10)
 * after the refinement steps sends the system to BiCGSTAB preconditioned by the same factors, and
 * the call FAILS (-1) if the limit is missed even then -- it never returns a silently wrong field.
 * refine_steps = refinement steps (+ 1000 + Krylov iterations when that fallback ran).
 * max_refine == 0: raw substitution, residual reported, no guard; < 0: no residual evaluation.   */
int fdfd_direct_solve_host(fdfd_direct* s, fdfd_op* op, const double* b_c128, double* x_c128, int nrhs,
                           int max_refine, double tol, double* relres, int* refine_steps);
int fdfd_direct_solve_dev(fdfd_direct* s, fdfd_op* op, const void* d_b, void* d_x, int nrhs,
                          int max_refine, double tol, double* relres, int* refine_steps);

/* Simulation.solve_fields in one call (simulation.py:113-178): b = scale * src (src real float64 or
 * complex128 on the host), x = A^-1 b with refinement as above, in-plane fields derived from the
 * device-resident x; the three fields come back in one pass (no re-upload of x).                 */
int fdfd_solve_fields_host(fdfd_direct* s, fdfd_op* op, const double* src, int src_is_real, double scale_re,
                           double scale_im, double* x_c128, double* f1_c128, double* f2_c128, int averaging,
                           int max_refine, double tol, double* relres, int* refine_steps);
/* The same with the factorisation inside the call when the handle holds none for the operator's current planes (the
 * whole of Simulation.solve_fields after an eps_r change, simulation.py:80-89 + 113-178): the factorisation is
 * queued first, src crosses PCIe on a second stream while it runs, the singular-pivot flag is read at the call's final
 * synchronisation.  An all-zero src returns zero fields (linalg.py:129-130).  factor_ms (may be NULL): device time of
 * the factorisation, 0 when the cached one was used. */
int fdfd_factor_solve_fields_host(fdfd_direct* s, fdfd_op* op, const double* src, int src_is_real, double scale_re,
                                  double scale_im, double* x_c128, double* f1_c128, double* f2_c128, int averaging,
                                  int max_refine, double tol, double* relres, int* refine_steps, double* factor_ms);
/* flags of the permittivity of the last assembly, evaluated on the device: bit 0 = some entry has an imaginary part,
 * bit 1 = some real part is negative (the argument check of simulation.py:256-265 for arrays too large to scan on the
 * host inside a timed solve) */
int fdfd_op_eps_flags(fdfd_op* op, int* flags);

/* ---- one grid split over several GPUs (no reference counterpart: the reference is single-process).
 * One process per GPU.  Rank 0 makes a 128-byte id, the host program hands it to the other ranks
 * (torch.distributed, MPI, a file ...), every rank then creates its communicator on its current
 * device.  NCCL is loaded at run time (fdfd_comm_load; path may be NULL), never linked.          */
int fdfd_comm_load(const char* libnccl_path);
int fdfd_comm_unique_id(void* id128);
int fdfd_comm_create(fdfd_comm** out, const void* id128, int rank, int world);
void fdfd_comm_destroy(fdfd_comm* c);
/* in-place sum over ranks of `count` doubles on the device (stream of `op`) */
int fdfd_comm_allreduce_sum_dev(fdfd_comm* c, fdfd_op* op, void* d_buf, double count);
/* direct solver on a sharded elimination tree: the levels added must come from ndplan.shard_plan
 * for this rank; subtrees are factorised without communication, the log2(world) levels above them
 * exchange one Schur block / ring vector per level point to point, the solution is summed over
 * ranks (every cell is written by exactly one).  Operator, b and x are replicated on every rank. */
int fdfd_direct_set_comm(fdfd_direct* s, fdfd_comm* c);
/* A front of a level shared by `gsize` ranks, distributed over them by block rows (ndplan.DistFront): plan levels
 * level0 .. level0 + nsteps - 1 are its elimination steps (those levels are added empty, nb = 0, on every rank).
 * Slots are compact: [pivot piece 0 | ... | pivot piece nsteps-1 | ring]; block j covers slots bstart[j]..bstart[j+1]
 * and its rows live on group rank bowner[j]; inv1/inv2[p] = position of slot p in the first/second child's ring or -1.
 * Call after fdfd_direct_set_comm and after every fdfd_direct_add_level, in elimination order.  Every rank of the
 * group takes part in the front's factorisation (owner inverts the pivot block and broadcasts it, every rank forms its
 * rows of G and updates its block rows of the Schur complement after an all-gather of the F_RE panel) and substitution. */
typedef struct {
    int level0, nsteps, gbase, gsize, n, nblk;
    const int* bstart;     /* [nblk + 1] */
    const int* bowner;     /* [nblk] group rank */
    int mc1, mc2;          /* ring sizes of the two children */
    const int* inv1;       /* [n] */
    const int* inv2;       /* [n] */
} fdfd_dist_front_desc;
int fdfd_direct_add_dist_front(fdfd_direct* s, const fdfd_dist_front_desc* d);
/* In-process communicators: `world` ranks inside ONE process (one host thread each; same call surface, transfers are
 * device copies behind a rendezvous).  out[world].  This is how the multi-rank code paths run on a single-GPU box;
 * fdfd_comm_abort wakes every rank blocked in the hub with an error (call it when a rank's thread fails). */
int fdfd_comm_create_local(fdfd_comm** out, int world);
void fdfd_comm_abort(fdfd_comm* c);
/* slab operator: rows [x0, x0 + nxl) of a gnx x ny grid (rows = the slow index of the reference's
 * C-ordered fields; a split along the other axis is the same call on the transposed problem).
 * Every array of a slab operator -- eps_r for fdfd_op_assemble_*, x / y / b of fdfd_op_apply_* and
 * fdfd_krylov_solve_* (precond and c12 must be NULL; see fdfd_slab_set_schwarz) -- has the EXTENDED layout (nxl + 2) x ny:
 * row 0 and row nxl + 1 mirror the neighbouring slabs' boundary rows (periodic in the rank index).
 * eps_r must be given with its halo rows filled; vector halos are exchanged by the library (NCCL
 * send/recv on the operator's stream) before every stencil application, and inner products are
 * summed over ranks on the device.  comm == NULL: a single slab that wraps onto itself.          */
int fdfd_slab_op_create(fdfd_op** out, fdfd_comm* comm, int gnx, int ny, int x0, int nxl, double omega,
                        double dl, int npml_x, int npml_y, int pol, double L0);
/* Restricted additive Schwarz preconditioner for the Krylov solve on slabs (what makes the slab path a SOLVER for
 * grids whose whole-grid factors do not fit: the reference call it serves is still Simulation.solve_fields,
 * simulation.py:113-178 -> linalg.py:123 solver_direct).  One subdomain per rank: the slab's rows plus `overlap`
 * rows of each neighbour plus `npml_sub` rows of artificial PML on both sides, a torus of nxl + 2 (overlap + npml_sub)
 * rows.  fdfd_schwarz_sub_create builds that subdomain's operator (an ordinary fdfd_op: assemble it with the
 * permittivity of global rows x0 - overlap - npml_sub ... periodic, factorise it with fdfd_direct_*);
 * fdfd_slab_set_schwarz attaches operator + factors to the slab operator (NULL, NULL detaches), after which
 * fdfd_krylov_solve_* with method 0 on the slab operator is right-preconditioned by it.  Per application: one
 * exchange of `overlap` rows with each neighbour and one substitution pass with the local factors.  The caller keeps
 * ownership of `sub` and `sub_factors` and must keep them alive while attached. */
int fdfd_schwarz_sub_create(fdfd_op** out, fdfd_op* slab, int overlap, int npml_sub);
int fdfd_slab_set_schwarz(fdfd_op* slab, fdfd_op* sub, fdfd_direct* sub_factors, int overlap, int npml_sub);

/* ---- Krylov solvers on the matrix-free stencil (no reference counterpart: the reference is
 * direct-only; these serve perturbed operators and the slab-decomposed multi-GPU path).
 * method: 0 = BiCGSTAB, 1 = COCG on the symmetrised operator, 2 = restarted GMRES (check_every is
 * then the restart length, <= 1 means 50; no c12).  precond may be NULL; when given (BiCGSTAB and
 * GMRES) its cached factorisation is the right preconditioner -- it may belong to a
 * nearby operator (previous Born iterate, linear part of the Newton Jacobian).  c12 (may be NULL)
 * adds the anti-linear term  c12 .* conj(x)  of the Newton Jacobian (nonlinear_solvers.py:134-135;
 * replaces linalg.py:152 solver_complex2real): the system is then only R-linear and is solved with
 * the real inner product, which real_inner != 0 also forces.  x holds the initial guess on entry. */
int fdfd_krylov_solve_host(fdfd_op* op, fdfd_direct* precond, const double* b_c128, double* x_c128,
                           int method, double tol, int maxiter, int fused, int check_every,
                           const double* c12_c128, int real_inner, int* iters, double* relres,
                           int* converged);
int fdfd_krylov_solve_dev(fdfd_op* op, fdfd_direct* precond, const void* d_b, void* d_x, int method,
                          double tol, int maxiter, int fused, int check_every, const void* d_c12,
                          int real_inner, int* iters, double* relres, int* converged);

/* ---- Born / Newton iteration of the Kerr problem on the device: replaces the loops of nonlinear_solvers.py:13-110
 * (born_solve, newton_solve, nl_eq_and_jac) together with Simulation.compute_nl (simulation.py:57-68).  Every Kerr
 * term of the reference is 3 chi region |E|^2 w(eps_r), so their sum is K |E|^2 with one complex plane K (host,
 * nx*ny).  op_nl: a work operator of the same grid holding the LINEAR eps_r (its eps_nl / planes are overwritten);
 * lin: the direct solver holding the factors of the linear operator (strategy 0, "reuse": they precondition
 * BiCGSTAB on every perturbed / Jacobian system; an exact factorisation through `work` is the fallback);
 * strategy 1 ("refactor"): `work` factorises the perturbed operator every iteration, as the reference does.
 * method 0 = Born, 1 = Newton (R-linear Jacobian solved in the real inner product instead of the reference's real
 * 2N x 2N LU, linalg.py:152-186).  E: start field in, converged field out.  conv[max_iter] receives
 * ||E_new - E_old|| / ||E_new|| per iteration (zero after the last one), *iters the number of iterations done,
 * *inner_iters the Krylov iterations spent.  Only that one scalar per iteration crosses to the host. */
int fdfd_nl_solve_host(fdfd_op* op_nl, fdfd_direct* lin, fdfd_direct* work, const double* K_c128, const double* b_c128,
                       double* E_c128, int method, int strategy, double conv_threshold, int max_iter, double* conv,
                       int* iters, int* inner_iters);

/* complex64 storage for the matrix-free path (interleaved float re, im = numpy complex64): vectors and eps_r
 * stream as 8-byte values (24 B/cell for the fused Ez stencil), arithmetic, inner products and iteration
 * scalars stay fp64.  Parity bar: relative L2 error <= 1e-4 against the fp64 reference.           */
int fdfd_op_apply_host_c64(fdfd_op* op, const float* x_c64, float* y_c64, int fused);
int fdfd_op_apply_dev_c64(fdfd_op* op, const void* d_x, void* d_y, int fused);
int fdfd_krylov_solve_host_c64(fdfd_op* op, const float* b_c64, float* x_c64, int method, double tol, int maxiter,
                               int fused, int check_every, int* iters, double* relres, int* converged);
int fdfd_krylov_solve_dev_c64(fdfd_op* op, const void* d_b, void* d_x, int method, double tol, int maxiter,
                              int fused, int check_every, int* iters, double* relres, int* converged);

/* ---- modal source: replaces source/mode.py:64-108 insert_mode's eigensolve (linalg.py:104
 * solver_eigs -> ARPACK shift-invert).  eps_line: n real relative permittivities along the
 * source plane; averaged != 0 applies the edge average of mode.py:82 (planes normal to y).
 * Returns the `order` eigenpairs of the 1-D waveguide operator closest to
 * (omega sqrt(mu0' eps0') neff)^2: vals[order], vecs[order][n] (unit 2-norm, real).              */
int fdfd_mode_solve_host(const double* eps_line, int n, double omega, double dl, int pol, double L0,
                         double neff, int order, int averaged, double* vals, double* vecs);

/* ---- test hook for the batched complex GEMM that carries the factorisation (zgemm.cuh):
 * C[b] = A[b] op(B[b]) (mode 0) or C[b] -= A[b] op(B[b]) (mode 1); packed row-major batches on the
 * host.  transb: 0 = B is K x N, 1 = B is N x K (C = A B^T).  lower != 0 (M == N): only the 64x64
 * tiles on or below the diagonal are computed, the rest of C is left untouched.                  */
int fdfd_zgemm_batched_host(const double* A, const double* B, double* C, int M, int N, int K, int batch,
                            int mode, int transb, int lower);

/* rows each thread of the fused Ez stencil marches (2, 4, 8; default 4) */
int fdfd_stencil_set_variant(int rows_per_thread, int complex64);
/* fused Hz stencil, A/B switch: chunk_rows 0 (default) = marching kernel with the rows per CTA chosen from the grid size,
 * 4 ... 1024 (multiple of 4) = that many rows per CTA, -4 / -8 = the one-shot kernel with 4 / 8 rows per thread;
 * halo_lanes 1 (default) = 30 stored columns per warp plus one halo lane on each side, 0 = 32 columns per warp with the
 * edge lanes loading their neighbours */
int fdfd_stencil_set_hz_variant(int chunk_rows, int halo_lanes);
/* kernel selection for A/B measurements (bit mask): bit 0 = tiled kernel only (default: persistent kernel for large
 * problems); bit 1 = textbook 4M complex products (default: 3M Karatsuba products, 6 tensor flops per complex MAC) */
int fdfd_zgemm_set_variant(int v);
/* 1 (default): levels of tiny fronts (k <= 32) run as one fused kernel each; 0: generic path everywhere */
int fdfd_direct_set_small_fronts(int enable);
/* device-only timing of one batched GEMM shape (CUDA events, `iters` launches) */
int fdfd_zgemm_bench(int M, int N, int K, int batch, int mode, int transb, int lower, int iters,
                     double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
