// Device-resident Maxwell operator: sc-PML factors, stencil planes, matrix-free apply.
#pragma once
#include "comm.cuh"
#include "common.cuh"

// fdfdpy/constants.py:3-6
#define FDFD_EPS0 8.85418782e-12
#define FDFD_MU0 1.25663706e-6

struct FdfdOp {
    int nx, ny;
    double omega, dl, L0;
    int npml_x, npml_y;
    int pol;            // 0 = Ez, 1 = Hz
    int averaging;      // Hz edge averaging (linalg.py:68-73)
    int has_nl;         // eps_nl present
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;   // fdfd_timer_start/stop
    // 1-D inverse stretch factors 1/s (pml.py:63-76), device
    cplx *isxf, *isxb, *isyf, *isyb;
    // coupling tables of the matrix-free Ez stencil: ax = [axm(nx) | axp(nx)], ay = [aym(ny) | ayp(ny)]
    cplx *ax, *ay;
    cplx32 *ax32, *ay32;   // the same tables in complex64 for the fp32 complex64 stencil
    // permittivity planes (device, nx*ny complex128)
    cplx *eps_r, *eps_nl;
    // five stencil planes c0,cxm,cxp,cym,cyp (device, 5*nx*ny)
    cplx *planes;
    // lazily allocated staging for the *_host entry points (4*nx*ny complex: b, x, f1, f2)
    cplx *io_buf;
    int* d_eps_flag;    // device flag: some eps_r entry has an imaginary part (Hz fused stencil: complex face weights)
    int eps_real;       // host copy: 1 all real, 0 not, -1 not read back yet
    unsigned long long version;   // bumped by every assembly: a factorisation remembers the version it belongs to
    int eps_flags;      // host copy of the device flags (bit 0: imaginary part present, bit 1: negative real part), -1 not read yet
    cudaStream_t up_stream;   // host -> device copies that run next to work already queued on `stream` (lazily created)
    cudaEvent_t ev_up;
    cplx32* eps32;      // complex64 copy of eps_r (| eps_nl) for the complex64 stencil, built on first use
    int eps32_valid;
    // slab of a grid split over several GPUs (halo = 1): nx counts the slab's rows PLUS one halo row on each
    // side, every vector has that extended layout, the stencil only writes rows 1..nx-2 and the halo rows of
    // its input are filled from the neighbouring ranks (periodic in the rank index) before it runs
    int halo, gnx, x0;
    cudaStream_t comm_stream;        // halo exchange runs here, overlapped with the interior rows
    cudaEvent_t ev_in, ev_halo;
    FdfdComm* comm;     // not owned; null with halo = 1 means a single slab wrapping onto itself
    struct SchwarzPre* schwarz;   // slab operators: restricted additive Schwarz preconditioner (krylov.cu), owned
    size_t n() const { return (size_t)nx * ny; }
};

int op_create(FdfdOp** out, int nx, int ny, double omega, double dl, int npml_x, int npml_y,
              int pol, double L0);
// slab operator: rows [x0, x0 + nxl) of a gnx x ny grid
int op_create_slab(FdfdOp** out, FdfdComm* comm, int gnx, int ny, int x0, int nxl, double omega, double dl,
                   int npml_x, int npml_y, int pol, double L0);
// Subdomain operator of a slab for the Schwarz preconditioner: a torus of nxl + 2 (overlap + npml_sub) rows whose
// x stretch factors are the global ones of rows x0 - ext ... (periodic) plus an artificial PML of npml_sub cells at
// both ends.  Assemble it with the permittivity of those rows.
int op_create_schwarz_sub(FdfdOp** out, const FdfdOp* slab, int overlap, int npml_sub);
void op_destroy(FdfdOp* op);
void schwarz_destroy(struct SchwarzPre* s);     // krylov.cu
// flags of the permittivity of the last assembly: bit 0 = an imaginary part is present, bit 1 = a real part is negative
int op_eps_flags(const FdfdOp* op, int* flags);
// fills the two halo rows of an extended-layout vector from the neighbouring slabs
int op_halo_exchange(const FdfdOp* op, void* d_x_ext, size_t elem_bytes, cudaStream_t st);
// eps_r / eps_nl are device pointers (eps_nl may be null); builds the five planes
int op_assemble_dev(FdfdOp* op, const cplx* d_eps_r, const cplx* d_eps_nl, int averaging);
// y = A x using the stored planes (any polarisation, nonlinearity included); nvec vectors back to back
int op_apply_planes(const FdfdOp* op, const cplx* d_x, cplx* d_y, int nvec);
// y = A x recomputing the coefficients from eps and the 1-D PML factors (Ez hot path, 48 B/cell)
int op_apply_fused(const FdfdOp* op, const cplx* d_x, cplx* d_y, int nvec);
// the same three for complex128 (cplx) or complex64 (cplx32) vectors; arithmetic is fp64 either way
template <class V> int op_apply_planes_t(const FdfdOp* op, const V* d_x, V* d_y, int nvec);
template <class V> int op_apply_fused_t(const FdfdOp* op, const V* d_x, V* d_y, int nvec);
template <class V> int op_residual_t(const FdfdOp* op, const V* d_b, const V* d_x, V* d_r, int nvec);
// r = b - A x (planes), returns nothing; used by iterative refinement
int op_residual(const FdfdOp* op, const cplx* d_b, const cplx* d_x, cplx* d_r, int nvec);
// staging buffer of the host entry points (allocated on first use, kept until op_destroy)
int op_io_buffer(FdfdOp* op, cplx** out);
// out[i] = scale * in[i] for n real (in_is_real) or complex inputs; in and out must not overlap
int op_scale_expand(const FdfdOp* op, const void* d_in, int in_is_real, cplx scale, cplx* d_out, size_t n);
// in-plane fields from the solved transverse field (simulation.py:138-176)
int op_derive_fields(const FdfdOp* op, const cplx* d_x, cplx* d_f1, cplx* d_f2, int averaging);
