// Maxwell operator on the device: sc-PML stretch factors (reference pml.py:7-41), the five
// stencil planes of A (linalg.py:39-114), matrix-free application and the derived in-plane
// fields (simulation.py:138-176).  All kernels are plain HBM-bound streaming kernels: the y
// index is the fastest one in memory (derivatives.py:9-11), so threads run along y.
#include <type_traits>
#include "operator.cuh"

thread_local char g_fdfd_err[512] = {0};
unsigned long long g_fdfd_launches = 0;
PhaseTiming g_phase_timing;

// ------------------------------------------------------------------------------------------
// PML: inverse stretch factors for one axis.  Restates create_sfactor's index rules
// (pml.py:29-40): low side i <= npml, high side i > n - npml, 'f' half-cell, 'b' full-cell.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pml_stretch(int i, int n, int npml, double hw, double omega, double L0, cplx& sf, cplx& sb) {
    sf = cmake(1.0, 0.0);
    sb = cmake(1.0, 0.0);
    if (npml >= 1) {
        const double eta0 = sqrt(FDFD_MU0 / FDFD_EPS0);
        double dw = npml * hw;
        double sig_max = -(4 + 1) * (-12.0) / (2 * eta0 * dw);
        double lf = -1.0, lb = -1.0;
        if (i <= npml) {
            lf = hw * (npml - i + 0.5);
            lb = hw * (npml - i + 1);
        } else if (i > n - npml) {
            lf = hw * (i - (n - npml) - 0.5);
            lb = hw * (i - (n - npml) - 1);
        }
        if (lf >= 0.0) {
            double tf = lf / dw, tb = lb / dw;
            double den = omega * FDFD_EPS0 * L0;
            sf = cmake(1.0, -(sig_max * (tf * tf * tf * tf)) / den);
            sb = cmake(1.0, -(sig_max * (tb * tb * tb * tb)) / den);
        }
    }
}
__global__ void pml_axis_kernel(cplx* __restrict__ inv_f, cplx* __restrict__ inv_b, int n, int npml,
                                double hw, double omega, double L0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cplx sf, sb;
    pml_stretch(i, n, npml, hw, omega, L0, sf, sb);
    inv_f[i] = crecip(sf);
    inv_b[i] = crecip(sb);
}

// Stretch factors of a Schwarz subdomain: local row i is global row (x0e + i) mod gnx and keeps that row's factor; the
// first and last npml_s rows add an artificial absorber with the reference's own grading (pml.py:7-18: m = 4,
// ln R = -12) on top of it, so what leaves the overlap region is absorbed instead of wrapping round the local torus.
__global__ void schwarz_sfactor_kernel(cplx* __restrict__ inv_f, cplx* __restrict__ inv_b, int nxe, int gnx, int x0e,
                                       int npml_g, int npml_s, double hw, double omega, double L0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nxe) return;
    const int g = ((x0e + i) % gnx + gnx) % gnx;
    cplx sf, sb;
    pml_stretch(g, gnx, npml_g, hw, omega, L0, sf, sb);
    double df = 0.0, db = 0.0;                       // depth into the artificial layer, in cells
    if (i < npml_s) {
        df = npml_s - i - 0.5;
        db = npml_s - i;
    } else if (i >= nxe - npml_s) {
        df = i - (nxe - npml_s) + 0.5;
        db = i - (nxe - npml_s) + 1.0;
    }
    if (db > 0.0) {
        const double eta0 = sqrt(FDFD_MU0 / FDFD_EPS0);
        const double dw = npml_s * hw, sig_max = 5.0 * 12.0 / (2 * eta0 * dw), den = omega * FDFD_EPS0 * L0;
        const double tf = df / npml_s, tb = db / npml_s;
        sf.y -= sig_max * (tf * tf * tf * tf) / den;
        sb.y -= sig_max * (tb * tb * tb * tb) / den;
    }
    inv_f[i] = crecip(sf);
    inv_b[i] = crecip(sb);
}

// ------------------------------------------------------------------------------------------
// plane assembly: one thread per cell
// ------------------------------------------------------------------------------------------
struct AsmParams {
    int nx, ny, pol, averaging;
    double omega, dx, dy, e0, m0;
};

__device__ __forceinline__ cplx edge_weight(const cplx* __restrict__ eps, int ix, int iy, int nx, int ny,
                                            int axis, const AsmParams& p) {
    // 1 / (eps0' * edge-averaged eps) on the lower face of cell (ix,iy) along `axis` (Hz), or 1/mu0' (Ez)
    if (p.pol == 0) return cmake(1.0 / p.m0, 0.0);
    cplx e = eps[(size_t)ix * ny + iy];
    if (p.averaging) {
        int jx = axis == 0 ? (ix == 0 ? nx - 1 : ix - 1) : ix;
        int jy = axis == 1 ? (iy == 0 ? ny - 1 : iy - 1) : iy;
        cplx e2 = eps[(size_t)jx * ny + jy];
        e = cmake((e2.x + e.x) / 2, (e2.y + e.y) / 2);
    }
    return crecip(cscale(e, p.e0));
}

__global__ void assemble_planes_kernel(cplx* __restrict__ planes, const cplx* __restrict__ eps_r,
                                       const cplx* __restrict__ eps_nl, const cplx* __restrict__ isxf,
                                       const cplx* __restrict__ isxb, const cplx* __restrict__ isyf,
                                       const cplx* __restrict__ isyb, AsmParams p) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)p.nx * p.ny;
    if (idx >= n) return;
    int ix = (int)(idx / p.ny), iy = (int)(idx % p.ny);
    int ixp = ix + 1 == p.nx ? 0 : ix + 1, iyp = iy + 1 == p.ny ? 0 : iy + 1;
    // flux weights on the lower faces: b = S_b^-1 * w / d
    cplx bx0 = cscale(cmul(isxb[ix], edge_weight(eps_r, ix, iy, p.nx, p.ny, 0, p)), 1.0 / p.dx);
    cplx bx1 = cscale(cmul(isxb[ixp], edge_weight(eps_r, ixp, iy, p.nx, p.ny, 0, p)), 1.0 / p.dx);
    cplx by0 = cscale(cmul(isyb[iy], edge_weight(eps_r, ix, iy, p.nx, p.ny, 1, p)), 1.0 / p.dy);
    cplx by1 = cscale(cmul(isyb[iyp], edge_weight(eps_r, ix, iyp, p.nx, p.ny, 1, p)), 1.0 / p.dy);
    cplx cxm = cscale(cmul(isxf[ix], bx0), 1.0 / p.dx);
    cplx cxp = cscale(cmul(isxf[ix], bx1), 1.0 / p.dx);
    cplx cym = cscale(cmul(isyf[iy], by0), 1.0 / p.dy);
    cplx cyp = cscale(cmul(isyf[iy], by1), 1.0 / p.dy);
    cplx shift;
    if (p.pol == 0) shift = cscale(eps_r[idx], p.omega * p.omega * p.e0);
    else shift = cmake(p.omega * p.omega * p.m0, 0.0);
    cplx c0 = csub(csub(shift, cadd(cxm, cxp)), cadd(cym, cyp));
    if (eps_nl) c0 = cadd(c0, cscale(eps_nl[idx], p.omega * p.omega * p.e0));
    planes[idx] = c0;
    planes[n + idx] = cxm;
    planes[2 * n + idx] = cxp;
    planes[3 * n + idx] = cym;
    planes[4 * n + idx] = cyp;
}

// ------------------------------------------------------------------------------------------
// stencil application from stored planes:  y = A x   or   r = b - A x
// ------------------------------------------------------------------------------------------
template <bool RESID, class V>
__global__ void __launch_bounds__(256)
stencil_planes_kernel(const cplx* __restrict__ planes, const V* __restrict__ x, const V* __restrict__ b,
                      V* __restrict__ y, int nx, int ny, int row0, int row1) {
    size_t n = (size_t)nx * ny;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)(row1 - row0) * ny) return;
    idx += (size_t)row0 * ny;                    // rows [row0, row1): a slab never writes its halo rows
    size_t voff = (size_t)blockIdx.y * n;
    const V* xv = x + voff;
    int ix = (int)(idx / ny), iy = (int)(idx % ny);
    size_t xm = (size_t)(ix == 0 ? nx - 1 : ix - 1) * ny + iy;
    size_t xp = (size_t)(ix + 1 == nx ? 0 : ix + 1) * ny + iy;
    size_t ym = (size_t)ix * ny + (iy == 0 ? ny - 1 : iy - 1);
    size_t yp = (size_t)ix * ny + (iy + 1 == ny ? 0 : iy + 1);
    cplx acc = cmul(ldg_c(planes + idx), vload(xv + idx));
    cfma(acc, ldg_c(planes + n + idx), vload(xv + xm));
    cfma(acc, ldg_c(planes + 2 * n + idx), vload(xv + xp));
    cfma(acc, ldg_c(planes + 3 * n + idx), vload(xv + ym));
    cfma(acc, ldg_c(planes + 4 * n + idx), vload(xv + yp));
    if (RESID) acc = csub(vload(b + voff + idx), acc);
    vstore(y + voff + idx, acc);
}

// ------------------------------------------------------------------------------------------
// fused matrix-free Ez stencil: coefficients rebuilt from eps_r (+eps_nl) and 1-D PML products.
// Algorithmic traffic 48 B/cell (x 16 + eps 16 + y 16) (+16 with eps_nl).  Each thread owns one
// y column and marches ROWS consecutive rows keeping the x-neighbours in registers.
// ------------------------------------------------------------------------------------------
int g_fused_rows = 4;          // A/B switch (fdfd_stencil_set_variant): rows marched per thread
int g_hz_chunk = 0;            // Hz: 0 = marching kernel, rows per CTA chosen from the grid size; 4 ... 128 = that many rows;
                               // -4 / -8 = the one-shot kernel with 4 / 8 rows per thread (round-2 baseline, kept for A/B runs)
int g_hz_halo_lanes = 1;       // marching Hz kernel: 30 stored columns per warp + 2 halo lanes (0: edge lanes load their neighbours)
int g_fused_rows32 = 4;        // complex64 vectors: 4 or 8 rows with two columns per thread (float4); 2 = one column per thread

__device__ __forceinline__ cplx shfl_up_c(cplx v) {
    return make_double2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ cplx shfl_down_c(cplx v) {
    return make_double2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}

// ax[ix] = (isxf[ix] isxb[ix], isxf[ix] isxb[ix+1]) / (mu0' dx^2): the two x-coupling coefficients of row ix
// (and the same along y), built once per operator so the stencil only streams x, eps and y.
__global__ void pml_products_kernel(cplx* __restrict__ am, cplx* __restrict__ ap, const cplx* __restrict__ isf,
                                    const cplx* __restrict__ isb, int n, double scale) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ip = i + 1 == n ? 0 : i + 1;
    am[i] = cscale(cmul(isf[i], isb[i]), scale);
    ap[i] = cscale(cmul(isf[i], isb[ip]), scale);
}

// Each thread owns one y column and marches ROWS consecutive rows.  The x-neighbours stay in registers, the
// y-neighbours come from the adjacent lanes by shuffle (only the two edge lanes of a warp load them), the
// coupling coefficients are 1-D tables (warp-uniform loads).  There is no barrier and no early exit in the
// full-CTA body, so every load of a thread is issued before the first one is consumed.
template <int ROWS, class V>
__global__ void __launch_bounds__(128)
stencil_fused_ez_kernel(const V* __restrict__ eps_r, const V* __restrict__ eps_nl,
                        const cplx* __restrict__ axm_t, const cplx* __restrict__ axp_t,
                        const cplx* __restrict__ aym_t, const cplx* __restrict__ ayp_t,
                        const V* __restrict__ x, V* __restrict__ y, int nx, int ny, double w2e0, int row0,
                        int row1) {
    const int ix0 = row0 + blockIdx.y * ROWS;
    const int rows = min(ROWS, row1 - ix0);
    const int iy_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = iy_raw < ny;
    const int iy = active ? iy_raw : ny - 1;            // inactive lanes shadow the last column (shuffles stay full-warp)
    const int lane = threadIdx.x & 31;
    const size_t n = (size_t)nx * ny;
    const size_t voff = (size_t)blockIdx.z * n;
    const V* xv = x + voff;
    const int iym = iy == 0 ? ny - 1 : iy - 1, iyp = iy + 1 == ny ? 0 : iy + 1;
    const bool load_dn = lane == 0 || iy == 0, load_up = lane == 31 || iy_raw >= ny - 1;
    const int ixm = ix0 == 0 ? nx - 1 : ix0 - 1;
    if (rows == ROWS) {
        cplx xc[ROWS + 2], e[ROWS];
        xc[0] = vload(xv + (size_t)ixm * ny + iy);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            xc[r + 1] = vload(xv + (size_t)(ix0 + r) * ny + iy);
            e[r] = vload(eps_r + (size_t)(ix0 + r) * ny + iy);
        }
        {
            int ixl = ix0 + ROWS == nx ? 0 : ix0 + ROWS;
            xc[ROWS + 1] = vload(xv + (size_t)ixl * ny + iy);
        }
        const cplx aym = ldg_c(aym_t + iy), ayp = ldg_c(ayp_t + iy);
        const cplx ay = cadd(aym, ayp);
        if (eps_nl) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) e[r] = cadd(e[r], vload(eps_nl + (size_t)(ix0 + r) * ny + iy));
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            // coefficient tables and the two edge-lane neighbours are cache hits (same lines as adjacent warps)
            const cplx axm = ldg_c(axm_t + ix0 + r), axp = ldg_c(axp_t + ix0 + r);
            cplx xd = shfl_up_c(xc[r + 1]), xu = shfl_down_c(xc[r + 1]);
            if (load_dn) xd = vload(xv + (size_t)(ix0 + r) * ny + iym);
            if (load_up) xu = vload(xv + (size_t)(ix0 + r) * ny + iyp);
            cplx c0 = csub(csub(cscale(e[r], w2e0), cadd(axm, axp)), ay);
            cplx acc = cmul(c0, xc[r + 1]);
            cfma(acc, axm, xc[r]);
            cfma(acc, axp, xc[r + 2]);
            cfma(acc, aym, xd);
            cfma(acc, ayp, xu);
            if (active) vstore(y + voff + (size_t)(ix0 + r) * ny + iy, acc);
        }
        return;
    }
    // ragged last CTA row of the grid
    const cplx aym = ldg_c(aym_t + iy), ayp = ldg_c(ayp_t + iy);
    const cplx ay = cadd(aym, ayp);
    cplx xl = vload(xv + (size_t)ixm * ny + iy);
    cplx xcur = vload(xv + (size_t)ix0 * ny + iy);
    for (int r = 0; r < rows; ++r) {
        int ix = ix0 + r;
        int ixp = ix + 1 == nx ? 0 : ix + 1;
        size_t row = (size_t)ix * ny;
        cplx xr = vload(xv + (size_t)ixp * ny + iy);
        cplx xd = vload(xv + row + iym);
        cplx xu = vload(xv + row + iyp);
        cplx e = vload(eps_r + row + iy);
        if (eps_nl) e = cadd(e, vload(eps_nl + row + iy));
        const cplx axm = ldg_c(axm_t + ix), axp = ldg_c(axp_t + ix);
        cplx c0 = csub(csub(cscale(e, w2e0), cadd(axm, axp)), ay);
        cplx acc = cmul(c0, xcur);
        cfma(acc, axm, xl);
        cfma(acc, axp, xr);
        cfma(acc, aym, xd);
        cfma(acc, ayp, xu);
        if (active) vstore(y + voff + row + iy, acc);
        xl = xcur;
        xcur = xr;
    }
}

// ------------------------------------------------------------------------------------------
// fused matrix-free Hz stencil:  A = Dxf ex^-1 Dxb + Dyf ey^-1 Dyb + w^2 mu  (linalg.py:67-98).
// The face weights 1 / (eps0' * edge-averaged eps) are rebuilt in registers from eps_r itself: the x-faces
// from the rows the thread marches (one reciprocal per row face, shared by the two cells it couples), the
// y-faces from the adjacent lanes by shuffle (each lane inverts its own lower face and borrows the upper one
// from lane + 1).  Same traffic as the Ez kernel: x 16 + eps 16 + y 16 = 48 B/cell against the 112 B/cell of the
// stored planes.  Tables: ax[ix] = (isxf[ix] isxb[ix], isxf[ix] isxb[ix+1]) / (eps0' dx^2), same along y.
// ------------------------------------------------------------------------------------------
// Face weights: W = cplx for a lossy (complex) permittivity, W = double when every eps_r entry is real -- the common
// lossless case, detected once per assembly -- which halves the arithmetic of the reciprocals and fluxes.
__device__ __forceinline__ double fast_rcp64(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));      // ~20 good bits; two Newton steps reach fp64
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return r;
}
template <bool AVG> __device__ __forceinline__ void face_weight(cplx e_lo, cplx e, cplx& w) {
    // reciprocal of the permittivity on the LOWER face of a cell holding e, e_lo = the cell below (linalg.py:68-73)
    const cplx z = AVG ? make_double2(0.5 * (e_lo.x + e.x), 0.5 * (e_lo.y + e.y)) : e;
    const double r = fast_rcp64(z.x * z.x + z.y * z.y);
    w = make_double2(z.x * r, -z.y * r);
}
template <bool AVG> __device__ __forceinline__ void face_weight(cplx e_lo, cplx e, double& w) {
    w = fast_rcp64(AVG ? 0.5 * (e_lo.x + e.x) : e.x);
}
__device__ __forceinline__ cplx wmul(cplx w, cplx d) { return cmul(w, d); }
__device__ __forceinline__ cplx wmul(double w, cplx d) { return make_double2(w * d.x, w * d.y); }

// The operator in flux form:  (A x)_c = sum over the four faces of  a_face * w_face * (x_neighbour - x_c)  + w^2 mu x_c,
// where a_face are the 1-D PML products and w_face = 1 / eps on the face.  The flux  w_face (x_c - x_below)  of a
// face is computed ONCE and used by both cells it couples: along x by the thread that marches the rows, along y by
// handing the lower-face flux of lane + 1 to lane (one shuffle) instead of shuffling weight and neighbour apart.
// (32-bit element offsets: the matrix-free kernels are used for grids below 2^31 cells; the address arithmetic of the
// dozen loads per cell is what this kernel issues most after the fp64 work, so it is kept to one IMAD.WIDE per load.)
template <int ROWS, class V, bool AVG, class W>
__global__ void __launch_bounds__(128, 6)
stencil_fused_hz_kernel(const V* __restrict__ eps_r, const V* __restrict__ eps_nl,
                        const cplx* __restrict__ axm_t, const cplx* __restrict__ axp_t,
                        const cplx* __restrict__ aym_t, const cplx* __restrict__ ayp_t,
                        const V* __restrict__ x, V* __restrict__ y, int nx, int ny, double w2m0, double w2e0,
                        int row0, int row1) {
    const int ix0 = row0 + blockIdx.y * ROWS;
    const int rows = min(ROWS, row1 - ix0);
    const int iy_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = iy_raw < ny;
    const unsigned uny = (unsigned)ny;
    const unsigned iy = active ? (unsigned)iy_raw : uny - 1u;
    const int lane = threadIdx.x & 31;
    const size_t voff = (size_t)blockIdx.z * nx * ny;
    const V* xv = x + voff;
    const unsigned iym = iy == 0 ? uny - 1u : iy - 1u, iyp = iy + 1u == uny ? 0u : iy + 1u;
    const bool load_dn = lane == 0 || iy == 0, load_up = lane == 31 || iy_raw >= ny - 1;
    const cplx aym = ldg_c(aym_t + iy), ayp = ldg_c(ayp_t + iy);
    // element offsets of column iy in rows ix0 - 1 .. ix0 + ROWS (a ragged last CTA re-reads the row after its last one)
    unsigned rb[ROWS + 2];
    rb[0] = (unsigned)(ix0 == 0 ? nx - 1 : ix0 - 1) * uny;
#pragma unroll
    for (int r = 0; r <= ROWS; ++r) {
        int ix = ix0 + (r < rows ? r : rows);
        if (ix >= nx) ix -= nx;
        rb[r + 1] = (unsigned)ix * uny;
    }
    cplx xc[ROWS + 2], e[ROWS + 2];
#pragma unroll
    for (int r = 0; r < ROWS + 2; ++r) {
        xc[r] = vload(xv + (rb[r] + iy));
        e[r] = vload(eps_r + (rb[r] + iy));
    }
    // x-fluxes on the ROWS + 1 row faces this thread touches: fx[r] = w (x_r - x_{r-1}), face below row ix0 + r
    cplx fx[ROWS + 1];
#pragma unroll
    for (int r = 0; r <= ROWS; ++r) {
        W w;
        face_weight<AVG>(e[r], e[r + 1], w);
        fx[r] = wmul(w, csub(xc[r + 1], xc[r]));
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (r < rows) {                                // block-uniform
            const unsigned rowo = rb[r + 1];
            const cplx axm = ldg_c(axm_t + ix0 + r), axp = ldg_c(axp_t + ix0 + r);
            const cplx ec = e[r + 1], xcc = xc[r + 1];
            cplx xd = shfl_up_c(xcc), ed = shfl_up_c(ec);
            if (load_dn) {
                xd = vload(xv + (rowo + iym));
                ed = vload(eps_r + (rowo + iym));
            }
            W wlo;
            face_weight<AVG>(ed, ec, wlo);
            const cplx fy_lo = wmul(wlo, csub(xcc, xd));          // flux through my lower y-face
            cplx fy_hi = shfl_down_c(fy_lo);                      // = lower-face flux of the cell above
            if (load_up) {
                W whi;
                face_weight<AVG>(ec, vload(eps_r + (rowo + iyp)), whi);
                fy_hi = wmul(whi, csub(vload(xv + (rowo + iyp)), xcc));
            }
            // A x = axp fx[r+1] - axm fx[r] + ayp fy_hi - aym fy_lo + w^2 mu x   (+ w^2 eps0 eps_nl x)
            cplx acc = make_double2(w2m0 * xcc.x, w2m0 * xcc.y);
            if (eps_nl) cfma(acc, cscale(vload(eps_nl + (rowo + iy)), w2e0), xcc);
            cfma(acc, axp, fx[r + 1]);
            cfma(acc, cneg(axm), fx[r]);
            cfma(acc, ayp, fy_hi);
            cfma(acc, cneg(aym), fy_lo);
            if (active) vstore(y + voff + (rowo + iy), acc);
        }
    }
}

// Marching variant of the Hz kernel: a CTA walks CHUNK consecutive rows in groups of four and issues the loads of the
// NEXT group (4 rows of x and eps) before the arithmetic of the current one, so every warp keeps 8 x 16 B of loads in
// flight under its own ~600-instruction compute phase instead of relying on other resident warps to cover it (the
// one-shot kernel above: 24 warps/SM, each alternating between a load phase and a long dependent fp64 phase, reached
// 0.68 of the HBM rate).  The x-face flux and the two boundary rows of a group carry over to the next group in
// registers, so a chunk re-reads only 2 rows per CHUNK.  The reciprocal uses one cubic correction step (3 DFMA)
// and the fluxes are subtracted with negated-operand FMAs (no DNEG), 40 fp64 instructions per cell instead of 63.
__device__ __forceinline__ double fast_rcp64c(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));      // >= 20 good bits
    const double e = fma(-d, r, 1.0);                           // 1/d = r (1 + e + e^2 + ...)
    const double t = fma(e, e, e);
    return fma(r, t, r);                                        // relative error e^3 <= 2^-60
}
template <bool AVG> __device__ __forceinline__ void face_weight_c(cplx e_lo, cplx e, cplx& w) {
    const cplx z = AVG ? make_double2(0.5 * (e_lo.x + e.x), 0.5 * (e_lo.y + e.y)) : e;
    const double r = fast_rcp64c(fma(z.x, z.x, z.y * z.y));
    w = make_double2(z.x * r, -z.y * r);
}
template <bool AVG> __device__ __forceinline__ void face_weight_c(cplx e_lo, cplx e, double& w) {
    w = fast_rcp64c(AVG ? fma(0.5, e_lo.x, 0.5 * e.x) : e.x);
}
__device__ __forceinline__ void cfms(cplx& a, cplx b, cplx c) {   // a -= b * c
    a.x = fma(-b.x, c.x, a.x);
    a.x = fma(b.y, c.y, a.x);
    a.y = fma(-b.x, c.y, a.y);
    a.y = fma(-b.y, c.x, a.y);
}
__device__ __forceinline__ double shfl_up_w(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ cplx shfl_up_w(cplx v) { return shfl_up_c(v); }
template <class W> __device__ __forceinline__ W eps_part(cplx e);
template <> __device__ __forceinline__ double eps_part<double>(cplx e) { return e.x; }
template <> __device__ __forceinline__ cplx eps_part<cplx>(cplx e) { return e; }
__device__ __forceinline__ cplx eps_full(double e) { return make_double2(e, 0.0); }
__device__ __forceinline__ cplx eps_full(cplx e) { return e; }

// HALO: a warp covers 30 columns plus one halo column on each side (lanes 0 and 31 load and compute their lower-face
// flux but do not store), so there is no divergent edge-lane code: without it EVERY warp executes the two predicated
// neighbour loads and a third face weight per cell because lanes 0 and 31 exist in every warp.
template <class V, bool AVG, class W, bool HALO>
__global__ void __launch_bounds__(128, 4)
stencil_march_hz_kernel(const V* __restrict__ eps_r, const V* __restrict__ eps_nl,
                        const cplx* __restrict__ axm_t, const cplx* __restrict__ axp_t,
                        const cplx* __restrict__ aym_t, const cplx* __restrict__ ayp_t,
                        const V* __restrict__ x, V* __restrict__ y, int nx, int ny, double w2m0, double w2e0,
                        int row0, int row1, int chunk) {
    const int ix0 = row0 + blockIdx.y * chunk;
    const int nrows = min(chunk, row1 - ix0);
    const int lane = threadIdx.x & 31;
    const unsigned uny = (unsigned)ny;
    int iy_raw;
    bool active;
    unsigned iy;
    if (HALO) {
        iy_raw = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 30 + lane - 1;          // -1 .. : 30 stored columns per warp
        active = lane >= 1 && lane <= 30 && iy_raw < ny;
        int c = iy_raw < 0 ? iy_raw + ny : iy_raw;
        c = c >= ny ? c - ny : c;                                                // the periodic neighbour of column ny - 1
        iy = (unsigned)min(c, ny - 1);                                           // lanes further out: any valid column
    } else {
        iy_raw = blockIdx.x * blockDim.x + threadIdx.x;
        active = iy_raw < ny;
        iy = active ? (unsigned)iy_raw : uny - 1u;
    }
    const size_t voff = (size_t)blockIdx.z * nx * ny;
    const V* xv = x + voff;
    V* yv = y + voff;
    const unsigned iym = iy == 0 ? uny - 1u : iy - 1u, iyp = iy + 1u == uny ? 0u : iy + 1u;
    const bool load_dn = !HALO && (lane == 0 || iy == 0), load_up = !HALO && (lane == 31 || iy_raw >= ny - 1);
    const cplx aym = ldg_c(aym_t + iy), ayp = ldg_c(ayp_t + iy);
    const int last = ix0 + nrows;                        // the row above the chunk (its lower face belongs to the chunk)
    auto rowoff = [&](int ix) -> unsigned {              // rows past `last` are never used: clamp, then wrap
        ix = min(ix, last);
        if (ix < 0) ix += nx;
        if (ix >= nx) ix -= nx;
        return (unsigned)ix * uny;
    };
    // window: rows ix0 - 1 .. ix0 + 4 of x and eps (eps keeps only its real part when W = double)
    cplx xw[6];
    W ew[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const unsigned o = rowoff(ix0 - 1 + r) + iy;
        xw[r] = vload(xv + o);
        ew[r] = eps_part<W>(vload(eps_r + o));
    }
    cplx fx0;                                            // flux through the face below the group's first row
    {
        W w;
        face_weight_c<AVG>(eps_full(ew[0]), eps_full(ew[1]), w);
        fx0 = wmul(w, csub(xw[1], xw[0]));
    }
    for (int g = 0; g < nrows; g += 4) {
        const int ixg = ix0 + g;
        const bool more = g + 4 < nrows;                 // block-uniform
        cplx xn[4], en[4];
        if (more) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned o = rowoff(ixg + 5 + r) + iy;
                xn[r] = vload(xv + o);
                en[r] = vload(eps_r + o);
            }
        }
        cplx fx[5];
        fx[0] = fx0;
#pragma unroll
        for (int r = 1; r <= 4; ++r) {
            W w;
            face_weight_c<AVG>(eps_full(ew[r]), eps_full(ew[r + 1]), w);
            fx[r] = wmul(w, csub(xw[r + 1], xw[r]));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (g + r < nrows) {                         // block-uniform
                const unsigned rowo = (unsigned)(ixg + r) * uny;
                const cplx axm = ldg_c(axm_t + ixg + r), axp = ldg_c(axp_t + ixg + r);
                const cplx xcc = xw[r + 1];
                const W ec = ew[r + 1];
                cplx xd = shfl_up_c(xcc);
                W ed = shfl_up_w(ec);
                if (!HALO && load_dn) {
                    xd = vload(xv + (rowo + iym));
                    ed = eps_part<W>(vload(eps_r + (rowo + iym)));
                }
                W wlo;
                face_weight_c<AVG>(eps_full(ed), eps_full(ec), wlo);
                const cplx fy_lo = wmul(wlo, csub(xcc, xd));
                cplx fy_hi = shfl_down_c(fy_lo);
                if (!HALO && load_up) {
                    W whi;
                    face_weight_c<AVG>(eps_full(ec), vload(eps_r + (rowo + iyp)), whi);
                    fy_hi = wmul(whi, csub(vload(xv + (rowo + iyp)), xcc));
                }
                cplx acc = make_double2(w2m0 * xcc.x, w2m0 * xcc.y);
                if (eps_nl) cfma(acc, cscale(vload(eps_nl + (rowo + iy)), w2e0), xcc);
                cfma(acc, axp, fx[r + 1]);
                cfms(acc, axm, fx[r]);
                cfma(acc, ayp, fy_hi);
                cfms(acc, aym, fy_lo);
                if (active) vstore(yv + (rowo + iy), acc);
            }
        }
        fx0 = fx[4];
        xw[0] = xw[4];
        xw[1] = xw[5];
        ew[0] = ew[4];
        ew[1] = ew[5];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            xw[r + 2] = xn[r];
            ew[r + 2] = eps_part<W>(en[r]);
        }
    }
}

// flag[0] = 1 if some entry of eps (n values) has a non-zero imaginary part
__global__ void eps_imag_kernel(const cplx* __restrict__ eps, size_t n, int* __restrict__ flag) {
    int any = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const cplx e = eps[i];
        any |= (e.y != 0.0 ? 1 : 0) | (e.x < 0.0 ? 2 : 0);
    }
    any = __reduce_or_sync(0xffffffffu, any);
    if (any && (threadIdx.x & 31) == 0) atomicOr(flag, any);
}

// complex64 variant with TWO adjacent y columns per thread: every access is a 16-byte float4 (the same bytes
// in flight per thread as the complex128 kernel), the pair's inner y-neighbours are the thread's own values,
// the outer ones come from the adjacent lanes.  Arithmetic is fp32 here (float coefficient tables): at
// 24 B/cell the fp64 version was bound by the float<->double conversion rate, not by HBM; the rounding is that
// of the complex64 storage itself.  Needs an even ny (16-byte alignment of every row).
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfmaf(float2& a, float2 b, float2 c) {
    a.x = fmaf(b.x, c.x, a.x);
    a.x = fmaf(-b.y, c.y, a.x);
    a.y = fmaf(b.x, c.y, a.y);
    a.y = fmaf(b.y, c.x, a.y);
}
__global__ void narrow_table_kernel(const cplx* __restrict__ in, cplx32* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2((float)in[i].x, (float)in[i].y);
}
template <int ROWS>
__global__ void __launch_bounds__(128)
stencil_fused_ez_c64x2_kernel(const cplx32* __restrict__ eps_r, const cplx32* __restrict__ eps_nl,
                              const cplx32* __restrict__ axm_t, const cplx32* __restrict__ axp_t,
                              const cplx32* __restrict__ aym_t, const cplx32* __restrict__ ayp_t,
                              const cplx32* __restrict__ x, cplx32* __restrict__ y, int nx, int ny, float w2e0,
                              int row0, int row1) {
    const int ix0 = row0 + blockIdx.y * ROWS;
    const int rows = min(ROWS, row1 - ix0);
    const int pair_raw = blockIdx.x * blockDim.x + threadIdx.x;          // column pair index
    const int npair = ny >> 1;
    const bool active = pair_raw < npair;
    const int pr = active ? pair_raw : npair - 1;
    const int iy = 2 * pr;                                               // columns iy, iy + 1
    const int lane = threadIdx.x & 31;
    const size_t n = (size_t)nx * ny;
    const size_t voff = (size_t)blockIdx.z * n;
    const cplx32* xv = x + voff;
    const int iym = iy == 0 ? ny - 1 : iy - 1, iyp = iy + 2 == ny ? 0 : iy + 2;
    const bool load_dn = lane == 0 || iy == 0, load_up = lane == 31 || pair_raw >= npair - 1;
    const int ixm = ix0 == 0 ? nx - 1 : ix0 - 1;
    const float4 aym = __ldg(reinterpret_cast<const float4*>(aym_t + iy)), ayp = __ldg(reinterpret_cast<const float4*>(ayp_t + iy));
    const float2 aym0 = make_float2(aym.x, aym.y), aym1 = make_float2(aym.z, aym.w);
    const float2 ayp0 = make_float2(ayp.x, ayp.y), ayp1 = make_float2(ayp.z, ayp.w);
    auto ld4 = [&](const cplx32* base, int ix) { return __ldg(reinterpret_cast<const float4*>(base + (size_t)ix * ny + iy)); };
    float4 xc[ROWS + 2], e[ROWS];
    xc[0] = ld4(xv, ixm);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int ix = min(ix0 + r, row1 - 1);                           // ragged last CTA: re-read the last row
        xc[r + 1] = ld4(xv, ix);
        e[r] = ld4(eps_r, ix);
        if (eps_nl) { float4 t = ld4(eps_nl, ix); e[r].x += t.x; e[r].y += t.y; e[r].z += t.z; e[r].w += t.w; }
    }
    {
        int ixl = ix0 + rows == nx ? 0 : ix0 + rows;
        float4 nxt = ld4(xv, ixl);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) if (r + 1 == rows) xc[r + 2] = nxt;       // the row after the last computed one
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (r < rows) {
            const int ix = ix0 + r;
            const float2 axm = __ldg(axm_t + ix), axp = __ldg(axp_t + ix);
            const float2 c_lo = make_float2(xc[r + 1].x, xc[r + 1].y), c_hi = make_float2(xc[r + 1].z, xc[r + 1].w);
            // outer neighbours: previous lane's upper column / next lane's lower column
            float2 dn = make_float2(__shfl_up_sync(0xffffffffu, c_hi.x, 1), __shfl_up_sync(0xffffffffu, c_hi.y, 1));
            float2 up = make_float2(__shfl_down_sync(0xffffffffu, c_lo.x, 1), __shfl_down_sync(0xffffffffu, c_lo.y, 1));
            if (load_dn) dn = __ldg(xv + (size_t)ix * ny + iym);
            if (load_up) up = __ldg(xv + (size_t)ix * ny + iyp);
            const float sx = axm.x + axp.x, sy = axm.y + axp.y;
            float2 c0 = make_float2(e[r].x * w2e0 - sx - (aym0.x + ayp0.x), e[r].y * w2e0 - sy - (aym0.y + ayp0.y));
            float2 acc0 = cmulf(c0, c_lo);
            cfmaf(acc0, axm, make_float2(xc[r].x, xc[r].y));
            cfmaf(acc0, axp, make_float2(xc[r + 2].x, xc[r + 2].y));
            cfmaf(acc0, aym0, dn);
            cfmaf(acc0, ayp0, c_hi);
            float2 c1 = make_float2(e[r].z * w2e0 - sx - (aym1.x + ayp1.x), e[r].w * w2e0 - sy - (aym1.y + ayp1.y));
            float2 acc1 = cmulf(c1, c_hi);
            cfmaf(acc1, axm, make_float2(xc[r].z, xc[r].w));
            cfmaf(acc1, axp, make_float2(xc[r + 2].z, xc[r + 2].w));
            cfmaf(acc1, aym1, c_lo);
            cfmaf(acc1, ayp1, up);
            if (active)
                *reinterpret_cast<float4*>(y + voff + (size_t)ix * ny + iy) = make_float4(acc0.x, acc0.y, acc1.x, acc1.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// derived in-plane fields
// ------------------------------------------------------------------------------------------
__global__ void derive_fields_kernel(const cplx* __restrict__ X, const cplx* __restrict__ eps_r,
                                     const cplx* __restrict__ eps_nl, const cplx* __restrict__ isxb,
                                     const cplx* __restrict__ isyb, cplx* __restrict__ f1,
                                     cplx* __restrict__ f2, AsmParams p) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)p.nx * p.ny;
    if (idx >= n) return;
    int ix = (int)(idx / p.ny), iy = (int)(idx % p.ny);
    int ixm = ix == 0 ? p.nx - 1 : ix - 1, iym = iy == 0 ? p.ny - 1 : iy - 1;
    size_t jx = (size_t)ixm * p.ny + iy, jy = (size_t)ix * p.ny + iym;
    cplx xc = X[idx];
    cplx dxb = cscale(cmul(isxb[ix], csub(xc, X[jx])), 1.0 / p.dx);
    cplx dyb = cscale(cmul(isyb[iy], csub(xc, X[jy])), 1.0 / p.dy);
    if (p.pol == 0) {
        // Hx = -1/(i w mu0') Dyb Ez = (i/(w mu0')) Dyb Ez ;  Hy = 1/(i w mu0') Dxb Ez
        double c = 1.0 / p.omega / p.m0;
        f1[idx] = cmake(-c * dyb.y, c * dyb.x);
        f2[idx] = cmake(c * dxb.y, -c * dxb.x);
    } else {
        cplx e = eps_r[idx], ex = eps_r[jx], ey = eps_r[jy];
        if (eps_nl) {
            e = cadd(e, eps_nl[idx]);
            ex = cadd(ex, eps_nl[jx]);
            ey = cadd(ey, eps_nl[jy]);
        }
        cplx wx = e, wy = e;
        if (p.averaging) {
            wx = cmake((ex.x + e.x) / 2, (ex.y + e.y) / 2);
            wy = cmake((ey.x + e.x) / 2, (ey.y + e.y) / 2);
        }
        // Ex = 1/(i w) Dyb Hz / ey_w ;  Ey = -1/(i w) Dxb Hz / ex_w
        cplx a = cdiv(dyb, cscale(wy, p.e0));
        cplx b = cdiv(dxb, cscale(wx, p.e0));
        double c = 1.0 / p.omega;
        f1[idx] = cmake(c * a.y, -c * a.x);
        f2[idx] = cmake(-c * b.y, c * b.x);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static AsmParams make_params(const FdfdOp* op) {
    AsmParams p;
    p.nx = op->nx;
    p.ny = op->ny;
    p.pol = op->pol;
    p.averaging = op->averaging;
    p.omega = op->omega;
    // grid spacing exactly as the reference derives it: L / N with L = N*dl (simulation.py:33-34, linalg.py:23-32)
    p.dx = ((double)op->nx * op->dl) / op->nx;
    p.dy = ((double)op->ny * op->dl) / op->ny;
    p.e0 = FDFD_EPS0 * op->L0;
    p.m0 = FDFD_MU0 * op->L0;
    return p;
}

// ext[j] = glob[(x0 - 1 + j) mod gnx]: the slab's slice of a per-row array, halo rows included
__global__ void slab_slice_kernel(cplx* __restrict__ ext, const cplx* __restrict__ glob, int next, int gnx, int x0) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= next) return;
    ext[j] = glob[((x0 - 1 + j) % gnx + gnx) % gnx];
}

static int op_create_impl(FdfdOp** out, int nx, int ny, double omega, double dl, int npml_x, int npml_y, int pol,
                          double L0, int halo, int gnx, int x0, FdfdComm* comm) {
    if (nx < 2 || ny < 2) FDFD_FAIL("grid must be at least 2x2, got %dx%d", nx, ny);
    if (pol != 0 && pol != 1) FDFD_FAIL("pol must be 0 (Ez) or 1 (Hz)");
    FdfdOp* op = new FdfdOp();
    memset(op, 0, sizeof(*op));
    op->nx = nx; op->ny = ny; op->omega = omega; op->dl = dl; op->L0 = L0;
    op->npml_x = npml_x; op->npml_y = npml_y; op->pol = pol; op->averaging = 1;
    op->halo = halo; op->gnx = gnx; op->x0 = x0; op->comm = comm;
    FDFD_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
    size_t n = op->n();
    FDFD_CHECK(cudaMalloc(&op->isxf, sizeof(cplx) * nx));
    FDFD_CHECK(cudaMalloc(&op->isxb, sizeof(cplx) * nx));
    FDFD_CHECK(cudaMalloc(&op->isyf, sizeof(cplx) * ny));
    FDFD_CHECK(cudaMalloc(&op->isyb, sizeof(cplx) * ny));
    FDFD_CHECK(cudaMalloc(&op->eps_r, sizeof(cplx) * n));
    FDFD_CHECK(cudaMalloc(&op->eps_nl, sizeof(cplx) * n));
    FDFD_CHECK(cudaMalloc(&op->planes, sizeof(cplx) * n * 5));
    FDFD_CHECK(cudaMalloc(&op->d_eps_flag, sizeof(int)));
    op->eps_real = -1;
    AsmParams p = make_params(op);
    if (!halo) {
        { pml_axis_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->isxf, op->isxb, nx, npml_x, p.dx, omega, L0); ++g_fdfd_launches; }
    } else {
        // the stretch factors of the WHOLE axis, then this slab's rows (and its neighbours' boundary rows)
        cplx *gf = nullptr, *gb = nullptr;
        FDFD_CHECK(cudaMalloc(&gf, sizeof(cplx) * gnx));
        FDFD_CHECK(cudaMalloc(&gb, sizeof(cplx) * gnx));
        { pml_axis_kernel<<<ceil_div(gnx, 128), 128, 0, op->stream>>>(gf, gb, gnx, npml_x, p.dx, omega, L0); ++g_fdfd_launches; }
        { slab_slice_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->isxf, gf, nx, gnx, x0); ++g_fdfd_launches; }
        { slab_slice_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->isxb, gb, nx, gnx, x0); ++g_fdfd_launches; }
        FDFD_CHECK(cudaStreamSynchronize(op->stream));
        cudaFree(gf); cudaFree(gb);
    }
    { pml_axis_kernel<<<ceil_div(ny, 128), 128, 0, op->stream>>>(op->isyf, op->isyb, ny, npml_y, p.dy, omega, L0); ++g_fdfd_launches; }
    if (halo && comm && comm->world > 1) {
        FDFD_CHECK(cudaStreamCreateWithFlags(&op->comm_stream, cudaStreamNonBlocking));
        FDFD_CHECK(cudaEventCreateWithFlags(&op->ev_in, cudaEventDisableTiming));
        FDFD_CHECK(cudaEventCreateWithFlags(&op->ev_halo, cudaEventDisableTiming));
    }
    // coupling-coefficient tables of the matrix-free Ez stencil
    FDFD_CHECK(cudaMalloc(&op->ax, sizeof(cplx) * 2 * nx));
    FDFD_CHECK(cudaMalloc(&op->ay, sizeof(cplx) * 2 * ny));
    // (Ez: divided by mu0'; Hz: by eps0', the permittivity itself is applied per face inside the kernel)
    const double med = pol == 0 ? p.m0 : p.e0;
    { pml_products_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->ax, op->ax + nx, op->isxf, op->isxb, nx, 1.0 / (med * p.dx * p.dx)); ++g_fdfd_launches; }
    { pml_products_kernel<<<ceil_div(ny, 128), 128, 0, op->stream>>>(op->ay, op->ay + ny, op->isyf, op->isyb, ny, 1.0 / (med * p.dy * p.dy)); ++g_fdfd_launches; }
    FDFD_CHECK(cudaMalloc(&op->ax32, sizeof(cplx32) * 2 * nx));
    FDFD_CHECK(cudaMalloc(&op->ay32, sizeof(cplx32) * 2 * ny));
    { narrow_table_kernel<<<ceil_div(2 * nx, 128), 128, 0, op->stream>>>(op->ax, op->ax32, 2 * nx); ++g_fdfd_launches; }
    { narrow_table_kernel<<<ceil_div(2 * ny, 128), 128, 0, op->stream>>>(op->ay, op->ay32, 2 * ny); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    *out = op;
    return 0;
}

int op_create(FdfdOp** out, int nx, int ny, double omega, double dl, int npml_x, int npml_y, int pol, double L0) {
    return op_create_impl(out, nx, ny, omega, dl, npml_x, npml_y, pol, L0, 0, nx, 0, nullptr);
}

int op_create_slab(FdfdOp** out, FdfdComm* comm, int gnx, int ny, int x0, int nxl, double omega, double dl,
                   int npml_x, int npml_y, int pol, double L0) {
    if (nxl < 1 || x0 < 0 || x0 + nxl > gnx) FDFD_FAIL("slab rows [%d, %d) outside the %d-row grid", x0, x0 + nxl, gnx);
    return op_create_impl(out, nxl + 2, ny, omega, dl, npml_x, npml_y, pol, L0, 1, gnx, x0, comm);
}

int op_create_schwarz_sub(FdfdOp** out, const FdfdOp* slab, int overlap, int npml_sub) {
    if (!slab->halo) FDFD_FAIL("the Schwarz subdomain belongs to a slab operator");
    const int nxl = slab->nx - 2, ext = overlap + npml_sub;
    if (overlap < 0 || npml_sub < 1) FDFD_FAIL("Schwarz subdomain: overlap >= 0 and npml_sub >= 1");
    if (overlap > nxl) FDFD_FAIL("Schwarz overlap (%d rows) exceeds the slab (%d rows)", overlap, nxl);
    FdfdOp* op = nullptr;
    // created like a whole-grid operator (own stream, y factors, tables), then the x factors are replaced
    if (op_create_impl(&op, nxl + 2 * ext, slab->ny, slab->omega, slab->dl, 0, slab->npml_y, slab->pol, slab->L0, 0,
                       nxl + 2 * ext, 0, nullptr))
        return -1;
    AsmParams p = make_params(op);
    const int nx = op->nx;
    { schwarz_sfactor_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->isxf, op->isxb, nx, slab->gnx, slab->x0 - ext,
                                                                        slab->npml_x, npml_sub, p.dx, op->omega, op->L0); ++g_fdfd_launches; }
    const double med = op->pol == 0 ? p.m0 : p.e0;
    { pml_products_kernel<<<ceil_div(nx, 128), 128, 0, op->stream>>>(op->ax, op->ax + nx, op->isxf, op->isxb, nx, 1.0 / (med * p.dx * p.dx)); ++g_fdfd_launches; }
    { narrow_table_kernel<<<ceil_div(2 * nx, 128), 128, 0, op->stream>>>(op->ax, op->ax32, 2 * nx); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    *out = op;
    return 0;
}

int op_halo_exchange(const FdfdOp* op, void* xv, size_t elem, cudaStream_t st) {
    if (!op->halo) return 0;
    const size_t row = elem * op->ny;                 // bytes per row (elem = 16: complex128, 8: complex64)
    char* x = static_cast<char*>(xv);
    char *first = x + row, *last = x + (size_t)(op->nx - 2) * row, *halo_lo = x, *halo_hi = x + (size_t)(op->nx - 1) * row;
    if (!op->comm || op->comm->world == 1) {         // one slab: the grid wraps onto itself
        FDFD_CHECK(cudaMemcpyAsync(halo_lo, last, row, cudaMemcpyDeviceToDevice, st));
        FDFD_CHECK(cudaMemcpyAsync(halo_hi, first, row, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const int w = op->comm->world, r = op->comm->rank, lower = (r + w - 1) % w, upper = (r + 1) % w;
    // my first row is the upper halo of the rank below, my last row the lower halo of the rank above
    return comm_halo_exchange(op->comm, first, last, halo_lo, halo_hi, lower, upper, row / 8, st);
}

void op_destroy(FdfdOp* op) {
    if (!op) return;
    if (op->schwarz) schwarz_destroy(op->schwarz);
    cudaFree(op->isxf); cudaFree(op->isxb); cudaFree(op->isyf); cudaFree(op->isyb);
    cudaFree(op->eps_r); cudaFree(op->eps_nl); cudaFree(op->planes); cudaFree(op->d_eps_flag);
    if (op->io_buf) cudaFree(op->io_buf);
    cudaFree(op->ax); cudaFree(op->ay); cudaFree(op->ax32); cudaFree(op->ay32);
    if (op->eps32) cudaFree(op->eps32);
    if (op->comm_stream) { cudaStreamDestroy(op->comm_stream); cudaEventDestroy(op->ev_in); cudaEventDestroy(op->ev_halo); }
    if (op->ev0) { cudaEventDestroy(op->ev0); cudaEventDestroy(op->ev1); }
    if (op->up_stream) { cudaStreamDestroy(op->up_stream); cudaEventDestroy(op->ev_up); }
    cudaStreamDestroy(op->stream);
    delete op;
}

int op_io_buffer(FdfdOp* op, cplx** out) {
    if (!op->io_buf) FDFD_CHECK(cudaMalloc(&op->io_buf, sizeof(cplx) * 4 * op->n()));
    *out = op->io_buf;
    return 0;
}

template <bool REAL>
__global__ void scale_expand_kernel(const void* __restrict__ in, cplx scale, cplx* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (REAL) {
        double v = static_cast<const double*>(in)[i];
        out[i] = make_double2(v * scale.x, v * scale.y);
    } else {
        out[i] = cmul(static_cast<const cplx*>(in)[i], scale);
    }
}

int op_scale_expand(const FdfdOp* op, const void* d_in, int in_is_real, cplx scale, cplx* d_out, size_t n) {
    if (in_is_real) scale_expand_kernel<true><<<ceil_div(n, 256), 256, 0, op->stream>>>(d_in, scale, d_out, n);
    else scale_expand_kernel<false><<<ceil_div(n, 256), 256, 0, op->stream>>>(d_in, scale, d_out, n);
    ++g_fdfd_launches;
    FDFD_CHECK(cudaGetLastError());
    return 0;
}

int op_assemble_dev(FdfdOp* op, const cplx* d_eps_r, const cplx* d_eps_nl, int averaging) {
    size_t n = op->n();
    op->averaging = averaging;
    op->has_nl = d_eps_nl != nullptr;
    op->eps32_valid = 0;
    ++op->version;
    if (d_eps_r != op->eps_r)
        FDFD_CHECK(cudaMemcpyAsync(op->eps_r, d_eps_r, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, op->stream));
    if (d_eps_nl && d_eps_nl != op->eps_nl)
        FDFD_CHECK(cudaMemcpyAsync(op->eps_nl, d_eps_nl, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, op->stream));
    AsmParams p = make_params(op);
    // one pass over eps_r sets two flags on the device: bit 0 = some entry has an imaginary part (lossless permittivity
    // gets the real-weight Hz stencil), bit 1 = some real part is negative (Simulation's argument check for large
    // arrays reads it here instead of scanning 134 MB on the host: 17 ms at 4096^2)
    FDFD_CHECK(cudaMemsetAsync(op->d_eps_flag, 0, sizeof(int), op->stream));
    { eps_imag_kernel<<<592, 256, 0, op->stream>>>(op->eps_r, n, op->d_eps_flag); ++g_fdfd_launches; }
    op->eps_real = -1;
    op->eps_flags = -1;
    { assemble_planes_kernel<<<ceil_div(n, 256), 256, 0, op->stream>>>(
        op->planes, op->eps_r, op->has_nl ? op->eps_nl : nullptr, op->isxf, op->isxb, op->isyf, op->isyb, p); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}

// Is every eps_r entry real?  Evaluated on the device after an assembly, read back lazily (one 4-byte copy) by the
// first fused Hz application that needs to choose between the real- and the complex-weight kernel.
static int op_eps_is_real(const FdfdOp* cop, int* out) {
    FdfdOp* op = const_cast<FdfdOp*>(cop);
    if (op->eps_real < 0) {
        int h = 1;
        FDFD_CHECK(cudaMemcpyAsync(&h, op->d_eps_flag, sizeof(int), cudaMemcpyDeviceToHost, op->stream));
        FDFD_CHECK(cudaStreamSynchronize(op->stream));
        op->eps_flags = h;
        op->eps_real = (h & 1) ? 0 : 1;
    }
    *out = op->eps_real;
    return 0;
}
int op_eps_flags(const FdfdOp* op, int* flags) {
    int real_eps = 0;
    if (op_eps_is_real(op, &real_eps)) return -1;
    *flags = op->eps_flags;
    return 0;
}

// Row ranges one stencil application is launched over.  Whole grid: one range.  Slab on several ranks: the
// halo exchange runs on the operator's communication stream WHILE the rows that do not touch a halo are
// computed; the first and last owned row follow once the halos have landed.
struct RowPlan {
    int nranges;
    int r0[3], r1[3];
    bool wait_halo_before[3];
};

template <class V>
static int slab_begin(const FdfdOp* op, const V* d_x, int nvec, RowPlan* plan) {
    plan->nranges = 1;
    plan->r0[0] = op->halo; plan->r1[0] = op->nx - op->halo; plan->wait_halo_before[0] = false;
    if (!op->halo) return 0;
    if (nvec != 1) FDFD_FAIL("slab operators apply one vector at a time");
    V* x = const_cast<V*>(d_x);
    const bool overlap = op->comm && op->comm->world > 1 && op->comm_stream && op->nx - 2 >= 3;
    if (!overlap) return op_halo_exchange(op, x, sizeof(V), op->stream);
    FDFD_CHECK(cudaEventRecord(op->ev_in, op->stream));
    FDFD_CHECK(cudaStreamWaitEvent(op->comm_stream, op->ev_in, 0));
    if (op_halo_exchange(op, x, sizeof(V), op->comm_stream)) return -1;
    FDFD_CHECK(cudaEventRecord(op->ev_halo, op->comm_stream));
    plan->nranges = 3;
    plan->r0[0] = 2; plan->r1[0] = op->nx - 2; plan->wait_halo_before[0] = false;
    plan->r0[1] = 1; plan->r1[1] = 2; plan->wait_halo_before[1] = true;
    plan->r0[2] = op->nx - 2; plan->r1[2] = op->nx - 1; plan->wait_halo_before[2] = false;
    return 0;
}

// complex64 copy of eps_r for the complex64 stencil (24 B/cell); rebuilt lazily after every assembly
__global__ void narrow_kernel(const cplx* __restrict__ in, cplx32* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2((float)in[i].x, (float)in[i].y);
}
static int eps32(const FdfdOp* cop, const cplx32** er, const cplx32** enl) {
    FdfdOp* op = const_cast<FdfdOp*>(cop);
    const size_t n = op->n();
    if (!op->eps32) FDFD_CHECK(cudaMalloc(&op->eps32, sizeof(cplx32) * 2 * n));
    if (!op->eps32_valid) {
        narrow_kernel<<<ceil_div(n, 256), 256, 0, op->stream>>>(op->eps_r, op->eps32, n);
        ++g_fdfd_launches;
        if (op->has_nl) { narrow_kernel<<<ceil_div(n, 256), 256, 0, op->stream>>>(op->eps_nl, op->eps32 + n, n); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
        op->eps32_valid = 1;
    }
    *er = op->eps32;
    *enl = op->has_nl ? op->eps32 + n : nullptr;
    return 0;
}
static int eps_of(const FdfdOp* op, const cplx** er, const cplx** enl) {
    *er = op->eps_r;
    *enl = op->has_nl ? op->eps_nl : nullptr;
    return 0;
}
static int eps_of(const FdfdOp* op, const cplx32** er, const cplx32** enl) { return eps32(op, er, enl); }

template <bool RESID, class V>
static int launch_planes(const FdfdOp* op, const V* d_x, const V* d_b, V* d_y, int nvec) {
    RowPlan plan;
    if (slab_begin(op, d_x, nvec, &plan)) return -1;
    for (int i = 0; i < plan.nranges; ++i) {
        if (plan.wait_halo_before[i]) FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
        dim3 grid(ceil_div((size_t)(plan.r1[i] - plan.r0[i]) * op->ny, 256), nvec);
        stencil_planes_kernel<RESID, V><<<grid, 256, 0, op->stream>>>(op->planes, d_x, d_b, d_y, op->nx, op->ny,
                                                                     plan.r0[i], plan.r1[i]);
        ++g_fdfd_launches;
    }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}

template <class V>
int op_apply_planes_t(const FdfdOp* op, const V* d_x, V* d_y, int nvec) {
    return launch_planes<false, V>(op, d_x, (const V*)nullptr, d_y, nvec);
}
template <class V>
int op_residual_t(const FdfdOp* op, const V* d_b, const V* d_x, V* d_r, int nvec) {
    return launch_planes<true, V>(op, d_x, d_b, d_r, nvec);
}
template <class V>
int op_apply_fused_t(const FdfdOp* op, const V* d_x, V* d_y, int nvec) {
    const V *er, *enl;
    if (eps_of(op, &er, &enl)) return -1;
    RowPlan plan;
    if (slab_begin(op, d_x, nvec, &plan)) return -1;
    AsmParams p = make_params(op);
    if (op->pol != 0 && op->n() >= (1ull << 31)) return op_apply_planes_t<V>(op, d_x, d_y, nvec);   // 32-bit offsets
    if (op->pol != 0) {
        const int rows = g_hz_chunk == -8 ? 8 : 4;
        int real_eps = 0;
        if (op_eps_is_real(op, &real_eps)) return -1;
        void (*kern)(const V*, const V*, const cplx*, const cplx*, const cplx*, const cplx*, const V*, V*, int, int, double,
                     double, int, int);
        if (real_eps) {
            if (op->averaging) kern = rows == 8 ? stencil_fused_hz_kernel<8, V, true, double> : stencil_fused_hz_kernel<4, V, true, double>;
            else kern = rows == 8 ? stencil_fused_hz_kernel<8, V, false, double> : stencil_fused_hz_kernel<4, V, false, double>;
        } else {
            if (op->averaging) kern = rows == 8 ? stencil_fused_hz_kernel<8, V, true, cplx> : stencil_fused_hz_kernel<4, V, true, cplx>;
            else kern = rows == 8 ? stencil_fused_hz_kernel<8, V, false, cplx> : stencil_fused_hz_kernel<4, V, false, cplx>;
        }
        if (g_hz_chunk >= 0) {
            // marching kernel.  Rows per CTA: long chunks amortise the two re-read rows and keep the prefetch pipeline
            // full, but the grid must still fill the machine (4 CTAs of 128 threads per SM x 148 SMs) about twice over
            const bool halo = g_hz_halo_lanes != 0;
            int chunk = g_hz_chunk;
            if (chunk == 0) {
                const long long cols = ceil_div(op->ny, halo ? 120 : 128), nrow = op->nx - 2 * op->halo;
                chunk = 4;
                for (int c = 64; c >= 8; c >>= 1)
                    if (cols * ((nrow + c - 1) / c) * nvec >= 2 * 4 * 148) { chunk = c; break; }
            }
            void (*mk)(const V*, const V*, const cplx*, const cplx*, const cplx*, const cplx*, const V*, V*, int, int,
                       double, double, int, int, int);
            if (real_eps) mk = op->averaging ? (halo ? stencil_march_hz_kernel<V, true, double, true> : stencil_march_hz_kernel<V, true, double, false>)
                                             : (halo ? stencil_march_hz_kernel<V, false, double, true> : stencil_march_hz_kernel<V, false, double, false>);
            else mk = op->averaging ? (halo ? stencil_march_hz_kernel<V, true, cplx, true> : stencil_march_hz_kernel<V, true, cplx, false>)
                                    : (halo ? stencil_march_hz_kernel<V, false, cplx, true> : stencil_march_hz_kernel<V, false, cplx, false>);
            for (int i = 0; i < plan.nranges; ++i) {
                if (plan.wait_halo_before[i]) FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
                dim3 grid(ceil_div(op->ny, halo ? 120 : 128), ceil_div(plan.r1[i] - plan.r0[i], chunk), nvec);
                mk<<<grid, 128, 0, op->stream>>>(er, enl, op->ax, op->ax + op->nx, op->ay, op->ay + op->ny, d_x, d_y, op->nx,
                                                 op->ny, p.omega * p.omega * p.m0, p.omega * p.omega * p.e0, plan.r0[i],
                                                 plan.r1[i], chunk);
                ++g_fdfd_launches;
            }
            FDFD_CHECK(cudaGetLastError());
            return 0;
        }
        for (int i = 0; i < plan.nranges; ++i) {
            if (plan.wait_halo_before[i]) FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
            dim3 grid(ceil_div(op->ny, 128), ceil_div(plan.r1[i] - plan.r0[i], rows), nvec);
            kern<<<grid, 128, 0, op->stream>>>(er, enl, op->ax, op->ax + op->nx, op->ay, op->ay + op->ny, d_x, d_y, op->nx,
                                               op->ny, p.omega * p.omega * p.m0, p.omega * p.omega * p.e0, plan.r0[i],
                                               plan.r1[i]);
            ++g_fdfd_launches;
        }
        FDFD_CHECK(cudaGetLastError());
        return 0;
    }
    if constexpr (std::is_same<V, cplx32>::value) {
        if ((op->ny & 1) == 0 && g_fused_rows32 != 2) {              // two columns per thread, float4 accesses
            const int rows32 = g_fused_rows32 == 8 ? 8 : 4;
            auto k2 = rows32 == 8 ? stencil_fused_ez_c64x2_kernel<8> : stencil_fused_ez_c64x2_kernel<4>;
            for (int i = 0; i < plan.nranges; ++i) {
                if (plan.wait_halo_before[i]) FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
                dim3 grid(ceil_div(op->ny / 2, 128), ceil_div(plan.r1[i] - plan.r0[i], rows32), nvec);
                k2<<<grid, 128, 0, op->stream>>>(er, enl, op->ax32, op->ax32 + op->nx, op->ay32, op->ay32 + op->ny, d_x, d_y,
                                                 op->nx, op->ny, (float)(p.omega * p.omega * p.e0), plan.r0[i], plan.r1[i]);
                ++g_fdfd_launches;
            }
            FDFD_CHECK(cudaGetLastError());
            return 0;
        }
    }
    const int rows = sizeof(V) == sizeof(cplx32) ? g_fused_rows32 : g_fused_rows;
    auto kern = rows == 8 ? stencil_fused_ez_kernel<8, V> : rows == 2 ? stencil_fused_ez_kernel<2, V> : stencil_fused_ez_kernel<4, V>;
    for (int i = 0; i < plan.nranges; ++i) {
        if (plan.wait_halo_before[i]) FDFD_CHECK(cudaStreamWaitEvent(op->stream, op->ev_halo, 0));
        dim3 grid(ceil_div(op->ny, 128), ceil_div(plan.r1[i] - plan.r0[i], rows), nvec);
        kern<<<grid, 128, 0, op->stream>>>(er, enl, op->ax, op->ax + op->nx, op->ay, op->ay + op->ny, d_x, d_y, op->nx,
                                           op->ny, p.omega * p.omega * p.e0, plan.r0[i], plan.r1[i]);
        ++g_fdfd_launches;
    }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}
template int op_apply_planes_t<cplx>(const FdfdOp*, const cplx*, cplx*, int);
template int op_apply_planes_t<cplx32>(const FdfdOp*, const cplx32*, cplx32*, int);
template int op_residual_t<cplx>(const FdfdOp*, const cplx*, const cplx*, cplx*, int);
template int op_residual_t<cplx32>(const FdfdOp*, const cplx32*, const cplx32*, cplx32*, int);
template int op_apply_fused_t<cplx>(const FdfdOp*, const cplx*, cplx*, int);
template int op_apply_fused_t<cplx32>(const FdfdOp*, const cplx32*, cplx32*, int);

int op_apply_planes(const FdfdOp* op, const cplx* d_x, cplx* d_y, int nvec) { return op_apply_planes_t<cplx>(op, d_x, d_y, nvec); }
int op_residual(const FdfdOp* op, const cplx* d_b, const cplx* d_x, cplx* d_r, int nvec) {
    return op_residual_t<cplx>(op, d_b, d_x, d_r, nvec);
}
int op_apply_fused(const FdfdOp* op, const cplx* d_x, cplx* d_y, int nvec) { return op_apply_fused_t<cplx>(op, d_x, d_y, nvec); }

int op_derive_fields(const FdfdOp* op, const cplx* d_x, cplx* d_f1, cplx* d_f2, int averaging) {
    AsmParams p = make_params(op);
    if (averaging >= 0) p.averaging = averaging;
    { derive_fields_kernel<<<ceil_div(op->n(), 256), 256, 0, op->stream>>>(
        d_x, op->eps_r, op->has_nl ? op->eps_nl : nullptr, op->isxb, op->isyb, d_f1, d_f2, p); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}
