// Batched multifrontal nested-dissection solver for the 5-point Maxwell stencil on a torus.
//
// Replaces the general sparse LU the reference calls in solver_direct (linalg.py:123-149).
// The operator is row-scaled by D = diag(sxf[ix] syf[iy]) on the fly: D A is complex SYMMETRIC (the
// forward stretch factors are the only asymmetry of the sc-PML operator, linalg.py:61-63 / 96-98),
// so every front keeps only its lower triangle.  Every level of the elimination tree is ONE batch
// of equally padded dense fronts
//        F = [ F_EE   .   ]      E: unknowns eliminated at this level (leaf interiors / shared lines)
//            [ F_RE  F_RR ]      R: the box ring handed to the parent
// and computes, with complex GEMMs on the FP64 tensor pipe (zgemm.cuh),
//        Einv = F_EE^-1                    blocked Gauss-Jordan, pivoting inside 64-wide tiles
//        G    = F_RE Einv                  (m x k) = (m x k)(k x k),       K = k
//        S    = F_RR - G F_RE^T            lower tiles only,               K = k
// Einv and G are the stored factors; the triangular solves of a classic multifrontal code become
// plain batched matrix-vector products:  ring = f_R - G f_E,  u_E = Einv f_E - G^T u_R.
#include <algorithm>
#include "direct.cuh"
#include "zgemm.cuh"
#include "mrhs.cuh"

// ------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx row_scale(const cplx* __restrict__ isxf, const cplx* __restrict__ isyf, int x, int y) {
    return crecip(cmul(isxf[x], isyf[y]));      // sxf[x] * syf[y]
}

__global__ void __launch_bounds__(64)
leaf_assemble_kernel(cplx* __restrict__ F, const cplx* __restrict__ planes, const cplx* __restrict__ isxf,
                     const cplx* __restrict__ isyf, const int* __restrict__ cls,
                     const int* __restrict__ k_cls, const int* __restrict__ x0, const int* __restrict__ y0,
                     const int* __restrict__ slot_lx, const int* __restrict__ slot_ly,
                     const int* __restrict__ slot_right, const int* __restrict__ slot_up, int kmax, int nmax,
                     int nx, int ny) {
    const long long b = blockIdx.x;
    const int c = cls[b];
    cplx* Fb = F + b * (long long)nmax * nmax;
    const size_t n = (size_t)nx * ny;
    for (int e = threadIdx.x; e < nmax * nmax; e += blockDim.x)
        if (e % nmax <= e / nmax) Fb[e] = make_double2(0.0, 0.0);
    __syncthreads();
    const int kc = k_cls[c];
    for (int s = threadIdx.x; s < nmax; s += blockDim.x) {
        if (s >= kc && s < kmax) Fb[s * nmax + s] = make_double2(1.0, 0.0);   // padded pivots
        int r = slot_right[c * nmax + s];
        if (r < 0) continue;                                                  // not an owner slot
        int u = slot_up[c * nmax + s];
        int x = x0[b] + slot_lx[c * nmax + s];
        int y = y0[b] + slot_ly[c * nmax + s];
        if (x >= nx) x -= nx;
        if (y >= ny) y -= ny;
        size_t node = (size_t)x * ny + y;
        const cplx d = row_scale(isxf, isyf, x, y);
        // scaled row `node`: its +x / +y couplings equal the scaled -x / -y couplings of the neighbours
        Fb[s * nmax + s] = cmul(planes[node], d);                                   // c0
        Fb[max(s, r) * nmax + min(s, r)] = cmul(planes[2 * n + node], d);           // cxp
        Fb[max(s, u) * nmax + min(s, u)] = cmul(planes[4 * n + node], d);           // cyp
    }
}

// One CTA assembles ASM_ROWS consecutive rows of one front (a warp per row, lanes along the row up to the
// diagonal): no integer division per element, no threads parked on the unused upper triangle, the row's own
// map entries are loaded once and the children's Schur rows are read along their fast index.
#define ASM_ROWS 32
__global__ void __launch_bounds__(256)
merge_assemble_kernel(cplx* __restrict__ F, const cplx* __restrict__ Fc, const int* __restrict__ cls,
                      const int* __restrict__ k_cls, const int* __restrict__ ch1, const int* __restrict__ ch2,
                      const int* __restrict__ inv1, const int* __restrict__ inv2, int kmax, int nmax,
                      int kc, int nc, long long cstride, int chunks) {
    // children's Schur blocks: batch stride cstride, leading dimension nc, first ring entry at (kc, kc)
    const long long b = blockIdx.x / chunks;
    const int r0 = (int)(blockIdx.x % chunks) * ASM_ROWS;
    const int c = cls[b];
    const int* i1 = inv1 + (size_t)c * nmax;
    const int* i2 = inv2 + (size_t)c * nmax;
    const cplx* S1 = Fc + (long long)ch1[b] * cstride + (size_t)kc * nc + kc;
    const cplx* S2 = Fc + (long long)ch2[b] * cstride + (size_t)kc * nc + kc;
    cplx* Fb = F + b * (long long)nmax * nmax;
    const int kcls = k_cls[c];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = r0 + w; p < min(r0 + ASM_ROWS, nmax); p += 8) {
        const int a1 = i1[p], a2 = i2[p];
        cplx* row = Fb + (size_t)p * nmax;
        for (int q = lane; q <= p; q += 32) {
            cplx v = make_double2(0.0, 0.0);
            if (a1 >= 0) {
                int bq = i1[q];
                if (bq >= 0) v = S1[(size_t)max(a1, bq) * nc + min(a1, bq)];
            }
            if (a2 >= 0) {
                int bq = i2[q];
                if (bq >= 0) v = cadd(v, S2[(size_t)max(a2, bq) * nc + min(a2, bq)]);
            }
            if (p == q && p >= kcls && p < kmax) v = make_double2(1.0, 0.0);
            row[q] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// tile inversion
// ------------------------------------------------------------------------------------------
// Register-resident in-place Gauss-Jordan inversion of a tile of up to TS x TS (TS = 32: 64 threads,
// TS = 64: 256 threads; each thread owns a 4 x 4 block) with IMPLICIT partial pivoting: rows are never
// moved, the pivot row of column c is the largest entry among the rows not used yet, found with one
// integer warp reduction on an order-preserving key.  Two barriers per column; only the pivot column
// and row travel through shared memory.  With pr(c) the pivot row of column c, the stored result M
// satisfies  A^-1[pc(r)][pr(c)] = M[r][c]  (pc = pr^-1), which the final store applies.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return r;
}
__device__ __forceinline__ cplx fast_crecip(cplx z) {
    // |z|^2 stays far inside the double range for operator entries (|z| ~ 1e-20 .. 1e20)
    double inv = fast_rcp(z.x * z.x + z.y * z.y);
    return make_double2(z.x * inv, -z.y * inv);
}

template <int TS>
struct TileInvSmem {
    cplx colbuf[2][TS];
    cplx rowbuf[TS];
    int prow_of_col[TS], pcol_of_row[TS];
};

template <int TS>
__device__ __forceinline__ void reg_tile_inverse(const cplx* S, int s_ld, int sym, int tw, cplx* D, int d_ld,
                                                 TileInvSmem<TS>& sm, int* info, int tid, int bar_id) {
    constexpr int GD = TS / 4, NTH = GD * GD;
    const int ty = tid / GD, tx = tid % GD, lane = tid & 31;
    const int R0 = ty * 4, C0 = tx * 4;
    auto bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(NTH) : "memory"); };
    cplx a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r = R0 + i, c = C0 + j;
            if (r < tw && c < tw) a[i][j] = (sym && c > r) ? S[(size_t)c * s_ld + r] : S[(size_t)r * s_ld + c];
            else a[i][j] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sm.colbuf[0][R0 + i] = a[i][0];
    }
    unsigned long long used = 0ull;
    for (int c = 0; c < tw; ++c) {
        const cplx* cb = sm.colbuf[c & 1];
        bar();
        // every warp finds the pivot row redundantly: key = high bits of |entry|^2, low bits = 63 - row
        unsigned key = 0;
#pragma unroll
        for (int h = 0; h < TS / 32; ++h) {
            int row = lane + 32 * h;
            if (row < tw && !((used >> row) & 1ull)) {
                unsigned hi = (unsigned)__double2hiint(cabs2(cb[row]));
                unsigned kk = (hi & 0xFFFFFFC0u) | (unsigned)(63 - row);
                key = kk > key ? kk : key;
            }
        }
        key = __reduce_max_sync(0xffffffffu, key);
        int p = 63 - (int)(key & 63u);
        const unsigned ex = (key >> 20) & 0x7FFu;
        if (ex == 0u || ex == 0x7FFu) {              // zero / denormal / inf / nan pivot: flag it, keep going
            if (ex == 0u) p = __ffsll((long long)(~used)) - 1;
            if (tid == 0) atomicExch(info, 1);
        }
        if (ty == (p >> 2)) {
            const int pi = p & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cplx t = a[0][j];
                if (pi == 1) t = a[1][j];
                if (pi == 2) t = a[2][j];
                if (pi == 3) t = a[3][j];
                sm.rowbuf[C0 + j] = (C0 + j == c) ? make_double2(1.0, 0.0) : t;
            }
        }
        bar();
        const cplx ipiv = fast_crecip(cb[p]);
        cplx pr[4], f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[j] = cmul(sm.rowbuf[C0 + j], ipiv);
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = cb[R0 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool piv = (R0 + i) == p;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cplx t = (C0 + j == c) ? make_double2(0.0, 0.0) : a[i][j];
                t.x -= f[i].x * pr[j].x - f[i].y * pr[j].y;
                t.y -= f[i].x * pr[j].y + f[i].y * pr[j].x;
                a[i][j] = piv ? pr[j] : t;
            }
        }
        used |= 1ull << p;
        if (tid == 0) { sm.prow_of_col[c] = p; sm.pcol_of_row[p] = c; }
        const int cn = c + 1;
        if (cn < tw && tx == (cn >> 2)) {
            const int cj = cn & 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cplx t = a[i][0];
                if (cj == 1) t = a[i][1];
                if (cj == 2) t = a[i][2];
                if (cj == 3) t = a[i][3];
                sm.colbuf[cn & 1][R0 + i] = t;
            }
        }
    }
    bar();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (R0 + i >= tw) continue;
        const int orow = sm.pcol_of_row[R0 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (C0 + j >= tw) continue;
            D[(size_t)orow * d_ld + sm.prow_of_col[C0 + j]] = a[i][j];
        }
    }
}

template <int TS>
__global__ void __launch_bounds__(TS * TS / 16)
tile_inverse_kernel(const cplx* src, long long s_stride, int s_ld, int tw, cplx* dst, long long d_stride, int d_ld,
                    int* __restrict__ info, int sym) {
    __shared__ TileInvSmem<TS> sm;
    reg_tile_inverse<TS>(src + (long long)blockIdx.x * s_stride, s_ld, sym, tw, dst + (long long)blockIdx.x * d_stride,
                         d_ld, sm, info, threadIdx.x, 1);
}

static void launch_tile_inverse(const cplx* src, long long s_stride, int s_ld, int tw, cplx* dst, long long d_stride,
                                int d_ld, int* info, int sym, long long nb, cudaStream_t st) {
    if (tw <= 32)
        tile_inverse_kernel<32><<<(unsigned)nb, 64, 0, st>>>(src, s_stride, s_ld, tw, dst, d_stride, d_ld, info, sym);
    else
        tile_inverse_kernel<64><<<(unsigned)nb, 256, 0, st>>>(src, s_stride, s_ld, tw, dst, d_stride, d_ld, info, sym);
    ++g_fdfd_launches;
}

// One warp inverts a k x k matrix (k <= KMAX <= 32) held one ROW per lane in registers: in-place
// Gauss-Jordan with implicit partial pivoting (see reg_tile_inverse), pivot search by one integer warp
// reduction, pivot row broadcast by shuffles; no shared memory, no barriers.  E / Einv are row-major
// with leading dimension ld; Einv may alias E.
template <int KMAX>
__device__ __forceinline__ void warp_invert(const cplx* E, cplx* Einv, int ld, int k, int* info) {
    const int lane = threadIdx.x & 31;
    cplx a[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        a[j] = (lane < k && j < k) ? E[lane * ld + j] : make_double2(lane == j ? 1.0 : 0.0, 0.0);
    unsigned used = 0u;
    int my_col = lane;                                  // column this row was the pivot of
    int prow_of_col = lane;                             // lane c keeps the pivot row of column c
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
        if (c < k) {                                    // warp-uniform
            unsigned key = 0u;
            if (lane < k && !((used >> lane) & 1u))
                key = ((unsigned)__double2hiint(cabs2(a[c])) & 0xFFFFFFE0u) | (unsigned)(31 - lane);
            key = __reduce_max_sync(0xffffffffu, key);
            int p = 31 - (int)(key & 31u);
            const unsigned ex = (key >> 20) & 0x7FFu;
            if (ex == 0u || ex == 0x7FFu) {
                if (ex == 0u) p = __ffs((int)(~used)) - 1;
                if (lane == 0) atomicExch(info, 1);
            }
            cplx piv;
            piv.x = __shfl_sync(0xffffffffu, a[c].x, p);
            piv.y = __shfl_sync(0xffffffffu, a[c].y, p);
            const cplx ipiv = fast_crecip(piv);
            const cplx f = a[c];
            const bool is_p = lane == p;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                cplx t = (j == c) ? make_double2(1.0, 0.0) : a[j];
                cplx pr;
                pr.x = __shfl_sync(0xffffffffu, t.x, p);
                pr.y = __shfl_sync(0xffffffffu, t.y, p);
                pr = cmul(pr, ipiv);
                cplx v = (j == c) ? make_double2(0.0, 0.0) : a[j];
                v.x -= f.x * pr.x - f.y * pr.y;
                v.y -= f.x * pr.y + f.y * pr.x;
                a[j] = is_p ? pr : v;
            }
            used |= 1u << p;
            if (is_p) my_col = c;
            if (lane == c) prow_of_col = p;
        }
    }
    // A^-1[pc(r)][pr(c)] = M[r][c]:  this lane's row goes to output row my_col, its entry j to column pr(j)
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        int oc = __shfl_sync(0xffffffffu, prow_of_col, j);
        if (lane < k && j < k) Einv[my_col * ld + oc] = a[j];
    }
}

// ------------------------------------------------------------------------------------------
// fused small-front level: one CTA per front does assemble -> Einv -> G -> S in shared memory.
// Used for the bottom of the tree (k <= 16), where a level is millions of tiny fronts and the
// generic path (four kernels, the full padded front written to and re-read from HBM) is pure memory
// traffic.  Global traffic here: child Schur blocks (or planes) in, Einv / G / compact S out.
// ------------------------------------------------------------------------------------------
struct SmallFrontArgs {
    int kind, kmax, mmax, nx, ny;
    const int *cls, *k_cls;
    const int *x0, *y0, *slot_lx, *slot_ly, *slot_right, *slot_up;     // leaf tables
    const int *ch1, *ch2, *inv1, *inv2;                                // merge tables
    const cplx *planes, *isxf, *isyf;
    const cplx* Sc;      // child Schur blocks: entry (a, b), a >= b, of child c at Sc[c * sc + (kc + a) * nc + kc + b]
    long long sc;
    int kc, nc;
    cplx *Einv, *G, *S;  // outputs; S compact [nb][mmax][mmax], lower triangle
    int* info;
};

__device__ __forceinline__ void small_front_put(cplx* W, cplx* R, cplx* Sm, int k, int m, int p, int q, cplx v) {
    // p >= q in front slot numbering (E slots first)
    if (p < k) {
        W[p * 2 * k + q] = v;
        W[q * 2 * k + p] = v;
    } else if (q < k) {
        R[(p - k) * k + q] = v;
    } else {
        Sm[(p - k) * m + (q - k)] = v;
    }
}

// sum of the two children's Schur entries that land on front entry (p, q), p >= q
__device__ __forceinline__ cplx small_front_gather(const SmallFrontArgs& a, const cplx* S1, const cplx* S2,
                                                   const int* i1, const int* i2, int p, int q) {
    cplx v = make_double2(0.0, 0.0);
    int a1 = i1[p], b1 = i1[q], a2 = i2[p], b2 = i2[q];
    if (a1 >= 0 && b1 >= 0) v = S1[(size_t)(a.kc + max(a1, b1)) * a.nc + a.kc + min(a1, b1)];
    if (a2 >= 0 && b2 >= 0) v = cadd(v, S2[(size_t)(a.kc + max(a2, b2)) * a.nc + a.kc + min(a2, b2)]);
    return v;
}

// KT / MT > 0: front sizes known at compile time (the levels of a power-of-two grid), so the index arithmetic
// folds to constants and the short inner products unroll; 0: taken from the arguments.
// A front is worked by WPF warps (1, 2, 4 or 8); a CTA holds several fronts (blockDim.x / (32 WPF)), each group with its
// own slice of shared memory and its own barrier -- the tiny leaf-side fronts no longer leave half a CTA idle while
// one warp inverts, and a resident CTA carries 4-8 fronts instead of one.
template <int WPF>
__device__ __forceinline__ void front_sync(int group) {
    if (WPF == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(WPF * 32) : "memory");
}
template <int KMAX, int KT, int MT, int WPF>
__global__ void small_front_kernel(SmallFrontArgs a, long long nb, int front_smem_cplx) {
    extern __shared__ __align__(16) unsigned char sf_smem[];
    const int k = KT ? KT : a.kmax, m = MT ? MT : a.mmax, n = k + m, w2 = 2 * k;
    const int group = threadIdx.x / (32 * WPF), tid = threadIdx.x % (32 * WPF), nt = 32 * WPF;
    const int fpc = blockDim.x / (32 * WPF);
    const long long b = (long long)blockIdx.x * fpc + group;
    if (b >= nb) return;                                // whole groups leave together: no barrier is left waiting
    cplx* W = reinterpret_cast<cplx*>(sf_smem) + (size_t)group * front_smem_cplx;   // [k][2k]: row r = [ E[r][:] | Einv[r][:] ]
    cplx* R = W + (size_t)k * w2;                     // [m][k]   F_RE
    cplx* Gs = R + (size_t)m * k;                     // [m][k]   G
    cplx* Sm = Gs + (size_t)m * k;                    // [m][m]   F_RR, leaf levels only
    const bool leaf = a.kind == 0;
    int* s_i1 = reinterpret_cast<int*>(W + (size_t)front_smem_cplx - (leaf ? 0 : 2 * ((n + 3) / 4)));   // [n] + [n] ints at the tail
    int* s_i2 = s_i1 + n;
    const int c = a.cls[b];
    const int kcls = a.k_cls[c];
    const cplx zero = make_double2(0.0, 0.0);
    const cplx *S1 = nullptr, *S2 = nullptr;
    {
        const int nz = k * w2 + m * k + (leaf ? m * k + m * m : 0);      // W, R (and Gs, Sm for a leaf)
        for (int e = tid; e < nz; e += nt) W[e] = zero;
    }
    if (!leaf) {
        for (int i = tid; i < n; i += nt) {
            s_i1[i] = a.inv1[(size_t)c * n + i];
            s_i2[i] = a.inv2[(size_t)c * n + i];
        }
        S1 = a.Sc + (long long)a.ch1[b] * a.sc;
        S2 = a.Sc + (long long)a.ch2[b] * a.sc;
    }
    front_sync<WPF>(group);
    for (int r = kcls + tid; r < k; r += nt) W[r * w2 + r] = make_double2(1.0, 0.0);       // padded pivots
    // ---- assemble F_EE, F_RE (and F_RR for a leaf)
    if (leaf) {
        const size_t ncell = (size_t)a.nx * a.ny;
        for (int s = tid; s < n; s += nt) {
            int r = a.slot_right[c * n + s];
            if (r < 0) continue;
            int u = a.slot_up[c * n + s];
            int x = a.x0[b] + a.slot_lx[c * n + s], y = a.y0[b] + a.slot_ly[c * n + s];
            if (x >= a.nx) x -= a.nx;
            if (y >= a.ny) y -= a.ny;
            size_t node = (size_t)x * a.ny + y;
            const cplx d = row_scale(a.isxf, a.isyf, x, y);
            small_front_put(W, R, Sm, k, m, s, s, cmul(a.planes[node], d));
            small_front_put(W, R, Sm, k, m, max(s, r), min(s, r), cmul(a.planes[2 * ncell + node], d));
            small_front_put(W, R, Sm, k, m, max(s, u), min(s, u), cmul(a.planes[4 * ncell + node], d));
        }
    } else {
        for (int e = tid; e < n * k; e += nt) {          // columns q < k of the lower triangle
            int p = e / k, q = e - p * k;
            if (q > p || (p == q && p >= kcls)) continue;                // padded pivot keeps its 1
            cplx v = small_front_gather(a, S1, S2, s_i1, s_i2, p, q);
            small_front_put(W, R, Sm, k, m, p, q, v);
        }
    }
    front_sync<WPF>(group);
    // ---- Einv (right half of W) by the group's first warp, one row per lane
    if (tid < 32) warp_invert<KMAX>(W, W + k, w2, k, a.info);
    front_sync<WPF>(group);
    // ---- G = F_RE Einv
    cplx* Eo = a.Einv + b * (long long)k * k;
    for (int e = tid; e < k * k; e += nt) Eo[e] = W[(e / k) * w2 + k + e % k];
    cplx* Go = a.G + b * (long long)m * k;
    for (int e = tid; e < m * k; e += nt) {
        int i = e / k, cc = e - i * k;
        cplx acc = zero;
#pragma unroll
        for (int l = 0; l < (KT ? KT : k); ++l) cfma(acc, R[i * k + l], W[l * w2 + k + cc]);
        Gs[e] = acc;
        Go[e] = acc;
    }
    front_sync<WPF>(group);
    // ---- S = F_RR - G F_RE^T (lower), F_RR gathered from the children on the fly
    cplx* So = a.S + b * (long long)m * m;
    if ((m & 1) == 0) {
        // 2 x 2 register blocks over the lower triangle only: 4 shared-memory loads feed 4 complex MACs (the scalar loop
        // below needs 8), and no thread is parked on the unused upper half
        const int mb = m >> 1, nblk2 = mb * (mb + 1) / 2;
        for (int e = tid; e < nblk2; e += nt) {
            int bi = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while ((bi + 1) * (bi + 2) / 2 <= e) ++bi;
            while (bi * (bi + 1) / 2 > e) --bi;
            const int bj = e - bi * (bi + 1) / 2, i0 = 2 * bi, j0 = 2 * bj;
            const bool diag = bi == bj;
            cplx a00, a01 = zero, a10, a11;
            if (leaf) {
                a00 = Sm[i0 * m + j0]; a10 = Sm[(i0 + 1) * m + j0]; a11 = Sm[(i0 + 1) * m + j0 + 1];
                if (!diag) a01 = Sm[i0 * m + j0 + 1];
            } else {
                a00 = small_front_gather(a, S1, S2, s_i1, s_i2, k + i0, k + j0);
                a10 = small_front_gather(a, S1, S2, s_i1, s_i2, k + i0 + 1, k + j0);
                a11 = small_front_gather(a, S1, S2, s_i1, s_i2, k + i0 + 1, k + j0 + 1);
                if (!diag) a01 = small_front_gather(a, S1, S2, s_i1, s_i2, k + i0, k + j0 + 1);
            }
            const cplx *g0p = Gs + i0 * k, *g1p = g0p + k, *r0p = R + j0 * k, *r1p = r0p + k;
#pragma unroll
            for (int l = 0; l < (KT ? KT : k); ++l) {
                const cplx g0 = g0p[l], g1 = g1p[l], r0 = r0p[l], r1 = r1p[l];
                a00.x -= g0.x * r0.x - g0.y * r0.y; a00.y -= g0.x * r0.y + g0.y * r0.x;
                a01.x -= g0.x * r1.x - g0.y * r1.y; a01.y -= g0.x * r1.y + g0.y * r1.x;
                a10.x -= g1.x * r0.x - g1.y * r0.y; a10.y -= g1.x * r0.y + g1.y * r0.x;
                a11.x -= g1.x * r1.x - g1.y * r1.y; a11.y -= g1.x * r1.y + g1.y * r1.x;
            }
            So[i0 * m + j0] = a00;
            if (!diag) So[i0 * m + j0 + 1] = a01;
            So[(i0 + 1) * m + j0] = a10;
            So[(i0 + 1) * m + j0 + 1] = a11;
        }
        return;
    }
    for (int e = tid; e < m * m; e += nt) {
        int i = e / m, j = e - i * m;
        if (j > i) continue;
        cplx acc = leaf ? Sm[e] : small_front_gather(a, S1, S2, s_i1, s_i2, k + i, k + j);
#pragma unroll
        for (int l = 0; l < (KT ? KT : k); ++l) {
            cplx g = Gs[i * k + l], r = R[j * k + l];
            acc.x -= g.x * r.x - g.y * r.y;
            acc.y -= g.x * r.y + g.y * r.x;
        }
        So[e] = acc;
    }
}

// complex entries of shared memory one front needs (the two child maps of a merge front ride at the tail)
static size_t small_front_cplx(int k, int m, bool leaf) {
    return (size_t)k * 2 * k + 2 * (size_t)m * k + (leaf ? (size_t)m * m : 2 * (size_t)((k + m + 3) / 4));
}
static size_t small_front_smem(int k, int m, bool leaf) { return sizeof(cplx) * small_front_cplx(k, m, leaf); }
static bool small_front_ok(int k, int m, bool leaf) {
    return k <= 16 && k + m <= 128 && small_front_smem(k, m, leaf) <= 200 * 1024;
}

// dst[b][c][r] = src[b][r][c] for an (rows x cols) block; 32 x 32 tiles through shared memory.
// mirror != 0 (square, src == dst): copies the strict lower triangle onto the upper one instead.
__global__ void __launch_bounds__(256)
transpose_kernel(const cplx* src, long long s_stride, int s_ld, cplx* dst, long long d_stride, int d_ld, int rows,
                 int cols, int mirror) {
    __shared__ cplx tile[32][33];
    const int tr = blockIdx.y, tc = blockIdx.x;
    if (mirror && tc > tr) return;
    const cplx* S = src + (long long)blockIdx.z * s_stride;
    cplx* D = dst + (long long)blockIdx.z * d_stride;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int i = ly; i < 32; i += 8) {
        int r = tr * 32 + i, c = tc * 32 + lx;
        if (r < rows && c < cols) tile[i][lx] = S[(size_t)r * s_ld + c];
    }
    __syncthreads();
    for (int i = ly; i < 32; i += 8) {
        int c = tc * 32 + i, r = tr * 32 + lx;       // element (r, c) of src goes to (c, r) of dst
        if (r < rows && c < cols && (!mirror || r > c)) D[(size_t)c * d_ld + r] = tile[lx][i];
    }
}

// Einv[b] (kmax x kmax, full) = symmetric expansion of the lower triangle of F[b][0:kmax, 0:kmax]
__global__ void __launch_bounds__(256)
sym_expand_kernel(const cplx* __restrict__ F, cplx* __restrict__ E, int kmax, int ld, long long fstride, int chunks) {
    const long long b = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const cplx* Fb = F + b * fstride;
    cplx* Eb = E + b * (long long)kmax * kmax;
    const long long total = (long long)kmax * kmax;
    const long long per = (total + chunks - 1) / chunks;
    const long long e0 = chunk * per, e1 = min(total, e0 + per);
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        int p = (int)(e / kmax), q = (int)(e % kmax);
        Eb[e] = Fb[(size_t)max(p, q) * ld + min(p, q)];
    }
}

// ------------------------------------------------------------------------------------------
// solve-phase kernels.  Front vectors are [slot][NR] with the right-hand-side index fastest.
// ------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(128)
leaf_gather_kernel(cplx* __restrict__ f, const cplx* __restrict__ rhs, const cplx* __restrict__ isxf,
                   const cplx* __restrict__ isyf, const int* __restrict__ cls,
                   const int* __restrict__ x0, const int* __restrict__ y0, const int* __restrict__ slot_lx,
                   const int* __restrict__ slot_ly, const int* __restrict__ slot_right, int nmax, int nx, int ny,
                   int nr_act, long long nb) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * nmax) return;
    long long b = gid / nmax;
    int s = (int)(gid % nmax);
    int c = cls[b];
    cplx v[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = make_double2(0.0, 0.0);
    if (slot_right[c * nmax + s] >= 0) {
        int x = x0[b] + slot_lx[c * nmax + s], y = y0[b] + slot_ly[c * nmax + s];
        if (x >= nx) x -= nx;
        if (y >= ny) y -= ny;
        size_t node = (size_t)x * ny + y, n = (size_t)nx * ny;
        const cplx d = row_scale(isxf, isyf, x, y);       // the factors belong to D A: solve D A x = D b
#pragma unroll
        for (int j = 0; j < NR; ++j)
            if (j < nr_act) v[j] = cmul(rhs[j * n + node], d);
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) f[gid * NR + j] = v[j];
}

template <int NR>
__global__ void __launch_bounds__(128)
leaf_scatter_kernel(const cplx* __restrict__ u, cplx* __restrict__ out, const int* __restrict__ cls,
                    const int* __restrict__ x0, const int* __restrict__ y0, const int* __restrict__ slot_lx,
                    const int* __restrict__ slot_ly, const int* __restrict__ slot_right, int nmax, int nx, int ny,
                    int nr_act, long long nb) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * nmax) return;
    long long b = gid / nmax;
    int s = (int)(gid % nmax);
    int c = cls[b];
    if (slot_right[c * nmax + s] < 0) return;
    int x = x0[b] + slot_lx[c * nmax + s], y = y0[b] + slot_ly[c * nmax + s];
    if (x >= nx) x -= nx;
    if (y >= ny) y -= ny;
    size_t node = (size_t)x * ny + y, n = (size_t)nx * ny;
#pragma unroll
    for (int j = 0; j < NR; ++j)
        if (j < nr_act) out[j * n + node] = u[gid * NR + j];
}

// f[b][p] = ring_child1[inv1[p]] + ring_child2[inv2[p]]
template <int NR>
__global__ void __launch_bounds__(128)
merge_gather_kernel(cplx* __restrict__ f, const cplx* __restrict__ ring_c, const int* __restrict__ cls,
                    const int* __restrict__ ch1, const int* __restrict__ ch2, const int* __restrict__ inv1,
                    const int* __restrict__ inv2, int nmax, int mc, long long nb) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * nmax) return;
    long long b = gid / nmax;
    int p = (int)(gid % nmax);
    int c = cls[b];
    int a1 = inv1[(size_t)c * nmax + p], a2 = inv2[(size_t)c * nmax + p];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        cplx v = make_double2(0.0, 0.0);
        if (a1 >= 0) v = ring_c[((long long)ch1[b] * mc + a1) * NR + j];
        if (a2 >= 0) v = cadd(v, ring_c[((long long)ch2[b] * mc + a2) * NR + j]);
        f[gid * NR + j] = v;
    }
}

// children pick their ring values out of the parent's solved front:  u_child[kc+i] = u_par[cmap[i]]
template <int NR>
__global__ void __launch_bounds__(128)
child_scatter_kernel(cplx* __restrict__ uc, const cplx* __restrict__ up, const int* __restrict__ cls,
                     const int* __restrict__ ch1, const int* __restrict__ ch2, const int* __restrict__ c1map,
                     const int* __restrict__ c2map, int nmax, int mc, int kc, int nc, long long nb) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * 2 * mc) return;
    long long b = gid / (2 * mc);
    int r = (int)(gid % (2 * mc));
    int which = r / mc, i = r % mc;
    int c = cls[b];
    int slot = (which ? c2map : c1map)[(size_t)c * mc + i];
    if (slot < 0) return;
    long long ch = which ? ch2[b] : ch1[b];
#pragma unroll
    for (int j = 0; j < NR; ++j) uc[(ch * nc + kc + i) * NR + j] = up[(b * nmax + slot) * NR + j];
}

// forward: yE = Einv f_E ; ring = f_R - G f_E.   One warp per output row; the 32 lanes are 32 / NR groups of
// NR lanes: a group walks every (32 / NR)-th column, its lanes take one right-hand side each (the matrix entry
// is one broadcast load per group, the NR vector entries one contiguous load), groups are summed by shuffles.
template <int NR>
__global__ void __launch_bounds__(256)
forward_mv_kernel(const cplx* __restrict__ Einv, const cplx* __restrict__ G, const cplx* __restrict__ f,
                  cplx* __restrict__ yE, cplx* __restrict__ ring, int kmax, int mmax, int nmax, long long nb) {
    constexpr int NG = 32 / NR;
    long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nb * nmax) return;
    const long long b = gw / nmax;
    const int r = (int)(gw % nmax);
    const int g = lane / NR, j = lane % NR;
    const cplx* row = r < kmax ? Einv + (b * kmax + r) * kmax : G + (b * mmax + (r - kmax)) * kmax;
    const cplx* v = f + b * nmax * NR;
    cplx acc = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int c = g; c < kmax; c += NG) cfma(acc, ldg_c(row + c), v[c * NR + j]);
#pragma unroll
    for (int o = 16; o >= NR; o >>= 1) {
        acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
    }
    if (g == 0) {
        if (r < kmax) yE[(b * kmax + r) * NR + j] = acc;
        else ring[(b * mmax + (r - kmax)) * NR + j] = csub(v[r * NR + j], acc);
    }
}

// the same with one THREAD per output row, for the levels of tiny fronts (k <= 16) where a warp per
// row would leave most lanes idle
template <int NR>
__global__ void __launch_bounds__(256)
forward_mv_thread_kernel(const cplx* __restrict__ Einv, const cplx* __restrict__ G, const cplx* __restrict__ f,
                         cplx* __restrict__ yE, cplx* __restrict__ ring, int kmax, int mmax, int nmax, long long nb) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * nmax) return;
    long long b = gid / nmax;
    int r = (int)(gid % nmax);
    const cplx* row = r < kmax ? Einv + (b * kmax + r) * kmax : G + (b * mmax + (r - kmax)) * kmax;
    const cplx* v = f + b * nmax * NR;
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    for (int c = 0; c < kmax; ++c) {
        cplx m = ldg_c(row + c);
#pragma unroll
        for (int j = 0; j < NR; ++j) cfma(acc[j], m, v[c * NR + j]);
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        if (r < kmax) yE[(b * kmax + r) * NR + j] = acc[j];
        else ring[(b * mmax + (r - kmax)) * NR + j] = csub(v[r * NR + j], acc[j]);
    }
}

// backward: u_E = yE - G^T u_R.  Lanes run along the E index (the contiguous one of G), the 8 warps
// of a CTA split the ring rows and reduce through shared memory (fixed order: deterministic).
template <int NR>
__global__ void __launch_bounds__(256)
backward_mvt_kernel(const cplx* __restrict__ G, const cplx* __restrict__ yE, cplx* __restrict__ u, int kmax,
                    int mmax, int nmax, long long nb) {
    __shared__ cplx part[8][32][NR];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long gid = (long long)blockIdx.x * 32 + lane;       // (front, E slot)
    const bool ok = gid < nb * kmax;
    const long long b = ok ? gid / kmax : 0;
    const int r = ok ? (int)(gid % kmax) : 0;
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    if (ok) {
        const cplx* col = G + b * (long long)mmax * kmax + r;
        const cplx* v = u + (b * nmax + kmax) * NR;
#pragma unroll 4
        for (int c = w; c < mmax; c += 8) {
            cplx m = ldg_c(col + (size_t)c * kmax);
#pragma unroll
            for (int j = 0; j < NR; ++j) cfma(acc[j], m, v[c * NR + j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) part[w][lane][j] = acc[j];
    __syncthreads();
    if (w == 0 && ok) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            cplx t = part[0][lane][j];
#pragma unroll
            for (int q = 1; q < 8; ++q) t = cadd(t, part[q][lane][j]);
            u[(b * nmax + r) * NR + j] = csub(yE[(b * kmax + r) * NR + j], t);
        }
    }
}

// the same with one THREAD per E slot, for the levels of tiny fronts (k <= 16, short rings): no shared memory,
// no barriers; the threads of a front read consecutive entries of a G row and the same ring entries.
template <int NR>
__global__ void __launch_bounds__(128)
backward_mvt_thread_kernel(const cplx* __restrict__ G, const cplx* __restrict__ yE, cplx* __restrict__ u, int kmax,
                           int mmax, int nmax, long long nb) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (front, E slot)
    if (gid >= nb * kmax) return;
    const long long b = gid / kmax;
    const int r = (int)(gid % kmax);
    const cplx* col = G + b * (long long)mmax * kmax + r;
    const cplx* v = u + (b * nmax + kmax) * NR;
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    for (int c = 0; c < mmax; ++c) {
        const cplx m = ldg_c(col + (size_t)c * kmax);
#pragma unroll
        for (int j = 0; j < NR; ++j) cfma(acc[j], m, v[c * NR + j]);
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) u[(b * nmax + r) * NR + j] = csub(yE[gid * NR + j], acc[j]);
}

// The same product for levels with FEW fronts and long rings (the top of the tree): the ring rows are also
// split over gridDim.y CTAs, each writes its partial sums, a second kernel adds them in a fixed order.
template <int NR>
__global__ void __launch_bounds__(256)
backward_mvt_split_kernel(const cplx* __restrict__ G, const cplx* __restrict__ u, cplx* __restrict__ tmp, int kmax,
                          int mmax, int nmax, long long nb, int rows_per_split) {
    __shared__ cplx part[8][32][NR];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long gid = (long long)blockIdx.x * 32 + lane;       // (front, E slot)
    const bool ok = gid < nb * kmax;
    const long long b = ok ? gid / kmax : 0;
    const int r = ok ? (int)(gid % kmax) : 0;
    const int c0 = blockIdx.y * rows_per_split, c1 = min(mmax, c0 + rows_per_split);
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    if (ok) {
        const cplx* col = G + b * (long long)mmax * kmax + r;
        const cplx* v = u + (b * nmax + kmax) * NR;
#pragma unroll 4
        for (int c = c0 + w; c < c1; c += 8) {
            cplx m = ldg_c(col + (size_t)c * kmax);
#pragma unroll
            for (int j = 0; j < NR; ++j) cfma(acc[j], m, v[c * NR + j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) part[w][lane][j] = acc[j];
    __syncthreads();
    if (w == 0 && ok) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            cplx t = part[0][lane][j];
#pragma unroll
            for (int q = 1; q < 8; ++q) t = cadd(t, part[q][lane][j]);
            tmp[((long long)blockIdx.y * nb * kmax + gid) * NR + j] = t;
        }
    }
}
template <int NR>
__global__ void backward_mvt_reduce_kernel(const cplx* __restrict__ tmp, const cplx* __restrict__ yE,
                                           cplx* __restrict__ u, int kmax, int nmax, long long nb, int nsplit) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nb * kmax) return;
    const long long b = gid / kmax;
    const int r = (int)(gid % kmax);
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        cplx t = make_double2(0.0, 0.0);
        for (int q = 0; q < nsplit; ++q) t = cadd(t, tmp[((long long)q * nb * kmax + gid) * NR + j]);
        u[(b * nmax + r) * NR + j] = csub(yE[gid * NR + j], t);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int upload_i32(int** dst, const int* src, size_t count) {
    *dst = nullptr;
    if (!src || count == 0) return 0;
    FDFD_CHECK(cudaMalloc(dst, count * sizeof(int)));
    FDFD_CHECK(cudaMemcpy(*dst, src, count * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

int nd_create(NdSolver** out, int nx, int ny, int tile) {
    if (tile < 1 || tile > 64) FDFD_FAIL("tile must be in 1..64");
    NdSolver* s = new NdSolver();
    s->nx = nx; s->ny = ny; s->tile = tile; s->factored = false;
    s->fact_op = nullptr; s->fact_version = 0;
    s->factor_bytes = 0; s->factor_flops = 0;
    s->ws_a = s->ws_b = s->ws_ring_a = s->ws_ring_b = s->ws_ye = nullptr;
    s->ws_vec_cap = s->ws_ring_cap = s->ws_ye_cap = 0;
    s->d_info = nullptr;
    s->fws = nullptr;
    s->fws_cap = 0;
    s->comm = nullptr;
    s->ws_refine = nullptr;
    s->ws_refine_cap = 0;
    s->ws_bsplit = nullptr;
    s->ws_bsplit_cap = 0;
    s->xchg = nullptr;
    s->xchg_cap = 0;
    s->dist_send = s->dist_recv = s->dist_panel = nullptr;
    s->dist_send_cap = s->dist_recv_cap = s->dist_panel_cap = 0;
    {
        int lo = 0, hi = 0;
        FDFD_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        FDFD_CHECK(cudaStreamCreateWithPriority(&s->la_stream, cudaStreamNonBlocking, hi));
        FDFD_CHECK(cudaEventCreateWithFlags(&s->la_ready, cudaEventDisableTiming));
        FDFD_CHECK(cudaEventCreateWithFlags(&s->la_done, cudaEventDisableTiming));
        FDFD_CHECK(cudaEventCreateWithFlags(&s->zg_fork, cudaEventDisableTiming));
        FDFD_CHECK(cudaEventCreateWithFlags(&s->zg_join, cudaEventDisableTiming));
    }
    FDFD_CHECK(cudaMalloc(&s->d_info, sizeof(int)));
    *out = s;
    return 0;
}

// A/B switch (environment FDFD_INPLACE_CHAINS=0): chain levels re-assemble their fronts instead of working in
// place on the previous Schur blocks
static bool inplace_chains_enabled() {
    const char* e = getenv("FDFD_INPLACE_CHAINS");
    return !(e && e[0] == '0');
}

int nd_add_level(NdSolver* s, const NdLevelDesc* d) {
    NdLevel L;
    memset(&L, 0, sizeof(L));
    L.kind = d->kind; L.nb = d->nb; L.kmax = d->kmax; L.mmax = d->mmax; L.nmax = d->kmax + d->mmax;
    L.ncls = d->ncls; L.child_mmax = d->child_mmax;
    L.send_to = d->send_to; L.recv_from = d->recv_from;
    if (L.nb > 1 && L.recv_from >= 0) FDFD_FAIL("a level that receives a child front must hold exactly one front");
    if (L.nb > 0 && L.send_to >= 0) FDFD_FAIL("a level that sends its child front away cannot own fronts");
    if (upload_i32(&L.cls, d->cls, d->nb)) return -1;
    if (upload_i32(&L.k_cls, d->k_cls, d->ncls)) return -1;
    if (d->kind == 0) {
        size_t t = (size_t)d->ncls * L.nmax;
        if (upload_i32(&L.x0, d->x0, d->nb) || upload_i32(&L.y0, d->y0, d->nb) ||
            upload_i32(&L.slot_lx, d->slot_lx, t) || upload_i32(&L.slot_ly, d->slot_ly, t) ||
            upload_i32(&L.slot_right, d->slot_right, t) || upload_i32(&L.slot_up, d->slot_up, t))
            return -1;
    } else {
        if (s->levels.empty()) FDFD_FAIL("merge level before any leaf level");
        size_t t = (size_t)d->ncls * d->child_mmax;
        if (upload_i32(&L.ch1, d->ch1, d->nb) || upload_i32(&L.ch2, d->ch2, d->nb) ||
            upload_i32(&L.c1map, d->c1map, t) || upload_i32(&L.c2map, d->c2map, t))
            return -1;
        // inverse maps: front slot -> child ring position
        std::vector<int> i1((size_t)d->ncls * L.nmax, -1), i2((size_t)d->ncls * L.nmax, -1);
        for (int c = 0; c < d->ncls; ++c)
            for (int i = 0; i < d->child_mmax; ++i) {
                int a = d->c1map[(size_t)c * d->child_mmax + i];
                int b = d->c2map[(size_t)c * d->child_mmax + i];
                if (a >= L.nmax || b >= L.nmax) FDFD_FAIL("child map out of range");
                if (a >= 0) i1[(size_t)c * L.nmax + a] = i;
                if (b >= 0) i2[(size_t)c * L.nmax + b] = i;
            }
        if (upload_i32(&L.inv1, i1.data(), i1.size()) || upload_i32(&L.inv2, i2.data(), i2.size())) return -1;
        // chain level whose front IS the previous level's Schur block, slot for slot (same fronts, one child,
        // identity map, no padded pivots): it is factorised in place, without an assembly pass
        const NdLevel& P = s->levels.back();
        bool ip = inplace_chains_enabled() && d->nb == P.nb && L.nmax == d->child_mmax && d->child_mmax == P.mmax &&
                  L.send_to < 0 && L.recv_from < 0 && P.kmax > 16;
        for (int b = 0; ip && b < d->nb; ++b) ip = d->ch1[b] == b && d->ch2[b] == b;
        for (int c = 0; ip && c < d->ncls; ++c) {
            ip = d->k_cls[c] == d->kmax;
            for (int i = 0; ip && i < d->child_mmax; ++i) {
                const int a = d->c1map[(size_t)c * d->child_mmax + i], b2 = d->c2map[(size_t)c * d->child_mmax + i];
                ip = b2 < 0 && (a == i || a < 0);
            }
        }
        L.inplace = ip ? 1 : 0;
    }
    s->levels.push_back(L);
    s->factored = false;
    return 0;
}

static void free_level_factors(NdLevel& L) {
    if (L.Einv) cudaFree(L.Einv);
    if (L.G) cudaFree(L.G);
    L.Einv = L.G = nullptr;
}

static void dist_destroy(NdSolver* s);
void nd_destroy(NdSolver* s) {
    if (!s) return;
    dist_destroy(s);
    for (auto& L : s->levels) {
        free_level_factors(L);
        int* ptrs[] = {L.cls, L.k_cls, L.ch1, L.ch2, L.c1map, L.c2map, L.inv1, L.inv2,
                       L.x0, L.y0, L.slot_lx, L.slot_ly, L.slot_right, L.slot_up};
        for (int* p : ptrs)
            if (p) cudaFree(p);
    }
    cudaFree(s->ws_a); cudaFree(s->ws_b); cudaFree(s->ws_ring_a); cudaFree(s->ws_ring_b); cudaFree(s->ws_ye);
    cudaFree(s->d_info);
    if (s->fws) cudaFree(s->fws);
    if (s->xchg) cudaFree(s->xchg);
    if (s->ws_refine) cudaFree(s->ws_refine);
    if (s->ws_bsplit) cudaFree(s->ws_bsplit);
    cudaStreamDestroy(s->la_stream);
    cudaEventDestroy(s->la_ready);
    cudaEventDestroy(s->la_done);
    cudaEventDestroy(s->zg_fork);
    cudaEventDestroy(s->zg_join);
    delete s;
}

int g_lookahead_enabled = 1;       // A/B switch (FDFD_LOOKAHEAD=0): pivot-block look-ahead on chain levels
int g_lookahead_helper = 1;        // A/B switch (FDFD_LA_HELPER=0): the SMs reserved for the look-ahead chain rejoin the Schur update
int g_small_front_enabled = 1;     // A/B switch (fdfd_direct_set_small_fronts): 0 = generic path on every level

static int chunks_for(long long per_front_elems, long long nb) {
    // enough CTAs to fill the machine when there are few big fronts, one CTA per front otherwise
    long long want = (148LL * 8 + nb - 1) / nb;
    long long maxc = (per_front_elems + 2047) / 2048;
    long long c = want < maxc ? want : maxc;
    return (int)(c < 1 ? 1 : c);
}

// ---- recursive symmetric block inversion ---------------------------------------------------
// E = [A B^T; B D] (full symmetric storage, in place):
//   Ainv = A^-1 (recursion);  T = B Ainv;  S = D - T B^T;  Sinv = S^-1 (recursion);
//   E21 = -Sinv T;  E11 = Ainv - T^T E21;  E12 = E21^T.
// n^3 / 2 complex MACs, all of them in large-K GEMMs; the base case is the register-resident 64 x 64
// tile inverse.  T and T^T live in a workspace stack (they must survive the recursion into S).
static int inv_split(int n) { return ((n / 2 + 63) / 64) * 64; }
static size_t inv_ws_need(int n) {
    if (n <= 64) return 0;
    int n1 = inv_split(n), n2 = n - n1;
    return std::max(inv_ws_need(n1), 2 * (size_t)n1 * n2 + inv_ws_need(n2));
}

static int launch_transpose(const cplx* src, long long s_stride, int s_ld, cplx* dst, long long d_stride, int d_ld,
                            int rows, int cols, int mirror, long long nb, cudaStream_t st) {
    if (nb > 65535) FDFD_FAIL("transpose batch too large");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)nb);
    transpose_kernel<<<grid, 256, 0, st>>>(src, s_stride, s_ld, dst, d_stride, d_ld, rows, cols, mirror);
    ++g_fdfd_launches;
    FDFD_CHECK(cudaGetLastError());
    return 0;
}

// ---- block Gauss-Jordan inversion -----------------------------------------------------------
// The recursive scheme above is all large-K GEMMs but a long serial chain of tiny launches (64 for a 512 block:
// ~1.4 ms, which at the top of the tree -- and on every step of a distributed front -- sits on the critical path).
// Block Gauss-Jordan with 64-wide pivot tiles needs TWO launches per pivot tile: the register-resident tile inverse
// P = M_pp^-1 (in-tile pivoting, as before) and one update kernel over all 64 x 64 tiles, out of place (src -> dst,
// ping-pong, so no tile is read after it was overwritten):
//     dst_pp = P        dst_pj = P src_pj        dst_ip = -src_ip P        dst_ij = src_ij - src_ip (P src_pj)
// n^3 complex MACs instead of n^3 / 2 (the symmetry is not used), which is nothing next to the latency it removes:
// the two 64^3 products of a tile run on the tensor pipe (DMMA, 3M) out of shared memory.
__device__ __forceinline__ void gj_load_tile(cplx* dst, int ld, const cplx* src, int n, int r0, int c0, int tid) {
    for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        dst[r * ld + c] = (r0 + r < n && c0 + c < n) ? src[(size_t)(r0 + r) * n + c0 + c] : make_double2(0.0, 0.0);
    }
}
// acc = A (64 x 64, row-major, lda) * B (64 x 64, row-major, ldb); warp (wm, wn) owns rows 16 wm.., columns 32 wn..
__device__ __forceinline__ void gj_tile_mm(const cplx* As, int lda, const cplx* Bs, int ldb, int wm, int wn, int gq, int tq,
                                           cplx (&out)[2][4][2]) {
    double cr[2][4][2], ci[2][4][2], t3[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t3[i][j][0] = t3[i][j][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < 64; kk += 4) {
        cplx a[2], b[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) a[mt] = As[(wm * 16 + mt * 8 + gq) * lda + kk + tq];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[(kk + tq) * ldb + wn * 32 + nt * 8 + gq];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                dmma884(cr[mt][nt][0], cr[mt][nt][1], a[mt].x, b[nt].x);
                dmma884(ci[mt][nt][0], ci[mt][nt][1], a[mt].y, b[nt].y);
                dmma884(t3[mt][nt][0], t3[mt][nt][1], a[mt].x + a[mt].y, b[nt].x + b[nt].y);
            }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                out[mt][nt][e] = make_double2(cr[mt][nt][e] - ci[mt][nt][e], t3[mt][nt][e] - cr[mt][nt][e] - ci[mt][nt][e]);
}
#define GJ_LDA 68
#define GJ_LDB 66
__global__ void __launch_bounds__(256)
gj_update_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst, const cplx* __restrict__ Pinv, int n, int p, int tw) {
    extern __shared__ __align__(16) unsigned char gj_smem[];
    cplx* Ps = reinterpret_cast<cplx*>(gj_smem);        // P, A-operand layout (also read as a B operand)
    cplx* Ys = Ps + 64 * GJ_LDA;                        // src_ip, A operand
    cplx* Xs = Ys + 64 * GJ_LDA;                        // src_pj, then R = P src_pj, B operand
    const int ti = blockIdx.y, tj = blockIdx.x, tid = threadIdx.x;
    const long long b = blockIdx.z;
    const cplx* S = src + b * (long long)n * n;
    cplx* D = dst + b * (long long)n * n;
    const cplx* P = Pinv + b * 4096;
    const int warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1, gq = lane >> 2, tq = lane & 3;
    const int r0 = ti * 64, c0 = tj * 64, p0 = p * 64;
    for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        Ps[r * GJ_LDA + c] = (r < tw && c < tw) ? P[e] : make_double2(0.0, 0.0);
    }
    if (ti == p && tj == p) {                            // the pivot tile itself: dst = P
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += 256) {
            const int r = e >> 6, c = e & 63;
            if (r < tw && c < tw) D[(size_t)(p0 + r) * n + p0 + c] = Ps[r * GJ_LDA + c];
        }
        return;
    }
    if (tj != p) gj_load_tile(Xs, GJ_LDB, S, n, p0, c0, tid);
    if (ti != p) gj_load_tile(Ys, GJ_LDA, S, n, r0, p0, tid);
    __syncthreads();
    cplx acc[2][4][2];
    if (tj != p) {
        gj_tile_mm(Ps, GJ_LDA, Xs, GJ_LDB, wm, wn, gq, tq, acc);            // R = P src_pj
        if (ti != p) {
            __syncthreads();                                               // every warp is done reading src_pj
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        Xs[(wm * 16 + mt * 8 + gq) * GJ_LDB + wn * 32 + nt * 8 + 2 * tq + e] = acc[mt][nt][e];
            __syncthreads();
            gj_tile_mm(Ys, GJ_LDA, Xs, GJ_LDB, wm, wn, gq, tq, acc);        // src_ip R
        }
    } else {
        gj_tile_mm(Ys, GJ_LDA, Ps, GJ_LDA, wm, wn, gq, tq, acc);            // src_ip P
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row = r0 + wm * 16 + mt * 8 + gq;
        if (row >= n) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = c0 + wn * 32 + nt * 8 + 2 * tq + e;
                if (col >= n) continue;
                cplx v = acc[mt][nt][e];
                if (ti == p) { /* row tile: R itself */ }
                else if (tj == p) v = make_double2(-v.x, -v.y);
                else {
                    const cplx o = S[(size_t)row * n + col];
                    v = make_double2(o.x - v.x, o.y - v.y);
                }
                D[(size_t)row * n + col] = v;
            }
    }
}
static size_t gj_ws_need(int n) { return n <= 64 ? 0 : (size_t)n * n + 4096; }
int g_block_gj = 1;                // A/B switch (FDFD_BLOCK_GJ=0): block Gauss-Jordan (default) or the recursive inversion
thread_local int g_gj_on_lookahead = 0;   // set by the distributed fronts where the concurrent update is short (see distfront.cuh)
int g_dist_gj_group = 8;           // FDFD_DIST_GJ_GROUP: distributed fronts of at least this many ranks invert look-ahead blocks by block GJ (0: never)
int g_block_gj_max_tiles = 256;    // ... used while one step's grid (fronts x tiles) stays within a few waves of CTAs: it trades
                                   // flops (2x, in latency-bound 64^3 products, one CTA per SM) for a short launch chain, which
                                   // pays at the top of the tree and on distributed fronts, not on levels of many fronts

// E (packed [nb][n][n], full storage) <- E^-1 in place; ws holds nb (n^2 + 4096) entries
static int gj_invert_batch(NdSolver* s, cplx* E, int n, long long nb, cplx* ws, cudaStream_t st) {
    PhaseScope ph(PH_PIVOT, st);
    if (nb > 65535) FDFD_FAIL("block inversion batch too large");
    const int nt = (n + 63) / 64;
    cplx* W = ws;
    cplx* Pscr = ws + (size_t)nb * n * n;
    constexpr size_t sm = sizeof(cplx) * (2 * 64 * GJ_LDA + 64 * GJ_LDB);
    static bool attr = false;
    if (!attr) {
        FDFD_CHECK(cudaFuncSetAttribute(gj_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attr = true;
    }
    cplx *src = E, *dst = W;
    if (nt & 1) {                                      // an odd number of ping-pong steps must START in the workspace
        FDFD_CHECK(cudaMemcpyAsync(W, E, sizeof(cplx) * (size_t)nb * n * n, cudaMemcpyDeviceToDevice, st));
        src = W; dst = E;
    }
    for (int p = 0; p < nt; ++p) {
        const int tw = std::min(64, n - 64 * p);
        launch_tile_inverse(src + (size_t)p * 64 * (n + 1), (long long)n * n, n, tw, Pscr, 4096, 64, s->d_info, 0, nb, st);
        dim3 grid(nt, nt, (unsigned)nb);
        gj_update_kernel<<<grid, 256, sm, st>>>(src, dst, Pscr, n, p, tw);
        ++g_fdfd_launches;
        std::swap(src, dst);
    }
    FDFD_CHECK(cudaGetLastError());
    s->factor_flops += 8.0 * (double)nb * (double)n * n * n;
    return 0;
}

static int sym_invert_batch(NdSolver* s, cplx* E, long long sE, int ld, int n, long long nb, cplx* ws,
                            cudaStream_t st) {
    if (n <= 64) {
        PhaseScope ph(PH_PIVOT, st);
        launch_tile_inverse(E, sE, ld, n, E, sE, ld, s->d_info, 0, nb, st);
        FDFD_CHECK(cudaGetLastError());
        return 0;
    }
    // (not on the look-ahead stream: there the inversion shares the machine with a persistent GEMM that leaves it 4-16
    // SMs, and a step's 64-256 one-per-SM CTAs would crawl through them; measured: 5.8 -> 7.5 ms on the level it delays)
    if (g_block_gj && ld == n && sE == (long long)n * n && (st != s->la_stream || g_gj_on_lookahead) &&
        nb * (long long)((n + 63) / 64) * ((n + 63) / 64) <= g_block_gj_max_tiles)
        return gj_invert_batch(s, E, n, nb, ws, st);
    const int n1 = inv_split(n), n2 = n - n1;
    cplx *A = E, *B = E + (size_t)n1 * ld, *D = B + n1, *Bt = E + n1;
    cplx *T = ws, *Tt = ws + (size_t)nb * n1 * n2, *ws_next = ws + 2 * (size_t)nb * n1 * n2;
    const long long sT = (long long)n1 * n2;
    if (sym_invert_batch(s, A, sE, ld, n1, nb, ws, st)) return -1;
    GemmBatch g;
    g.batch = (int)nb;
    {   // T = B Ainv  (Ainv symmetric: NT form)
        PhaseScope ph(PH_ROWGEMM, st);
        g.transb = 1; g.lower = 0; g.mode = 0;
        g.A = B; g.sA = sE; g.lda = ld;
        g.B = A; g.sB = sE; g.ldb = ld;
        g.C = T; g.sC = sT; g.ldc = n1;
        g.M = n2; g.N = n1; g.K = n1;
        if (zgemm_batched(g, st)) return -1;
    }
    {   // S = D - T B^T  (lower tiles, then mirrored: the recursion wants full storage)
        PhaseScope ph(PH_UPDATE, st);
        g.transb = 1; g.lower = 1; g.mode = 1;
        g.A = T; g.sA = sT; g.lda = n1;
        g.B = B; g.sB = sE; g.ldb = ld;
        g.C = D; g.sC = sE; g.ldc = ld;
        g.M = n2; g.N = n2; g.K = n1;
        if (zgemm_batched(g, st)) return -1;
    }
    {
        PhaseScope ph(PH_COPY, st);
        if (launch_transpose(D, sE, ld, D, sE, ld, n2, n2, 1, nb, st)) return -1;
        if (launch_transpose(T, sT, n1, Tt, sT, n2, n2, n1, 0, nb, st)) return -1;
    }
    if (sym_invert_batch(s, D, sE, ld, n2, nb, ws_next, st)) return -1;
    {   // E21 = -Sinv T
        PhaseScope ph(PH_ROWGEMM, st);
        g.transb = 0; g.lower = 0; g.mode = 2;
        g.A = D; g.sA = sE; g.lda = ld;
        g.B = T; g.sB = sT; g.ldb = n1;
        g.C = B; g.sC = sE; g.ldc = ld;
        g.M = n2; g.N = n1; g.K = n2;
        if (zgemm_batched(g, st)) return -1;
    }
    {   // E11 = Ainv - T^T E21  (lower tiles + mirror)
        PhaseScope ph(PH_UPDATE, st);
        g.transb = 0; g.lower = 1; g.mode = 1;
        g.A = Tt; g.sA = sT; g.lda = n2;
        g.B = B; g.sB = sE; g.ldb = ld;
        g.C = A; g.sC = sE; g.ldc = ld;
        g.M = n1; g.N = n1; g.K = n2;
        if (zgemm_batched(g, st)) return -1;
    }
    {
        PhaseScope ph(PH_COPY, st);
        if (launch_transpose(A, sE, ld, A, sE, ld, n1, n1, 1, nb, st)) return -1;
        if (launch_transpose(B, sE, ld, Bt, sE, ld, n2, n1, 0, nb, st)) return -1;      // E12 = E21^T
    }
    s->factor_flops += 8.0 * (double)nb * ((double)n2 * n1 * n1 + 0.5 * (double)n2 * n2 * n1 +
                                           (double)n2 * n2 * n1 + 0.5 * (double)n1 * n1 * n2);
    return 0;
}

#include "distfront.cuh"

// one arena for all transient factorisation buffers: two ping-pong front batches plus the workspace
// stack of the block inversion, sized for the largest level; allocated once and kept
static int ensure_factor_workspace(NdSolver* s) {
    size_t maxF = 0, maxW = 0, maxX = 0;
    for (size_t li = 0; li < s->levels.size(); ++li) {
        NdLevel& L = s->levels[li];
        // one extra slot where the parent level receives its second child from another rank
        // (levels factorised in place live in the buffer of the level that started their chain)
        size_t lj = li + 1;
        while (lj < s->levels.size() && s->levels[lj].inplace) ++lj;
        const bool extra = lj < s->levels.size() && s->levels[lj].recv_from >= 0;
        const size_t nb = L.nb + (extra ? 1 : 0), nmax = L.nmax;
        maxF = std::max(maxF, nb * nmax * nmax);
        maxW = std::max(maxW, (size_t)L.nb * std::max(inv_ws_need(L.kmax), gj_ws_need(L.kmax)));
        if (L.send_to >= 0 || L.recv_from >= 0) maxX = std::max(maxX, (size_t)L.child_mmax * L.child_mmax);
    }
    for (const NdDistFront* f : s->dist) maxW = std::max(maxW, std::max(inv_ws_need(f->kmax_step), gj_ws_need(f->kmax_step)));
    if (maxX > s->xchg_cap) {
        if (s->xchg) cudaFree(s->xchg);
        s->xchg = nullptr;
        s->xchg_cap = 0;
        FDFD_CHECK(cudaMalloc(&s->xchg, sizeof(cplx) * maxX));
        s->xchg_cap = maxX;
    }
    size_t need = 2 * maxF + maxW;
    if (need > s->fws_cap) {
        if (s->fws) cudaFree(s->fws);
        s->fws = nullptr;
        s->fws_cap = 0;
        FDFD_CHECK(cudaMalloc(&s->fws, sizeof(cplx) * std::max(need, (size_t)1)));
        s->fws_cap = need;
    }
    s->fws_F[0] = s->fws;
    s->fws_F[1] = s->fws + maxF;
    s->fws_W = s->fws + 2 * maxF;
    return 0;
}

int nd_factor(NdSolver* s, const FdfdOp* op, bool defer_check) {
    if (s->levels.empty()) FDFD_FAIL("no levels in the plan");
    if (op->nx != s->nx || op->ny != s->ny) FDFD_FAIL("operator / plan shape mismatch");
    cudaStream_t st = op->stream;
    s->factored = false;
    if (ensure_factor_workspace(s)) return -1;
    {
        const char* e = getenv("FDFD_LOOKAHEAD");
        g_lookahead_enabled = !(e && e[0] == '0');
        e = getenv("FDFD_LA_HELPER");
        g_lookahead_helper = !(e && e[0] == '0');
        e = getenv("FDFD_BLOCK_GJ");
        g_block_gj = !(e && e[0] == '0');
        e = getenv("FDFD_DIST_GJ_GROUP");
        if (e) g_dist_gj_group = atoi(e);
        e = getenv("FDFD_BLOCK_GJ_MAXTILES");
        if (e && atoi(e) > 0) g_block_gj_max_tiles = atoi(e);
    }
    FDFD_CHECK(cudaMemsetAsync(s->d_info, 0, sizeof(int), st));
    // the previous level's Schur blocks: batch base, first ring entry at (prev_k, prev_k), leading dimension
    // prev_n, batch stride prev_stride, ring size prev_m
    cplx* Fprev = nullptr;
    bool lookahead_done = false;
    int prev_k = 0, prev_n = 0, prev_m = 0, cur = 1;
    long long prev_stride = 0;
    s->factor_bytes = 0;
    s->factor_flops = 0;
    for (size_t li = 0; li < s->levels.size(); ++li) {
        NdLevel& L = s->levels[li];
        g_phase_timing.level = (int)li;
        int dj = 0;
        if (NdDistFront* df = dist_at_level(s, (int)li, &dj)) {
            // a front shared by several ranks: levels li .. li + nsteps - 1 are its elimination steps
            // (the local child's batch is padded to its level's largest shape class: ring entries keep their positions)
            if (dj == 0 && (!Fprev || prev_m < df->mc[df->cidx])) FDFD_FAIL("distributed front: local child size mismatch");
            const cplx* cbase = dj == 0 ? Fprev + (size_t)prev_k * prev_n + prev_k : nullptr;
            if (dist_factor(s, df, dj > 0 ? s->dist[dj - 1] : nullptr, cbase, prev_n, st)) return -1;
            li += df->nsteps - 1;
            continue;
        }
        const long long nb = L.nb;
        const int nmax = L.nmax, kmax = L.kmax, mmax = L.mmax;
        // A chain level that continues the elimination of the same fronts works IN PLACE on the previous Schur
        // blocks (its front is that block, slot for slot); every other level assembles into the other buffer.
        cplx* F;
        int ld;
        long long fstride;
        if (L.inplace && nb > 0) {
            F = Fprev + (size_t)prev_k * prev_n + prev_k;
            ld = prev_n;
            fstride = prev_stride;
        } else {
            cur ^= 1;
            F = s->fws_F[cur];
            ld = nmax;
            fstride = (long long)nmax * nmax;
        }
        if (L.send_to >= 0 || L.recv_from >= 0) {
            // sharded tree: the second child's Schur block crosses NVLink as a packed m x m square and lands
            // in slot 1 of the child batch, in the layout the local child has (ch2 of this level points there)
            if (!s->comm) FDFD_FAIL("sharded elimination plan without a communicator (fdfd_direct_set_comm)");
            const size_t m = (size_t)prev_m, pitch_f = sizeof(cplx) * prev_n, pitch_x = sizeof(cplx) * m;
            if ((int)m != L.child_mmax || !Fprev) FDFD_FAIL("sharded plan: child block size mismatch");
            PhaseScope ph(PH_COPY, st);
            if (L.send_to >= 0) {
                FDFD_CHECK(cudaMemcpy2DAsync(s->xchg, pitch_x, Fprev + (size_t)prev_k * prev_n + prev_k, pitch_f, pitch_x, m,
                                             cudaMemcpyDeviceToDevice, st));
                if (comm_send(s->comm, s->xchg, 2 * m * m, L.send_to, st)) return -1;
            } else {
                if (comm_recv(s->comm, s->xchg, 2 * m * m, L.recv_from, st)) return -1;
                FDFD_CHECK(cudaMemcpy2DAsync(Fprev + prev_stride + (size_t)prev_k * prev_n + prev_k, pitch_f,
                                             s->xchg, pitch_x, pitch_x, m, cudaMemcpyDeviceToDevice, st));
            }
        }
        if (nb == 0) continue;          // this rank's part of the tree ended below this level
        // factor storage is allocated on the first factorisation and reused afterwards
        if (!L.Einv) FDFD_CHECK(cudaMalloc(&L.Einv, sizeof(cplx) * (size_t)nb * kmax * kmax));
        if (mmax > 0 && !L.G) FDFD_CHECK(cudaMalloc(&L.G, sizeof(cplx) * (size_t)nb * mmax * kmax));
        if (small_front_ok(kmax, mmax, L.kind == 0) && mmax > 0 && g_small_front_enabled) {
            // bottom of the tree: one fused kernel per level, compact Schur blocks handed to the parent
            SmallFrontArgs a;
            a.kind = L.kind; a.kmax = kmax; a.mmax = mmax; a.nx = s->nx; a.ny = s->ny;
            a.cls = L.cls; a.k_cls = L.k_cls;
            a.x0 = L.x0; a.y0 = L.y0; a.slot_lx = L.slot_lx; a.slot_ly = L.slot_ly;
            a.slot_right = L.slot_right; a.slot_up = L.slot_up;
            a.ch1 = L.ch1; a.ch2 = L.ch2; a.inv1 = L.inv1; a.inv2 = L.inv2;
            a.planes = op->planes; a.isxf = op->isxf; a.isyf = op->isyf;
            a.Sc = Fprev; a.sc = prev_stride; a.kc = prev_k; a.nc = prev_n;
            a.Einv = L.Einv; a.G = L.G; a.S = F; a.info = s->d_info;
            // warps per front by front size, fronts per CTA by what fits (<= 8 groups, <= 200 KB, <= 512 threads)
            const size_t fsm = small_front_cplx(kmax, mmax, L.kind == 0);
            const int wpf = nmax <= 40 ? 1 : (nmax <= 64 ? 4 : 8);
            int fpc = (int)std::min<size_t>(8, (200 * 1024) / (fsm * sizeof(cplx)));
            fpc = std::max(1, std::min(fpc, 512 / (32 * wpf)));
            const size_t smem = fsm * sizeof(cplx) * fpc;
            const int threads = fpc * wpf * 32;
            typedef void (*SfKern)(SmallFrontArgs, long long, int);
            auto pick = [&](auto kmax_c, auto kt_c, auto mt_c) -> SfKern {
                constexpr int KM = decltype(kmax_c)::value, KT = decltype(kt_c)::value, MT = decltype(mt_c)::value;
                return wpf == 1 ? small_front_kernel<KM, KT, MT, 1> : wpf == 2 ? small_front_kernel<KM, KT, MT, 2>
                       : wpf == 4 ? small_front_kernel<KM, KT, MT, 4> : small_front_kernel<KM, KT, MT, 8>;
            };
#define SF_PICK(KM, KT, MT) pick(std::integral_constant<int, KM>(), std::integral_constant<int, KT>(), std::integral_constant<int, MT>())
            SfKern kern = kmax <= 3 ? SF_PICK(3, 0, 0) : kmax <= 7 ? SF_PICK(7, 0, 0) : kmax <= 9 ? SF_PICK(9, 0, 0)
                          : kmax <= 12 ? SF_PICK(12, 0, 0) : SF_PICK(16, 0, 0);
            // the bottom levels of a power-of-two grid (4-cell leaves): sizes as compile-time constants
            if (kmax == 9 && mmax == 16) kern = SF_PICK(9, 9, 16);
            else if (kmax == 3 && mmax == 24) kern = SF_PICK(3, 3, 24);
            else if (kmax == 7 && mmax == 32) kern = SF_PICK(7, 7, 32);
            else if (kmax == 7 && mmax == 48) kern = SF_PICK(7, 7, 48);
            else if (kmax == 15 && mmax == 64) kern = SF_PICK(16, 15, 64);
            else if (kmax == 15 && mmax == 96) kern = SF_PICK(16, 15, 96);
#undef SF_PICK
            if (smem > 48 * 1024)
                FDFD_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            {
                PhaseScope ph(PH_SMALL, st);
                kern<<<(unsigned)((nb + fpc - 1) / fpc), threads, smem, st>>>(a, nb, (int)fsm);
                ++g_fdfd_launches;
            }
            FDFD_CHECK(cudaGetLastError());
            s->factor_flops += 8.0 * (double)nb * ((double)kmax * kmax * kmax + (double)mmax * kmax * kmax +
                                                   0.5 * (double)mmax * mmax * kmax);
            s->factor_bytes += sizeof(cplx) * ((size_t)nb * kmax * kmax + (size_t)nb * mmax * kmax);
            Fprev = F;
            prev_k = 0;
            prev_n = mmax;
            prev_m = mmax;
            prev_stride = (long long)mmax * mmax;
            continue;
        }
        if (!L.inplace) {
            PhaseScope ph(PH_ASSEMBLE, st);
            if (L.kind == 0) {
                leaf_assemble_kernel<<<(unsigned)nb, 64, 0, st>>>(F, op->planes, op->isxf, op->isyf, L.cls, L.k_cls,
                                                                  L.x0, L.y0, L.slot_lx, L.slot_ly, L.slot_right,
                                                                  L.slot_up, kmax, nmax, s->nx, s->ny);
            } else {
                int chunks = (nmax + ASM_ROWS - 1) / ASM_ROWS;
                merge_assemble_kernel<<<(unsigned)(nb * chunks), 256, 0, st>>>(F, Fprev, L.cls, L.k_cls, L.ch1, L.ch2,
                                                                               L.inv1, L.inv2, kmax, nmax, prev_k,
                                                                               prev_n, prev_stride, chunks);
            }
            ++g_fdfd_launches;
        }
        FDFD_CHECK(cudaGetLastError());
        // ---- Einv = F_EE^-1
        if (lookahead_done) {
            // inverted on the side stream while the previous level's Schur update ran
            FDFD_CHECK(cudaStreamWaitEvent(st, s->la_done, 0));
            lookahead_done = false;
        } else if (kmax <= 64) {
            PhaseScope ph(PH_PIVOT, st);
            launch_tile_inverse(F, fstride, ld, kmax, L.Einv, (long long)kmax * kmax, kmax, s->d_info, 1, nb, st);
        } else {
            {
                PhaseScope ph(PH_EXTRACT, st);
                int chunks = chunks_for((long long)kmax * kmax, nb);
                sym_expand_kernel<<<(unsigned)(nb * chunks), 256, 0, st>>>(F, L.Einv, kmax, ld, fstride, chunks);
                ++g_fdfd_launches;
            }
            if (sym_invert_batch(s, L.Einv, (long long)kmax * kmax, kmax, kmax, nb, s->fws_W, st)) return -1;
        }
        FDFD_CHECK(cudaGetLastError());
        if (mmax > 0) {
            GemmBatch g;
            // ---- G = F_RE Einv = F_RE Einv^T (Einv is symmetric: both operands k-contiguous)
            g.transb = 1; g.lower = 0; g.mode = 0; g.batch = (int)nb;
            g.A = F + (size_t)kmax * ld; g.sA = fstride; g.lda = ld;
            g.B = L.Einv; g.sB = (long long)kmax * kmax; g.ldb = kmax;
            g.C = L.G; g.sC = (long long)mmax * kmax; g.ldc = kmax;
            g.M = mmax; g.N = kmax; g.K = kmax;
            {
                PhaseScope ph(PH_GGEMM, st);
                if (zgemm_batched(g, st)) return -1;
            }
            // ---- S = F_RR - G F_RE^T, lower tiles only, in place (the parent assembles from it)
            g.transb = 1; g.lower = 1; g.mode = 1;
            g.A = L.G; g.sA = (long long)mmax * kmax; g.lda = kmax;
            g.B = F + (size_t)kmax * ld; g.sB = fstride; g.ldb = ld;
            g.C = F + (size_t)kmax * ld + kmax; g.sC = fstride; g.ldc = ld;
            g.M = mmax; g.N = mmax; g.K = kmax;
            // LOOK-AHEAD: when the next level continues this front in place (a chain step), its pivot block is the
            // leading k1 x k1 corner of this Schur block.  That corner is updated first, then inverted on a
            // high-priority side stream (the latency-bound recursive block inversion, ~1.5-3 ms) while the rest of
            // the update -- the bulk of the level -- runs on all but a few SMs of the main stream.
            NdLevel* Nx = li + 1 < s->levels.size() ? &s->levels[li + 1] : nullptr;
            const bool la = g_lookahead_enabled && Nx && Nx->inplace && Nx->nb == nb && Nx->kmax > 64 && Nx->kmax < mmax &&
                            nb <= 16 && !dist_at_level(s, (int)li + 1, nullptr);
            if (la) {
                const int k1 = Nx->kmax;
                if (!Nx->Einv) FDFD_CHECK(cudaMalloc(&Nx->Einv, sizeof(cplx) * (size_t)nb * k1 * k1));
                GemmBatch c = g;
                {   // corner: rows and columns [0, k1)
                    PhaseScope ph(PH_SCHUR, st);
                    c.M = k1; c.N = k1;
                    if (zgemm_batched(c, st)) return -1;
                }
                FDFD_CHECK(cudaEventRecord(s->la_ready, st));
                FDFD_CHECK(cudaStreamWaitEvent(s->la_stream, s->la_ready, 0));
                {
                    cplx* Fn = F + (size_t)kmax * ld + kmax;           // the next level's front (a view)
                    int chunks = chunks_for((long long)k1 * k1, nb);
                    sym_expand_kernel<<<(unsigned)(nb * chunks), 256, 0, s->la_stream>>>(Fn, Nx->Einv, k1, ld, fstride, chunks);
                    ++g_fdfd_launches;
                    const bool timing = g_phase_timing.on;
                    g_phase_timing.on = false;                          // phase events belong to the main stream
                    int rc = sym_invert_batch(s, Nx->Einv, (long long)k1 * k1, k1, k1, nb, s->fws_W, s->la_stream);
                    g_phase_timing.on = timing;
                    if (rc) return -1;
                }
                FDFD_CHECK(cudaEventRecord(s->la_done, s->la_stream));
                lookahead_done = true;
                const int reserve = nb <= 4 ? 4 : (nb <= 8 ? 8 : 16);
                g_zgemm_max_ctas = 148 - reserve;
                PhaseScope ph(PH_SCHUR, st);
                // rows [k1, m): the rectangle left of the corner's columns, then the lower square
                c = g;
                c.lower = 0;
                c.A = g.A + (size_t)k1 * kmax; c.C = g.C + (size_t)k1 * ld;
                c.M = mmax - k1; c.N = k1;
                int rc = zgemm_batched(c, st);
                c.lower = 1;
                c.B = g.B + (size_t)k1 * ld; c.C = g.C + (size_t)k1 * ld + k1;
                c.M = mmax - k1; c.N = mmax - k1;
                // the SMs left to the look-ahead chain join this update once that chain is through (helper launch on
                // the look-ahead stream, tiles handed out dynamically)
                ZgemmHelper helper = {s->la_stream, reserve, s->zg_fork, s->zg_join};
                if (g_lookahead_helper) g_zgemm_helper = &helper;
                if (!rc) rc = zgemm_batched(c, st);
                g_zgemm_helper = nullptr;
                g_zgemm_max_ctas = 148;
                if (rc) return -1;
            } else {
                PhaseScope ph(PH_SCHUR, st);
                if (zgemm_batched(g, st)) return -1;
            }
            s->factor_flops += 8.0 * (double)nb * ((double)mmax * kmax * kmax + 0.5 * (double)mmax * mmax * kmax);
        }
        s->factor_bytes += sizeof(cplx) * ((size_t)nb * kmax * kmax + (size_t)nb * mmax * kmax);
        Fprev = F;
        prev_k = kmax;
        prev_n = ld;
        prev_m = mmax;
        prev_stride = fstride;
    }
    g_phase_timing.level = -1;
    if (s->comm && s->comm->world > 1 && comm_allreduce_max_i32(s->comm, s->d_info, 1, st)) return -1;
    s->factored = true;
    s->fact_op = op;
    s->fact_version = op->version;
    if (defer_check) return 0;          // the caller keeps queueing work and calls nd_factor_check at its own sync point
    return nd_factor_check(s, op);
}

// waits for the factorisation and reads the singular-pivot flag
int nd_factor_check(NdSolver* s, const FdfdOp* op) {
    int info = 0;
    FDFD_CHECK(cudaMemcpyAsync(&info, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, op->stream));
    FDFD_CHECK(cudaStreamSynchronize(op->stream));
    if (info) {
        s->factored = false;
        FDFD_FAIL("direct solver: a pivot block is numerically singular (no inter-block pivoting)");
    }
    return 0;
}

template <int NR>
static int nd_solve_chunk(NdSolver* s, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nr_act) {
    cudaStream_t st = op->stream;
    const size_t nlev = s->levels.size();
    // workspace sizing
    size_t vec_need = 0, ring_need = 1, ye_need = 0;
    for (size_t li = 0; li < nlev; ++li) {
        NdLevel& L = s->levels[li];
        const size_t slots = (size_t)L.nb + ((li + 1 < nlev && s->levels[li + 1].recv_from >= 0) ? 1 : 0);
        vec_need = std::max(vec_need, slots * L.nmax * NR);
        ring_need = std::max(ring_need, slots * std::max(L.mmax, 1) * NR);
        L.ye_off = ye_need;
        ye_need += (size_t)L.nb * L.kmax * NR;
    }
    ye_need = std::max(ye_need, (size_t)1);
    // grow-only workspaces: pointer and capacity are reset BEFORE the new allocation, so a failed cudaMalloc
    // (out of memory is realistic next to tens of GB of factors) leaves a consistent, empty workspace behind
    if (vec_need > s->ws_vec_cap) {
        cudaFree(s->ws_a); cudaFree(s->ws_b);
        s->ws_a = s->ws_b = nullptr;
        s->ws_vec_cap = 0;
        FDFD_CHECK(cudaMalloc(&s->ws_a, sizeof(cplx) * vec_need));
        FDFD_CHECK(cudaMalloc(&s->ws_b, sizeof(cplx) * vec_need));
        s->ws_vec_cap = vec_need;
    }
    if (ring_need > s->ws_ring_cap) {
        cudaFree(s->ws_ring_a); cudaFree(s->ws_ring_b);
        s->ws_ring_a = s->ws_ring_b = nullptr;
        s->ws_ring_cap = 0;
        FDFD_CHECK(cudaMalloc(&s->ws_ring_a, sizeof(cplx) * ring_need));
        FDFD_CHECK(cudaMalloc(&s->ws_ring_b, sizeof(cplx) * ring_need));
        s->ws_ring_cap = ring_need;
    }
    if (ye_need > s->ws_ye_cap) {
        cudaFree(s->ws_ye);
        s->ws_ye = nullptr;
        s->ws_ye_cap = 0;
        FDFD_CHECK(cudaMalloc(&s->ws_ye, sizeof(cplx) * ye_need));
        s->ws_ye_cap = ye_need;
    }
    cplx *f = s->ws_a, *ring_prev = s->ws_ring_a, *ring_cur = s->ws_ring_b;
    // ---- forward (leaves -> root)
    {
    for (size_t li = 0; li < nlev; ++li) {
        NdLevel& L = s->levels[li];
        g_phase_timing.level = (int)li;
        PhaseScope phf(PH_SOLVE_FWD, st);
        int dj = 0;
        if (NdDistFront* df = dist_at_level(s, (int)li, &dj)) {
            const cplx* my_ring = dj == 0 ? ring_prev : s->dist[dj - 1]->vec + (size_t)s->dist[dj - 1]->kfull * NR;
            if (dist_forward<NR>(s, df, my_ring, st)) return -1;
            li += df->nsteps - 1;
            continue;
        }
        const long long nb = L.nb;
        // sharded tree: the second child's ring right-hand side arrives in slot 1 of the child batch
        const size_t ring_cnt = 2 * (size_t)L.child_mmax * NR;
        if (L.send_to >= 0 && comm_send(s->comm, ring_prev, ring_cnt, L.send_to, st)) return -1;
        if (L.recv_from >= 0 && comm_recv(s->comm, ring_prev + (size_t)L.child_mmax * NR, ring_cnt, L.recv_from, st))
            return -1;
        if (nb == 0) continue;
        long long tot = nb * L.nmax;
        if (L.kind == 0)
            { leaf_gather_kernel<NR><<<ceil_div(tot, 128), 128, 0, st>>>(f, d_b, op->isxf, op->isyf, L.cls, L.x0, L.y0, L.slot_lx, L.slot_ly,
                                                                       L.slot_right, L.nmax, s->nx, s->ny, nr_act, nb); ++g_fdfd_launches; }
        else
            { merge_gather_kernel<NR><<<ceil_div(tot, 128), 128, 0, st>>>(f, ring_prev, L.cls, L.ch1, L.ch2, L.inv1,
                                                                        L.inv2, L.nmax, L.child_mmax, nb); ++g_fdfd_launches; }
        if (L.kmax <= 16) {
            { forward_mv_thread_kernel<NR><<<ceil_div(tot, 256), 256, 0, st>>>(L.Einv, L.G, f, s->ws_ye + L.ye_off, ring_cur,
                                                                             L.kmax, L.mmax, L.nmax, nb); ++g_fdfd_launches; }
        } else if constexpr (NR >= 8) {
            // many right-hand sides: the factor blocks stream once through the tensor-pipe kernel (mrhs.cuh)
            MrhsArgs a{};
            a.batch = nb; a.lda = L.kmax; a.K = L.kmax;
            a.V = f; a.sV = (long long)L.nmax * NR;
            a.A = L.Einv; a.sA = (long long)L.kmax * L.kmax; a.M = L.kmax;
            a.D = nullptr; a.C = s->ws_ye + L.ye_off; a.sC = (long long)L.kmax * NR;
            if (mrhs_launch<NR, false>(a, &s->ws_bsplit, &s->ws_bsplit_cap, st)) return -1;
            if (L.mmax > 0) {
                a.A = L.G; a.sA = (long long)L.mmax * L.kmax; a.M = L.mmax;
                a.D = f + (size_t)L.kmax * NR; a.sD = (long long)L.nmax * NR;
                a.C = ring_cur; a.sC = (long long)L.mmax * NR;
                if (mrhs_launch<NR, false>(a, &s->ws_bsplit, &s->ws_bsplit_cap, st)) return -1;
            }
        } else {
            { forward_mv_kernel<NR><<<ceil_div(tot * 32, 256), 256, 0, st>>>(L.Einv, L.G, f, s->ws_ye + L.ye_off, ring_cur,
                                                                           L.kmax, L.mmax, L.nmax, nb); ++g_fdfd_launches; }
        }
        FDFD_CHECK(cudaGetLastError());
        std::swap(ring_prev, ring_cur);
    }
    }
    // ---- backward (root -> leaves)
    cplx *u = s->ws_a, *u_par = s->ws_b;
    for (size_t li = nlev; li-- > 0;) {
        NdLevel& L = s->levels[li];
        g_phase_timing.level = (int)li;
        PhaseScope phb(PH_SOLVE_BWD, st);
        {
            bool handled = false;
            for (size_t dj = 0; dj < s->dist.size() && !handled; ++dj) {
                NdDistFront* df = s->dist[dj];
                if ((size_t)(df->level0 + df->nsteps - 1) != li) continue;
                if (dist_backward<NR>(s, df, dj + 1 < s->dist.size() ? s->dist[dj + 1] : nullptr, st)) return -1;
                li = (size_t)df->level0;          // the loop's decrement moves on to the level below the front
                handled = true;
            }
            if (handled) continue;
        }
        const long long nb = L.nb;
        const NdLevel* Pp = li + 1 < nlev ? &s->levels[li + 1] : nullptr;
        const bool remote_child = Pp && Pp->recv_from >= 0, from_parent = Pp && Pp->send_to >= 0;
        const size_t vec_cnt = (size_t)L.nmax * NR;
        if (nb + (remote_child ? 1 : 0) == 0) continue;
        FDFD_CHECK(cudaMemsetAsync(u, 0, sizeof(cplx) * (size_t)(nb + (remote_child ? 1 : 0)) * vec_cnt, st));
        if (Pp && Pp->nb > 0) {
            const NdLevel& P = *Pp;
            long long tot = (long long)P.nb * 2 * P.child_mmax;
            { child_scatter_kernel<NR><<<ceil_div(tot, 128), 128, 0, st>>>(u, u_par, P.cls, P.ch1, P.ch2, P.c1map, P.c2map,
                                                                         P.nmax, P.child_mmax, L.kmax, L.nmax, P.nb); ++g_fdfd_launches; }
        }
        if (!s->dist.empty() && (size_t)s->dist[0]->level0 == li + 1 && nb > 0) {
            // my local front is a child of the first distributed front: its ring solution comes out of that front's vector
            const NdDistFront* df = s->dist[0];
            { dist_child_ring_kernel<NR><<<ceil_div(df->mc[df->cidx], 128), 128, 0, st>>>(u, df->vec, df->d_cmap_mine,
                                                                                       df->mc[df->cidx], L.kmax); ++g_fdfd_launches; }
        }
        // sharded tree: the remote child's ring solution (slot 1) goes back to the rank that owns it
        if (remote_child && comm_send(s->comm, u + vec_cnt, 2 * vec_cnt, Pp->recv_from, st)) return -1;
        if (from_parent && comm_recv(s->comm, u, 2 * vec_cnt, Pp->send_to, st)) return -1;
        if (nb == 0) { std::swap(u, u_par); continue; }
        const int ecta = ceil_div(nb * L.kmax, 32);
        int nsplit = std::min(592 / ecta, L.mmax / 64);
        if (L.kmax <= 16 && L.mmax <= 128) {
            { backward_mvt_thread_kernel<NR><<<ceil_div(nb * L.kmax, 128), 128, 0, st>>>(L.G, s->ws_ye + L.ye_off, u, L.kmax,
                                                                                       L.mmax, L.nmax, nb); ++g_fdfd_launches; }
        } else if constexpr (NR >= 8) {
            // u_E = yE - G^T u_R for all NR right-hand sides at once on the tensor pipe (mrhs.cuh)
            MrhsArgs a{};
            a.batch = nb; a.lda = L.kmax; a.M = L.kmax; a.K = L.mmax;
            a.A = L.G; a.sA = (long long)L.mmax * L.kmax;
            a.V = u + (size_t)L.kmax * NR; a.sV = (long long)L.nmax * NR;
            a.D = s->ws_ye + L.ye_off; a.sD = (long long)L.kmax * NR;
            a.C = u; a.sC = (long long)L.nmax * NR;
            if (L.mmax > 0) {
                if (mrhs_launch<NR, true>(a, &s->ws_bsplit, &s->ws_bsplit_cap, st)) return -1;
            } else {
                FDFD_CHECK(cudaMemcpy2DAsync(u, sizeof(cplx) * L.nmax * NR, s->ws_ye + L.ye_off, sizeof(cplx) * L.kmax * NR,
                                             sizeof(cplx) * L.kmax * NR, nb, cudaMemcpyDeviceToDevice, st));
            }
        } else if (nsplit >= 2) {
            // few fronts, long rings: also split the ring rows over CTAs (two-pass, fixed summation order)
            const size_t need = (size_t)nsplit * nb * L.kmax * NR;
            if (need > s->ws_bsplit_cap) {
                if (s->ws_bsplit) cudaFree(s->ws_bsplit);
                s->ws_bsplit = nullptr;
                s->ws_bsplit_cap = 0;
                FDFD_CHECK(cudaMalloc(&s->ws_bsplit, sizeof(cplx) * need));
                s->ws_bsplit_cap = need;
            }
            const int rows_per = ceil_div(L.mmax, nsplit);
            if constexpr (NR < 8) {
                dim3 grid(ecta, nsplit);
                { backward_mvt_split_kernel<NR><<<grid, 256, 0, st>>>(L.G, u, s->ws_bsplit, L.kmax, L.mmax, L.nmax, nb, rows_per); ++g_fdfd_launches; }
                { backward_mvt_reduce_kernel<NR><<<ceil_div(nb * L.kmax, 128), 128, 0, st>>>(s->ws_bsplit, s->ws_ye + L.ye_off, u,
                                                                                              L.kmax, L.nmax, nb, nsplit); ++g_fdfd_launches; }
            }
        } else {
            if constexpr (NR < 8) {
                { backward_mvt_kernel<NR><<<ecta, 256, 0, st>>>(L.G, s->ws_ye + L.ye_off, u, L.kmax,
                                                                L.mmax, L.nmax, nb); ++g_fdfd_launches; }
            }
        }
        if (L.kind == 0) {
            long long tot = nb * L.nmax;
            { leaf_scatter_kernel<NR><<<ceil_div(tot, 128), 128, 0, st>>>(u, d_x, L.cls, L.x0, L.y0, L.slot_lx, L.slot_ly,
                                                                        L.slot_right, L.nmax, s->nx, s->ny, nr_act, nb); ++g_fdfd_launches; }
        }
        FDFD_CHECK(cudaGetLastError());
        std::swap(u, u_par);
    }
    g_phase_timing.level = -1;
    return 0;
}

int nd_solve(NdSolver* s, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs) {
    if (!s->factored) FDFD_FAIL("nd_solve called before nd_factor");
    const size_t n = (size_t)s->nx * s->ny;
    const bool sharded = s->comm && s->comm->world > 1;
    // sharded tree: every cell is written by exactly one rank; the sum over ranks is the whole field
    if (sharded) FDFD_CHECK(cudaMemsetAsync(d_x, 0, sizeof(cplx) * n * nrhs, op->stream));
    for (int j0 = 0; j0 < nrhs;) {
        int rem = nrhs - j0, rc;
        if (rem >= 16) { rc = 16; if (nd_solve_chunk<16>(s, op, d_b + j0 * n, d_x + j0 * n, 16)) return -1; }
        else if (rem >= 8) { rc = 8; if (nd_solve_chunk<8>(s, op, d_b + j0 * n, d_x + j0 * n, 8)) return -1; }
        else if (rem > 2) { rc = rem < 4 ? rem : 4; if (nd_solve_chunk<4>(s, op, d_b + j0 * n, d_x + j0 * n, rc)) return -1; }
        else if (rem == 2) { rc = 2; if (nd_solve_chunk<2>(s, op, d_b + j0 * n, d_x + j0 * n, 2)) return -1; }
        else { rc = 1; if (nd_solve_chunk<1>(s, op, d_b + j0 * n, d_x + j0 * n, 1)) return -1; }
        j0 += rc;
    }
    if (sharded && comm_allreduce_sum(s->comm, d_x, 2 * n * nrhs, op->stream)) return -1;
    return 0;
}
