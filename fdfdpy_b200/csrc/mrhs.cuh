// Substitution with MANY right-hand sides (NR = 8 or 16 per pass) on the FP64 tensor pipe.
//
// With one right-hand side the triangular-solve phase of the multifrontal method is a stream of matrix-vector
// products (forward_mv / backward_mvt in direct.cu): every factor entry is read once and used once, HBM-bound.  With
// 16 right-hand sides riding along (BASELINE config 4: one factorisation reused by 16 sources per omega; the
// reference's README to-do) every factor entry feeds 16 complex multiply-adds = 96 tensor flops (3M) per 16 bytes,
// which is the FP64 tensor pipe's rate at the HBM rate -- so the product has to run on DMMA, as a skinny GEMM
//        C [M x NR]  =  D - A' [M x K] V [K x NR]            A' = A (forward) or A^T (backward, A = G stored [m][k])
// whose A operand (Einv or G of one front) streams through shared memory exactly once per pass for all NR columns.
// V / D / C are front vectors in the solver's [slot][NR] layout.  One CTA = 128 rows x NR columns, 8 warps of
// 16 rows, cp.async ring of 3 stages over K in steps of 16, 3M complex products (zgemm.cuh).  Few-front levels with
// long K (the backward product at the top of the tree) are split over K into gridDim.z slices that write partial
// sums, added in a fixed order by mrhs_reduce_kernel.
#pragma once
#include "zgemm.cuh"

struct MrhsArgs {
    const cplx* A; long long sA; int lda;     // factor block of front b: A + b * sA
    const cplx* V; long long sV;              // [K][NR]
    const cplx* D; long long sD;              // [M][NR] or nullptr (C = A' V)
    cplx* C; long long sC;                    // [M][NR]
    int M, K;
    int kslices, kper;                        // split of K (kper = K values per slice, multiple of 16)
    cplx* part;                               // kslices > 1: partial sums [slice][batch][M][NR]
    long long batch;
};

template <int NR, bool TRANS, int BM, int STAGES>
__global__ void __launch_bounds__(BM * 2)
mrhs_dmma_kernel(MrhsArgs a) {
    constexpr int BK = 16, NT = NR / 8, NTHR = BM * 2;          // one warp per 16 rows
    constexpr int LDA = TRANS ? BM + 2 : BK + 4;                 // conflict-free fragment reads (see zgemm.cuh)
    constexpr int A_ELEMS = TRANS ? BK * LDA : BM * LDA;
    constexpr int LDV = NR + 2, V_ELEMS = BK * LDV;
    extern __shared__ __align__(16) unsigned char mr_smem[];
    cplx* As = reinterpret_cast<cplx*>(mr_smem);
    cplx* Vs = As + STAGES * A_ELEMS;
    const long long b = blockIdx.y;
    const int m_base = blockIdx.x * BM;
    const int k_lo = blockIdx.z * a.kper, k_hi = min(a.K, k_lo + a.kper);
    const cplx* __restrict__ A = a.A + b * a.sA;
    const cplx* __restrict__ V = a.V + b * a.sV;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int KT = (k_hi - k_lo + BK - 1) / BK;

    auto load_tile = [&](int stage, int kt) {
        const int k0 = k_lo + kt * BK;
        cplx* as = As + stage * A_ELEMS;
        cplx* vs = Vs + stage * V_ELEMS;
        if (!TRANS) {
            // A tile: 128 rows x 16 k, k contiguous in memory
#pragma unroll
            for (int i = tid; i < BM * BK; i += NTHR) {
                const int r = i >> 4, c = i & 15;
                const int gm = m_base + r, gk = k0 + c;
                const bool ok = gm < a.M && gk < k_hi;
                cp_async16(as + r * LDA + c, ok ? A + (size_t)gm * a.lda + gk : A, ok);
            }
        } else {
            // A^T tile: 16 rows of the stored matrix (the K index) x 128 columns (the M index), M contiguous
#pragma unroll
            for (int i = tid; i < BM * BK; i += NTHR) {
                const int r = i / BM, c = i % BM;
                const int gk = k0 + r, gm = m_base + c;
                const bool ok = gm < a.M && gk < k_hi;
                cp_async16(as + r * LDA + c, ok ? A + (size_t)gk * a.lda + gm : A, ok);
            }
        }
        for (int i = tid; i < BK * NR; i += NTHR) {
            const int r = i / NR, c = i % NR;
            const bool ok = k0 + r < k_hi;
            cp_async16(vs + r * LDV + c, ok ? V + (size_t)(k0 + r) * NR + c : V, ok);
        }
    };

    double t1[2][NT][2], t2[2][NT][2], t3[2][NT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
            t1[i][j][0] = t1[i][j][1] = t2[i][j][0] = t2[i][j][1] = t3[i][j][0] = t3[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < KT) load_tile((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const cplx* as = As + (kt % STAGES) * A_ELEMS;
        const cplx* vs = Vs + (kt % STAGES) * V_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            cplx av[2], bv[NT];
            double as_[2], bs_[NT];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int r = warp * 16 + mt * 8 + gq;
                av[mt] = TRANS ? as[(kk + tq) * LDA + r] : as[r * LDA + kk + tq];
                as_[mt] = av[mt].x + av[mt].y;
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                bv[nt] = vs[(kk + tq) * LDV + nt * 8 + gq];
                bs_[nt] = bv[nt].x + bv[nt].y;
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dmma884(t1[mt][nt][0], t1[mt][nt][1], av[mt].x, bv[nt].x);
                    dmma884(t2[mt][nt][0], t2[mt][nt][1], av[mt].y, bv[nt].y);
                    dmma884(t3[mt][nt][0], t3[mt][nt][1], as_[mt], bs_[nt]);
                }
        }
    }
    cp_async_wait<0>();
    // epilogue: each thread owns, per (mt, nt), two adjacent right-hand sides of one row
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row = m_base + warp * 16 + mt * 8 + gq;
        if (row >= a.M) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 8 + 2 * tq;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                cplx v = make_double2(t1[mt][nt][e] - t2[mt][nt][e], t3[mt][nt][e] - t1[mt][nt][e] - t2[mt][nt][e]);
                const size_t o = (size_t)row * NR + col + e;
                if (a.kslices > 1) {
                    a.part[((size_t)blockIdx.z * a.batch + b) * a.M * NR + o] = v;
                } else {
                    if (a.D) {
                        const cplx d = a.D[b * a.sD + o];
                        v = make_double2(d.x - v.x, d.y - v.y);
                    }
                    a.C[b * a.sC + o] = v;
                }
            }
        }
    }
}

// C = D - sum over the K slices (fixed order)
template <int NR>
__global__ void mrhs_reduce_kernel(MrhsArgs a) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)a.M * NR;
    if (e >= a.batch * per) return;
    const long long b = e / per, o = e % per;
    cplx t = make_double2(0.0, 0.0);
    for (int q = 0; q < a.kslices; ++q) t = cadd(t, a.part[((size_t)q * a.batch + b) * per + o]);
    if (a.D) {
        const cplx d = a.D[b * a.sD + o];
        t = make_double2(d.x - t.x, d.y - t.y);
    }
    a.C[b * a.sC + o] = t;
}

template <int NR, bool TRANS, int BM, int STAGES>
constexpr size_t mrhs_smem_bytes() {
    return sizeof(cplx) * STAGES * ((TRANS ? 16 * (BM + 2) : BM * (16 + 4)) + 16 * (NR + 2));
}

template <int NR, bool TRANS, int BM, int STAGES>
static int mrhs_launch_cfg(const MrhsArgs& a, cudaStream_t st) {
    constexpr size_t sm = mrhs_smem_bytes<NR, TRANS, BM, STAGES>();
    static bool attr = false;
    if (!attr) {
        FDFD_CHECK(cudaFuncSetAttribute(mrhs_dmma_kernel<NR, TRANS, BM, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        attr = true;
    }
    dim3 grid((a.M + BM - 1) / BM, (unsigned)a.batch, a.kslices);
    mrhs_dmma_kernel<NR, TRANS, BM, STAGES><<<grid, BM * 2, sm, st>>>(a);
    ++g_fdfd_launches;
    return 0;
}

// scratch: device buffer for the K-slice partial sums, grown on demand (owned by the caller)
template <int NR, bool TRANS>
static int mrhs_launch(MrhsArgs a, cplx** scratch, size_t* scratch_cap, cudaStream_t st) {
    if (a.M <= 0 || a.batch <= 0) return 0;
    // short reductions (the many mid-size fronts: K of a few k-steps) take 64-row tiles and a 2-stage ring, four CTAs
    // per SM, instead of one 128-row CTA whose pipeline never fills
    const bool small = a.K <= 128;
    const int mtiles = (a.M + (small ? 63 : 127)) / (small ? 64 : 128);
    // enough CTAs for the machine: split K when the level has few fronts and a long reduction
    long long ctas = (long long)mtiles * a.batch;
    int ks = 1;
    if (ctas < 148 && a.K >= 512) ks = (int)std::min<long long>((296 + ctas - 1) / ctas, a.K / 128);
    if (ks < 1) ks = 1;
    a.kper = ((a.K + ks - 1) / ks + 15) / 16 * 16;
    a.kslices = (a.K + a.kper - 1) / a.kper;
    if (a.kslices > 1) {
        const size_t need = (size_t)a.kslices * a.batch * a.M * NR;
        if (need > *scratch_cap) {
            if (*scratch) cudaFree(*scratch);
            *scratch = nullptr;
            *scratch_cap = 0;
            FDFD_CHECK(cudaMalloc(scratch, sizeof(cplx) * need));
            *scratch_cap = need;
        }
        a.part = *scratch;
    }
    if (a.batch > 65535) FDFD_FAIL("mrhs: batch too large for one launch");
    if (small ? mrhs_launch_cfg<NR, TRANS, 64, 2>(a, st) : mrhs_launch_cfg<NR, TRANS, 128, 3>(a, st)) return -1;
    if (a.kslices > 1) {
        mrhs_reduce_kernel<NR><<<ceil_div(a.batch * a.M * NR, 256), 256, 0, st>>>(a);
        ++g_fdfd_launches;
    }
    FDFD_CHECK(cudaGetLastError());
    return 0;
}
