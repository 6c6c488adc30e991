// Batched complex128 GEMM on the FP64 tensor pipe (DMMA, mma.sync m8n8k4 f64) for sm_100a.
//
//   C[b] (M x N)  =  C[b] - A[b] (M x K) * op(B[b])          (mode 1, the Schur / sweep update)
//   C[b]          =  A[b] * op(B[b])                          (mode 0)
// op(B) = B (K x N, transb 0) or B^T with B stored N x K (transb 1: both operands have k contiguous,
// the natural form of the symmetric Schur update  S -= G F_RE^T).  lower != 0 computes only the
// 64x64 tiles on or below the diagonal (M == N; the strict upper triangle of C is never touched
// outside diagonal tiles).
//
// All matrices are row-major interleaved complex (re, im) with leading dimensions and 64-bit
// batch strides.  A complex product is THREE real DMMA products (Karatsuba / "3M"):
//        T1 = Ar Br,  T2 = Ai Bi,  T3 = (Ar + Ai)(Br + Bi);   Re C = T1 - T2,  Im C = T3 - T1 - T2,
// i.e. 6 real flops per complex multiply-add on the tensor pipe instead of the 8 of the textbook form (M3 = false
// keeps that "4M" form: the pipe is the limiter of this kernel, so 3M is a straight 4/3 on the mainloop).  3M is
// normwise backward stable like 4M; what it gives up is componentwise accuracy of a small imaginary part next to
// a large real one, which the fp64 residual refinement of the solver does not depend on.
//
// Warp tile 16 (M) x 32 (N) complex; CTA = WM x WN warps.
#pragma once
#include <vector>
#include "common.cuh"

struct GemmBatch {
    const cplx* A; long long sA; int lda;
    const cplx* B; long long sB; int ldb;
    cplx* C; long long sC; int ldc;
    int M, N, K, batch;
    int mode;   // 0: C = AB, 1: C -= AB, 2: C = -AB
    int transb; // 0: B is K x N, 1: B is N x K (C = A B^T)
    int lower;  // 1: only tiles with tile_row >= tile_col
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;     // src-size 0 -> the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Tiles are staged interleaved (re, im) by cp.async (LDGSTS) into a STAGES-deep ring; fragment loads
// are LDS.128.  Leading dimensions LDA = 4 mod 8 and LDB = 2 mod 8 complex make every 8-lane phase of
// those loads hit eight different 16-byte bank groups.
// (mma.sync m16n8k16 f64 was tried: ptxas lowers it to eight DMMA.8x8x4 on sm_100a and it ran ~6 % slower.)
template <int WM, int WN, int STAGES, bool TB, bool M3>
__global__ void __launch_bounds__(WM * WN * 32)
zgemm_dmma_kernel(GemmBatch g, int tiles_m, int tiles_n) {
    constexpr int BM = 16 * WM, BN = 32 * WN, BK = 16, NT = WM * WN * 32;
    constexpr int LDA = BK + 4, LDB = TB ? BK + 4 : BN + 2;
    constexpr int A_ELEMS = BM * LDA, B_ELEMS = TB ? BN * LDB : BK * LDB;
    extern __shared__ __align__(16) unsigned char zg_smem[];
    cplx* As = reinterpret_cast<cplx*>(zg_smem);            // [STAGES][BM][LDA]
    cplx* Bs = As + STAGES * A_ELEMS;                        // [STAGES][BK][LDB]

    long long bid = blockIdx.x;
    const int tn = (int)(bid % tiles_n);
    bid /= tiles_n;
    const int tm = (int)(bid % tiles_m);
    const long long b = bid / tiles_m;
    if (g.lower && tn * BN > tm * BM + BM - 1) return;       // tile strictly above the diagonal
    const cplx* __restrict__ A = g.A + b * g.sA;
    const cplx* __restrict__ B = g.B + b * g.sB;
    cplx* __restrict__ C = g.C + b * g.sC;
    const int m_base = tm * BM, n_base = tn * BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int gq = lane >> 2, tq = lane & 3;

    // 4M: cr, ci (t3 unused);  3M: cr = T1, ci = T2, t3 = T3
    double cr[2][4][2], ci[2][4][2], t3[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t3[i][j][0] = t3[i][j][1] = 0.0;

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kt * BK;
        cplx* as = As + stage * A_ELEMS;
        cplx* bs = Bs + stage * B_ELEMS;
#pragma unroll
        for (int i = tid; i < BM * BK; i += NT) {
            int r = i / BK, c = i % BK;
            int gm = m_base + r, gk = k0 + c;
            bool ok = gm < g.M && gk < g.K;
            cp_async16(as + r * LDA + c, ok ? A + (size_t)gm * g.lda + gk : A, ok);
        }
        if (TB) {
#pragma unroll
            for (int i = tid; i < BN * BK; i += NT) {
                int r = i / BK, c = i % BK;
                int gn = n_base + r, gk = k0 + c;
                bool ok = gn < g.N && gk < g.K;
                cp_async16(bs + r * LDB + c, ok ? B + (size_t)gn * g.ldb + gk : B, ok);
            }
        } else {
#pragma unroll
            for (int i = tid; i < BK * BN; i += NT) {
                int r = i / BN, c = i % BN;
                int gk = k0 + r, gn = n_base + c;
                bool ok = gk < g.K && gn < g.N;
                cp_async16(bs + r * LDB + c, ok ? B + (size_t)gk * g.ldb + gn : B, ok);
            }
        }
    };

    const int KT = (g.K + BK - 1) / BK;
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
        const cplx* bs = Bs + (kt % STAGES) * B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            cplx a[2], bq[4];
            double nai[2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                a[mt] = as[(wm * 16 + mt * 8 + gq) * LDA + kk + tq];
                nai[mt] = -a[mt].y;
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                bq[nt] = TB ? bs[(wn * 32 + nt * 8 + gq) * LDB + kk + tq] : bs[(kk + tq) * LDB + wn * 32 + nt * 8 + gq];
            if (M3) {
                double as_[2], bs_[4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) as_[mt] = a[mt].x + a[mt].y;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) bs_[nt] = bq[nt].x + bq[nt].y;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], a[mt].x, bq[nt].x);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a[mt].y, bq[nt].y);
                        dmma884(t3[mt][nt][0], t3[mt][nt][1], as_[mt], bs_[nt]);
                    }
            } else {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], a[mt].x, bq[nt].x);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a[mt].x, bq[nt].y);
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], nai[mt], bq[nt].y);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a[mt].y, bq[nt].x);
                    }
            }
        }
    }
    cp_async_wait<0>();
    if (M3) {               // (T1, T2, T3) -> (Re, Im)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double p1 = cr[mt][nt][e], p2 = ci[mt][nt][e];
                    cr[mt][nt][e] = p1 - p2;
                    ci[mt][nt][e] = t3[mt][nt][e] - p1 - p2;
                }
    }

    // epilogue: each thread owns, per (mt, nt), two adjacent complex entries of one row
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        int row = m_base + wm * 16 + mt * 8 + gq;
        if (row >= g.M) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int col = n_base + wn * 32 + nt * 8 + 2 * tq;
            cplx* p = C + (size_t)row * g.ldc + col;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < g.N) {
                    cplx v = make_double2(cr[mt][nt][e], ci[mt][nt][e]);
                    if (g.mode == 1) {
                        cplx o = p[e];
                        v = make_double2(o.x - v.x, o.y - v.y);
                    } else if (g.mode == 2) {
                        v = make_double2(-v.x, -v.y);
                    }
                    p[e] = v;
                }
            }
        }
    }
}

// Persistent variant for the large updates: one CTA per SM walks its share of the 64x64 output
// tiles; the cp.async ring runs ACROSS tiles (no pipeline refill per tile) and the C tile of a
// mode-1 update is prefetched into registers under the last k-tile's DMMAs, so neither the
// operand nor the accumulator-read latency is exposed.  With `lower` the tile list of a batch is
// the triangle  t = tm (tm + 1) / 2 + tn,  tn <= tm  (row-major, so concurrently running CTAs
// share the A panel of one or two tile rows).
struct TileCoord { unsigned b, tm, tn; };
__device__ __forceinline__ TileCoord decode_tile(unsigned tile, unsigned tiles_per_batch, unsigned tiles_n, int lower) {
    TileCoord t;
    t.b = tile / tiles_per_batch;
    unsigned r = tile - t.b * tiles_per_batch;
    if (lower) {
        unsigned tm = (unsigned)((sqrtf(8.0f * (float)r + 1.0f) - 1.0f) * 0.5f);
        while ((tm + 1) * (tm + 2) / 2 <= r) ++tm;
        while (tm * (tm + 1) / 2 > r) --tm;
        t.tm = tm;
        t.tn = r - tm * (tm + 1) / 2;
    } else {
        t.tm = r / tiles_n;
        t.tn = r - t.tm * tiles_n;
    }
    return t;
}

template <int STAGES, bool TB, bool M3, int BK>
__global__ void __launch_bounds__(256, 1)
zgemm_dmma_persistent_kernel(GemmBatch g, int tiles_m, int tiles_n, unsigned tiles_per_batch, long long total_tiles,
                             unsigned* __restrict__ counter) {
    constexpr int WN = 2, BM = 64, BN = 64;
    constexpr int RPP = 256 / BK, NPASS = 64 / RPP;      // rows one pass of the 256 threads covers / passes per tile
    constexpr int LDA = BK + 4, LDB = TB ? BK + 4 : BN + 2;
    constexpr int A_ELEMS = BM * LDA, B_ELEMS = TB ? BN * LDB : BK * LDB;
    extern __shared__ __align__(16) unsigned char zg_smem[];
    cplx* As = reinterpret_cast<cplx*>(zg_smem);
    cplx* Bs = As + STAGES * A_ELEMS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int gq = lane >> 2, tq = lane & 3;
    const int KT = (g.K + BK - 1) / BK;
    const unsigned ntiles = (unsigned)total_tiles;
    // Tile hand-out.  counter == nullptr: static, CTA b walks tiles b, b + gridDim.x, ...  Otherwise DYNAMIC: tiles come
    // from a global atomic counter, so CTAs that start late (an SM still busy with look-ahead work of another stream)
    // or that belong to a second, helper launch on another stream sharing the same counter simply take fewer tiles.
    // Thread 0 draws tile q + 2 of this CTA when the load cursor moves on to tile q; the draw reaches the other
    // threads through shared memory behind the k-loop's barrier (every move of the load cursor is at least one
    // barrier after the previous one), eight slots so that the compute cursor (<= 3 tiles behind) still finds its own.
    __shared__ unsigned s_tiles[8];
    const bool dyn = counter != nullptr;
    if (!dyn && blockIdx.x >= ntiles) return;
    if (dyn) {
        if (tid == 0) {
            s_tiles[0] = atomicAdd(counter, 1u);
            s_tiles[1] = atomicAdd(counter, 1u);
            s_tiles[2] = atomicAdd(counter, 1u);
        }
        __syncthreads();
        if (s_tiles[0] >= ntiles) return;
    }
    unsigned q_ld = 0;                                    // sequence number of the tile the load cursor is on

    // load cursor (runs STAGES-1 k-tiles ahead of the compute cursor).  Each thread copies four 16-byte
    // chunks of the A tile and four of the B tile per k-tile; their global pointers only advance by a
    // constant per k-tile, so the hot loop carries two pointers and two row masks and no index math.
    const int a_r = tid / BK, a_c = tid % BK;                       // A (and B^T): rows a_r + RPP i, k column a_c
    const int b_r = TB ? a_r : tid >> 6, b_c = TB ? a_c : tid & 63; // B (K x N): k rows b_r + 4 i, n column b_c
    constexpr int BPASS = TB ? NPASS : BK / 4;
    const long long a_step = (long long)RPP * g.lda, b_step = TB ? (long long)RPP * g.ldb : 4LL * g.ldb;
    const long long b_adv = TB ? (long long)BK : (long long)BK * g.ldb;
    unsigned ld_tile = dyn ? s_tiles[0] : blockIdx.x;
    int ld_kt = 0, ld_stage = 0;
    unsigned amask = 0, bmask = 0;
    const cplx *pa = g.A, *pb = g.B;
    auto decode_load = [&]() {
        TileCoord t = decode_tile(ld_tile, tiles_per_batch, (unsigned)tiles_n, g.lower);
        const int ld_m = (int)t.tm * BM, ld_n = (int)t.tn * BN;
        pa = g.A + (long long)t.b * g.sA + (long long)(ld_m + a_r) * g.lda + a_c;
        amask = 0;
#pragma unroll
        for (int i = 0; i < NPASS; ++i) amask |= (ld_m + a_r + RPP * i < g.M ? 1u : 0u) << i;
        if (TB) {
            pb = g.B + (long long)t.b * g.sB + (long long)(ld_n + b_r) * g.ldb + b_c;
            bmask = 0;
#pragma unroll
            for (int i = 0; i < NPASS; ++i) bmask |= (ld_n + b_r + RPP * i < g.N ? 1u : 0u) << i;
        } else {
            pb = g.B + (long long)t.b * g.sB + (long long)b_r * g.ldb + ld_n + b_c;
            bmask = ld_n + b_c < g.N ? 0xFFFFu : 0u;
        }
    };
    decode_load();
    auto issue_load = [&]() {
        const int k0 = ld_kt * BK;
        cplx* as = As + ld_stage * A_ELEMS + a_r * LDA + a_c;
        cplx* bs = Bs + ld_stage * B_ELEMS + b_r * LDB + b_c;
        const bool kok = k0 + a_c < g.K;
#pragma unroll
        for (int i = 0; i < NPASS; ++i) {
            bool ok = kok && ((amask >> i) & 1u);
            cp_async16(as + i * RPP * LDA, ok ? pa + i * a_step : g.A, ok);
        }
#pragma unroll
        for (int i = 0; i < BPASS; ++i) {
            bool ok = TB ? (kok && ((bmask >> i) & 1u)) : (bmask && k0 + b_r + 4 * i < g.K);
            cp_async16(bs + i * (TB ? RPP : 4) * LDB, ok ? pb + i * b_step : g.B, ok);
        }
        pa += BK;
        pb += b_adv;
        ld_stage = ld_stage + 1 == STAGES ? 0 : ld_stage + 1;
        if (++ld_kt == KT) {
            ld_kt = 0;
            if (dyn) {
                ++q_ld;
                ld_tile = s_tiles[q_ld & 7u];
                if (tid == 0) s_tiles[(q_ld + 2u) & 7u] = atomicAdd(counter, 1u);
            } else {
                ld_tile += gridDim.x;
            }
            if (ld_tile < ntiles) decode_load();
            else amask = bmask = 0;         // past the end: the copies degenerate to zero fills of a dead stage
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        issue_load();
        cp_async_commit();
    }
    double cr[2][4][2], ci[2][4][2], t3[2][4][2];       // 3M: cr = T1, ci = T2, t3 = T3
    int stage = 0;
    // operand fragments of one k-step (4 k values): a[mt], b[nt] and, for 3M, their re + im sums
    struct Frag { cplx a[2], b[4]; double as_[2], bs_[4]; };
    auto load_frag = [&](Frag& f, const cplx* as, const cplx* bs, int kk) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) f.a[mt] = as[(wm * 16 + mt * 8 + gq) * LDA + kk + tq];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            f.b[nt] = TB ? bs[(wn * 32 + nt * 8 + gq) * LDB + kk + tq] : bs[(kk + tq) * LDB + wn * 32 + nt * 8 + gq];
        // the sums (3M) / the negated imaginary part (4M) are formed one k-step AHEAD of the DMMAs that read them
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) f.as_[mt] = M3 ? f.a[mt].x + f.a[mt].y : -f.a[mt].y;
        if (M3) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) f.bs_[nt] = f.b[nt].x + f.b[nt].y;
        }
    };
    unsigned q_c = 0;
    for (unsigned tile = dyn ? s_tiles[0] : blockIdx.x; tile < ntiles; tile = dyn ? s_tiles[++q_c & 7u] : tile + gridDim.x) {
        const TileCoord tc = decode_tile(tile, tiles_per_batch, (unsigned)tiles_n, g.lower);
        cplx* __restrict__ C = g.C + (long long)tc.b * g.sC;
        const int m_base = (int)tc.tm * BM, n_base = (int)tc.tn * BN;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t3[i][j][0] = t3[i][j][1] = 0.0;
      for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const bool last = kt == KT - 1;
        if (last && g.mode == 1) {
            // the C tile of an update is pulled into L2 / L1 under the last k-tile's DMMAs (no registers held)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int row = m_base + wm * 16 + mt * 8 + gq;
                if (row < g.M) {
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = n_base + wn * 32 + nt * 8 + 2 * tq;
                        if (col < g.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(C + (size_t)row * g.ldc + col));
                    }
                }
            }
        }
        const cplx* as = As + stage * A_ELEMS;
        const cplx* bs = Bs + stage * B_ELEMS;
        stage = stage + 1 == STAGES ? 0 : stage + 1;
        Frag f[2];
        load_frag(f[0], as, bs, 0);
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            if (ks + 1 < BK / 4) load_frag(f[(ks + 1) & 1], as, bs, 4 * (ks + 1));
            const Frag& c = f[ks & 1];
            if (M3) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], c.a[mt].x, c.b[nt].x);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], c.a[mt].y, c.b[nt].y);
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma884(t3[mt][nt][0], t3[mt][nt][1], c.as_[mt], c.bs_[nt]);
            } else {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], c.a[mt].x, c.b[nt].x);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], c.a[mt].x, c.b[nt].y);
                    }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], c.as_[mt], c.b[nt].y);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], c.a[mt].y, c.b[nt].x);
                    }
            }
            if (ks == 0) {
                // refill the stage consumed in the previous iteration; issued here, under the first
                // k-step's DMMAs, so the tensor pipe is not idle while the copies are set up
                issue_load();
                cp_async_commit();
            }
        }
        if (last) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                int row = m_base + wm * 16 + mt * 8 + gq;
                if (row >= g.M) continue;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    int col = n_base + wn * 32 + nt * 8 + 2 * tq;
                    cplx* p = C + (size_t)row * g.ldc + col;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (col + e < g.N) {
                            cplx v = M3 ? make_double2(cr[mt][nt][e] - ci[mt][nt][e],
                                                       t3[mt][nt][e] - cr[mt][nt][e] - ci[mt][nt][e])
                                        : make_double2(cr[mt][nt][e], ci[mt][nt][e]);
                            if (g.mode == 1) {
                                const cplx o = p[e];
                                v = make_double2(o.x - v.x, o.y - v.y);
                            } else if (g.mode == 2) v = make_double2(-v.x, -v.y);
                            p[e] = v;
                        }
                    }
                }
            }
        }
      }
    }
    cp_async_wait<0>();
}

template <int WM, int WN, int STAGES, bool TB, int BK = 16>
constexpr size_t zgemm_smem_bytes() {
    return sizeof(cplx) * STAGES * ((16 * WM) * (BK + 4) + (TB ? (32 * WN) * (BK + 4) : BK * (32 * WN + 2)));
}

// optional live timing of every GEMM launch (CUDA events on the launching stream); see capi.cu
struct ZgemmTiming {
    bool on = false;
    std::vector<cudaEvent_t> ev;      // start/stop pairs
    std::vector<double> flops;        // algorithmic real flops of each launch (8 per complex MAC)
    std::vector<double> tflops_exec;  // real flops the tensor pipe executed (6 per complex MAC with 3M)
    std::vector<int> big;             // 1: 64x64-tile kernel, 0: 32x32-tile kernel
};
extern ZgemmTiming g_zgemm_timing;
extern thread_local int g_zgemm_max_ctas;  // CTAs of the persistent kernel (148 = one per SM; fewer leaves SMs to a side stream)
// Helper launch: while set, a persistent GEMM issued on stream S is also launched with `ctas` CTAs on `stream` (ordered
// after whatever that stream already holds, i.e. the look-ahead work the main launch left SMs free for); both launches
// draw their tiles from one atomic counter, and S waits for the helper before it goes on.
struct ZgemmHelper {
    cudaStream_t stream;
    int ctas;
    cudaEvent_t ev_fork, ev_join;
};
extern thread_local ZgemmHelper* g_zgemm_helper;
unsigned* zgemm_tile_counter(cudaStream_t stream);     // a zeroed (stream-ordered) counter out of a device ring; capi.cu
extern int g_zgemm_variant;   // bit 0: always the tiled kernel (default: persistent kernel for large problems);
                              // bit 1: textbook 4M complex products (default: 3M); bit 2: persistent kernel with
                              // k-tiles of 16 x 4 stages (default: 32 x 3 stages); bit 3: static tile hand-out
                              // (default: dynamic, from an atomic counter)

template <bool TB, bool M3>
static inline int zgemm_launch(const GemmBatch& g, cudaStream_t stream) {
    if (g.M <= 32 && g.N <= 32) {
        int tm = (g.M + 31) / 32, tn = (g.N + 31) / 32;
        long long blocks = (long long)tm * tn * g.batch;
        constexpr size_t sm = zgemm_smem_bytes<2, 1, 2, TB>();
        zgemm_dmma_kernel<2, 1, 2, TB, M3><<<(unsigned)blocks, 64, sm, stream>>>(g, tm, tn);
        ++g_fdfd_launches;
        return 0;
    }
    int tm = (g.M + 63) / 64, tn = (g.N + 63) / 64;
    long long per_batch = g.lower ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn;
    long long blocks = per_batch * g.batch;
    if (blocks > 2147483647LL || (long long)tm * tn * g.batch > 2147483647LL) {
        snprintf(g_fdfd_err, sizeof(g_fdfd_err), "zgemm grid too large");
        return -1;
    }
    if ((g_zgemm_variant & 1) == 0 && blocks >= 148) {
        // k-tiles of 32 with a 3-stage ring (221 KB): half the barriers and post-barrier bubbles of the 16 x 4-stage
        // form (bit 2 of the variant keeps that one for A/B runs), the same bytes in flight
        if ((g_zgemm_variant & 4) == 0) {
            constexpr int ST = 3, BK = 32;
            constexpr size_t sm = zgemm_smem_bytes<4, 2, ST, TB, BK>();
            static bool attr_p = false;
            if (!attr_p) {
                cudaFuncSetAttribute(zgemm_dmma_persistent_kernel<ST, TB, M3, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sm);
                attr_p = true;
            }
            unsigned* cnt = (g_zgemm_variant & 8) ? nullptr : zgemm_tile_counter(stream);
            ZgemmHelper* h = cnt ? g_zgemm_helper : nullptr;
            if (h) {
                cudaEventRecord(h->ev_fork, stream);                     // the counter is zero from here on
                cudaStreamWaitEvent(h->stream, h->ev_fork, 0);
            }
            zgemm_dmma_persistent_kernel<ST, TB, M3, BK><<<g_zgemm_max_ctas, 256, sm, stream>>>(g, tm, tn, (unsigned)per_batch, blocks, cnt);
            if (h) {
                zgemm_dmma_persistent_kernel<ST, TB, M3, BK><<<h->ctas, 256, sm, h->stream>>>(g, tm, tn, (unsigned)per_batch, blocks, cnt);
                ++g_fdfd_launches;
                cudaEventRecord(h->ev_join, h->stream);
                cudaStreamWaitEvent(stream, h->ev_join, 0);
            }
        } else {
            constexpr int ST = 4, BK = 16;
            constexpr size_t sm = zgemm_smem_bytes<4, 2, ST, TB, BK>();
            static bool attr_p = false;
            if (!attr_p) {
                cudaFuncSetAttribute(zgemm_dmma_persistent_kernel<ST, TB, M3, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sm);
                attr_p = true;
            }
            zgemm_dmma_persistent_kernel<ST, TB, M3, BK><<<g_zgemm_max_ctas, 256, sm, stream>>>(g, tm, tn, (unsigned)per_batch, blocks, nullptr);
        }
    } else {
        constexpr size_t sm = zgemm_smem_bytes<4, 2, 3, TB>();
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(zgemm_dmma_kernel<4, 2, 3, TB, M3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            attr_set = true;
        }
        zgemm_dmma_kernel<4, 2, 3, TB, M3><<<(unsigned)((long long)tm * tn * g.batch), 256, sm, stream>>>(g, tm, tn);
    }
    ++g_fdfd_launches;
    return 0;
}

static inline int zgemm_batched(const GemmBatch& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    if (g.lower && g.M != g.N) {
        snprintf(g_fdfd_err, sizeof(g_fdfd_err), "zgemm: lower needs a square C");
        return -1;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_zgemm_timing.on) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    const bool m3 = (g_zgemm_variant & 2) == 0;
    int rc = g.transb ? (m3 ? zgemm_launch<true, true>(g, stream) : zgemm_launch<true, false>(g, stream))
                      : (m3 ? zgemm_launch<false, true>(g, stream) : zgemm_launch<false, false>(g, stream));
    if (rc) return rc;
    if (g_zgemm_timing.on) {
        cudaEventRecord(e1, stream);
        g_zgemm_timing.ev.push_back(e0);
        g_zgemm_timing.ev.push_back(e1);
        // flops actually computed: a lower-masked launch does (about) half of M*N*K
        double mn = g.lower ? 0.5 * (double)g.M * ((double)g.N + 64.0) : (double)g.M * (double)g.N;
        g_zgemm_timing.flops.push_back(8.0 * mn * g.K * g.batch);
        g_zgemm_timing.tflops_exec.push_back((m3 ? 6.0 : 8.0) * mn * g.K * g.batch);
        g_zgemm_timing.big.push_back((g.M <= 32 && g.N <= 32) ? 0 : 1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_fdfd_err, sizeof(g_fdfd_err), "zgemm launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}
