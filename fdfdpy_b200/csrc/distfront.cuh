// Distributed fronts: the fronts of the levels that several ranks share (the top log2(world) merges of the sharded
// elimination tree, ndplan.DistFront) are stored, factorised and substituted by ALL ranks of their group.
//
//   layout    rows of the compact front are dealt to the group in blocks (pivot pieces, then ring pieces; serpentine
//             ownership so the triangular Schur work balances); a rank keeps its blocks back to back, full row width
//             (F[nloc][n], lower part valid).
//   assembly  extend-add of the two children's Schur blocks as a personalised all-to-all: every rank packs, for every
//             destination, the entries of ITS part of ITS child that land on the destination's rows (dense rows, zeros
//             elsewhere), the destination sums what arrives.  g - 1 send/recv rounds, no index lists on the wire.
//   step s    the owner of pivot block s inverts it (recursive block inversion, as on one GPU) and broadcasts Einv;
//             every rank forms its rows of G = F_RE Einv; the F_RE panel is all-gathered (block broadcasts in one
//             group) and every rank updates its own block rows of S -= G F_RE^T with the same DMMA GEMM kernel.
//   solve     front vectors are replicated in the group; forward: the pivot owner broadcasts its (already updated)
//             right-hand side piece, every rank updates the rows it owns; backward: G^T u_R is summed over the ranks
//             in a fixed order (all-gather of the partial sums), so every rank holds the same bits.
// All sub-group collectives are grouped NCCL send/recv on the world communicator (groups of 2, 4, 8 ranks).
// This file is included by direct.cu (it uses its static helpers).
#pragma once

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// out[lp][q] (q <= p = drow[lp]) = entry of MY part of MY child that lands on front entry (p, q), else 0.
// Child entry (a, b), a >= b, of the child's ring lives at cbase[croff[a] + b] (croff[a] < 0: row a is not mine);
// croff == nullptr: a local child, cbase[a * cld + b].
__global__ void __launch_bounds__(256)
dist_pack_kernel(cplx* __restrict__ out, const int* __restrict__ drow, const int* __restrict__ inv,
                 const cplx* __restrict__ cbase, const long long* __restrict__ croff, long long cld, int n) {
    const int lp = blockIdx.x;
    const int p = drow[lp];
    const int a = inv[p];
    cplx* row = out + (size_t)lp * n;
    for (int q = threadIdx.x; q <= p; q += blockDim.x) {
        cplx v = make_double2(0.0, 0.0);
        const int b = a >= 0 ? inv[q] : -1;
        if (b >= 0) {
            const int hi = max(a, b), lo = min(a, b);
            const long long off = croff ? croff[hi] : (long long)hi * cld;
            if (off >= 0) v = cbase[off + lo];
        }
        row[q] = v;
    }
}
__global__ void __launch_bounds__(256)
dist_add_kernel(cplx* __restrict__ F, const cplx* __restrict__ in, const int* __restrict__ drow, int n) {
    const int lp = blockIdx.x;
    const int p = drow[lp];
    cplx* row = F + (size_t)lp * n;
    const cplx* src = in + (size_t)lp * n;
    for (int q = threadIdx.x; q <= p; q += blockDim.x) row[q] = cadd(row[q], src[q]);
}

// f[p] = ring1[inv1[p]] + ring2[inv2[p]]
template <int NR>
__global__ void dist_gather_kernel(cplx* __restrict__ f, const cplx* __restrict__ ring1, const cplx* __restrict__ ring2,
                                   const int* __restrict__ inv1, const int* __restrict__ inv2, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int a1 = inv1[p], a2 = inv2[p];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        cplx v = make_double2(0.0, 0.0);
        if (a1 >= 0) v = ring1[(size_t)a1 * NR + j];
        if (a2 >= 0) v = cadd(v, ring2[(size_t)a2 * NR + j]);
        f[(size_t)p * NR + j] = v;
    }
}
// dst[i] = src[map[i]]
template <int NR>
__global__ void dist_pick_kernel(cplx* __restrict__ dst, const cplx* __restrict__ src, const int* __restrict__ map, int cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int s = map[i];
#pragma unroll
    for (int j = 0; j < NR; ++j) dst[(size_t)i * NR + j] = src[(size_t)s * NR + j];
}

// forward step: rows 0..k-1: yE[col0 + r] = Einv[r][:] fE ; rows k..k+mloc-1: f[lslot[r0 + i]] -= G[i][:] fE.
// One warp per row, lanes stride the pivot index, NR right-hand sides per lane.
template <int NR>
__global__ void __launch_bounds__(256)
dist_fwd_step_kernel(const cplx* __restrict__ Einv, const cplx* __restrict__ G, cplx* __restrict__ f, cplx* __restrict__ yE,
                     const int* __restrict__ lslot, int r0, int mloc, int k, int col0) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= k + mloc) return;
    const cplx* mrow = row < k ? Einv + (size_t)row * k : G + (size_t)(row - k) * k;
    const cplx* fE = f + (size_t)col0 * NR;
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    for (int c = lane; c < k; c += 32) {
        const cplx mv = ldg_c(mrow + c);
#pragma unroll
        for (int j = 0; j < NR; ++j) cfma(acc[j], mv, fE[(size_t)c * NR + j]);
    }
#pragma unroll
    for (int j = 0; j < NR; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[j].x += __shfl_down_sync(0xffffffffu, acc[j].x, o);
            acc[j].y += __shfl_down_sync(0xffffffffu, acc[j].y, o);
        }
    if (lane == 0) {
        if (row < k) {
#pragma unroll
            for (int j = 0; j < NR; ++j) yE[(size_t)(col0 + row) * NR + j] = acc[j];
        } else {
            const size_t slot = lslot[r0 + row - k];
#pragma unroll
            for (int j = 0; j < NR; ++j) f[slot * NR + j] = csub(f[slot * NR + j], acc[j]);
        }
    }
}

// backward: tmp[split][c] = sum over my rows i of this split of G[i][c] u[lslot[r0 + i]]   (lanes along c)
template <int NR>
__global__ void __launch_bounds__(128)
dist_bwd_partial_kernel(const cplx* __restrict__ G, const cplx* __restrict__ u, const int* __restrict__ lslot, int r0,
                        int mloc, int k, int rows_per_split, cplx* __restrict__ tmp) {
    __shared__ cplx part[4][32][NR];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int i0 = blockIdx.y * rows_per_split, i1 = min(mloc, i0 + rows_per_split);
    cplx acc[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = make_double2(0.0, 0.0);
    if (c < k) {
        for (int i = i0 + w; i < i1; i += 4) {
            const cplx mv = ldg_c(G + (size_t)i * k + c);
            const size_t slot = lslot[r0 + i];
#pragma unroll
            for (int j = 0; j < NR; ++j) cfma(acc[j], mv, u[slot * NR + j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) part[w][lane][j] = acc[j];
    __syncthreads();
    if (w == 0 && c < k) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            cplx t = part[0][lane][j];
#pragma unroll
            for (int q = 1; q < 4; ++q) t = cadd(t, part[q][lane][j]);
            tmp[((size_t)blockIdx.y * k + c) * NR + j] = t;
        }
    }
}
// out[c] = sum over splits (fixed order)
template <int NR>
__global__ void dist_bwd_reduce_kernel(const cplx* __restrict__ tmp, cplx* __restrict__ out, int k, int nsplit) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * NR) return;
    cplx t = make_double2(0.0, 0.0);
    for (int q = 0; q < nsplit; ++q) t = cadd(t, tmp[(size_t)q * k * NR + e]);
    out[e] = t;
}
// u[col0 + c] = yE[col0 + c] - sum over the group ranks (fixed order) of gat[rank][c]
template <int NR>
__global__ void dist_bwd_finish_kernel(cplx* __restrict__ u, const cplx* __restrict__ yE, const cplx* __restrict__ gat,
                                       int g, int k, int col0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * NR) return;
    cplx t = make_double2(0.0, 0.0);
    for (int q = 0; q < g; ++q) t = cadd(t, gat[(size_t)q * k * NR + e]);
    u[(size_t)col0 * NR + e] = csub(yE[(size_t)col0 * NR + e], t);
}
// local child below the first distributed front: its ring solution out of the front vector
template <int NR>
__global__ void dist_child_ring_kernel(cplx* __restrict__ uc, const cplx* __restrict__ u, const int* __restrict__ cmap,
                                       int mc, int kc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mc) return;
    const int s = cmap[i];
#pragma unroll
    for (int j = 0; j < NR; ++j) uc[(size_t)(kc + i) * NR + j] = u[(size_t)s * NR + j];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
#define DIST_NR_MAX 16

static int dist_upload(int** dst, const std::vector<int>& v) { return upload_i32(dst, v.data(), v.size()); }

int nd_add_dist_front(NdSolver* s, const NdDistFrontDesc* d) {
    if (!s->comm) FDFD_FAIL("distributed front without a communicator (fdfd_direct_set_comm first)");
    const int rank = s->comm->rank;
    if (d->gsize < 2 || rank < d->gbase || rank >= d->gbase + d->gsize) FDFD_FAIL("rank %d is not in the front's group", rank);
    if (d->level0 < 1 || d->level0 + d->nsteps > (int)s->levels.size()) FDFD_FAIL("distributed front: level range out of the plan");
    if (d->nsteps < 1 || d->nsteps > d->nblk) FDFD_FAIL("distributed front: bad block structure");
    NdDistFront* f = new NdDistFront();
    f->level0 = d->level0; f->nsteps = d->nsteps; f->gbase = d->gbase; f->gsize = d->gsize;
    f->grank = rank - d->gbase; f->cidx = f->grank >= d->gsize / 2 ? 1 : 0;
    f->n = d->n; f->nblk = d->nblk;
    f->bstart.assign(d->bstart, d->bstart + d->nblk + 1);
    f->bowner.assign(d->bowner, d->bowner + d->nblk);
    if (f->bstart[0] != 0 || f->bstart[d->nblk] != d->n) FDFD_FAIL("distributed front: blocks do not cover the front");
    f->kfull = f->bstart[d->nsteps];
    f->m = d->n - f->kfull;
    f->mc[0] = d->mc1; f->mc[1] = d->mc2;
    f->kmax_step = 0;
    for (int sidx = 0; sidx < d->nsteps; ++sidx) f->kmax_step = std::max(f->kmax_step, f->bstart[sidx + 1] - f->bstart[sidx]);
    // local rows
    f->lrow0.assign(d->nblk, -1);
    f->nloc_of.assign(d->gsize, 0);
    std::vector<std::vector<int>> rows(d->gsize);
    for (int j = 0; j < d->nblk; ++j) {
        const int o = f->bowner[j];
        if (o < 0 || o >= d->gsize) FDFD_FAIL("distributed front: bad block owner");
        if (o == f->grank) f->lrow0[j] = f->nloc_of[o];
        for (int p = f->bstart[j]; p < f->bstart[j + 1]; ++p) rows[o].push_back(p);
        f->nloc_of[o] += f->bstart[j + 1] - f->bstart[j];
    }
    f->nloc = f->nloc_of[f->grank];
    f->d_rows_of.assign(d->gsize, nullptr);
    for (int o = 0; o < d->gsize; ++o)
        if (dist_upload(&f->d_rows_of[o], rows[o])) return -1;
    f->d_lslot = f->d_rows_of[f->grank];
    f->r0.assign(d->nsteps, 0);
    for (int sidx = 0; sidx < d->nsteps; ++sidx) {
        int r = 0;
        for (int j = 0; j <= sidx; ++j)
            if (f->bowner[j] == f->grank) r += f->bstart[j + 1] - f->bstart[j];
        f->r0[sidx] = r;
    }
    // maps
    std::vector<int> inv1(d->inv1, d->inv1 + d->n), inv2(d->inv2, d->inv2 + d->n);
    if (dist_upload(&f->d_inv[0], inv1) || dist_upload(&f->d_inv[1], inv2)) return -1;
    const std::vector<int>& mine = f->cidx ? inv2 : inv1;
    std::vector<int> cmap(f->mc[f->cidx], -1);
    for (int p = 0; p < d->n; ++p)
        if (mine[p] >= 0) {
            if (mine[p] >= (int)cmap.size()) FDFD_FAIL("distributed front: child map out of range");
            cmap[mine[p]] = p;
        }
    for (int v : cmap)
        if (v < 0) FDFD_FAIL("distributed front: a child ring node is missing from the front");
    if (dist_upload(&f->d_cmap_mine, cmap)) return -1;
    // where my rows of the final ring sit inside F (for the parent's assembly)
    f->d_ring_off = nullptr;
    if (f->m > 0) {
        std::vector<long long> off(f->m, -1);
        for (int j = d->nsteps; j < d->nblk; ++j)
            if (f->bowner[j] == f->grank)
                for (int p = f->bstart[j]; p < f->bstart[j + 1]; ++p)
                    off[p - f->kfull] = (long long)(f->lrow0[j] + p - f->bstart[j]) * d->n + f->kfull;
        FDFD_CHECK(cudaMalloc(&f->d_ring_off, sizeof(long long) * f->m));
        FDFD_CHECK(cudaMemcpy(f->d_ring_off, off.data(), sizeof(long long) * f->m, cudaMemcpyHostToDevice));
    }
    f->Einv.assign(d->nsteps, nullptr);
    f->G.assign(d->nsteps, nullptr);
    f->F = nullptr;
    f->vec = f->yE = f->oring = f->gat = nullptr;
    // consistency with the fronts registered before (front j + 1 is the parent of front j)
    if (!s->dist.empty()) {
        const NdDistFront* c = s->dist.back();
        if (c->level0 + c->nsteps != f->level0 || c->m != f->mc[f->cidx] || 2 * c->gsize != f->gsize)
            FDFD_FAIL("distributed fronts must be added in elimination order, each the parent of the one before");
    } else {
        const NdLevel& L = s->levels[f->level0 - 1];
        if (L.nb != 1) FDFD_FAIL("the level below the first distributed front must hold exactly one local front");
    }
    s->dist.push_back(f);
    s->factored = false;
    return 0;
}

static void dist_destroy(NdSolver* s) {
    for (NdDistFront* f : s->dist) {
        for (cplx* p : f->Einv) if (p) cudaFree(p);
        for (cplx* p : f->G) if (p) cudaFree(p);
        for (int* p : f->d_rows_of) if (p) cudaFree(p);
        cudaFree(f->d_inv[0]); cudaFree(f->d_inv[1]); cudaFree(f->d_cmap_mine);
        if (f->d_ring_off) cudaFree(f->d_ring_off);
        if (f->F) cudaFree(f->F);
        if (f->vec) cudaFree(f->vec);
        if (f->yE) cudaFree(f->yE);
        if (f->oring) cudaFree(f->oring);
        if (f->gat) cudaFree(f->gat);
        delete f;
    }
    s->dist.clear();
    if (s->dist_send) cudaFree(s->dist_send);
    if (s->dist_recv) cudaFree(s->dist_recv);
    if (s->dist_panel) cudaFree(s->dist_panel);
    s->dist_send = s->dist_recv = s->dist_panel = nullptr;
    s->dist_send_cap = s->dist_recv_cap = s->dist_panel_cap = 0;
}

static int dist_grow(cplx** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    FDFD_CHECK(cudaMalloc(p, sizeof(cplx) * need));
    *cap = need;
    return 0;
}

static NdDistFront* dist_at_level(NdSolver* s, int li, int* index) {
    for (size_t j = 0; j < s->dist.size(); ++j)
        if (s->dist[j]->level0 == li) {
            if (index) *index = (int)j;
            return s->dist[j];
        }
    return nullptr;
}

// broadcast inside the front's group (grouped point-to-point: sub-communicators are not needed for 2..8 ranks)
static int dist_bcast(NdSolver* s, const NdDistFront* f, void* buf, size_t doubles, int root, cudaStream_t st) {
    if (doubles == 0) return 0;
    if (comm_group_begin(s->comm)) return -1;
    int rc = 0;
    if (f->grank == root) {
        for (int p = 0; p < f->gsize && !rc; ++p)
            if (p != root) rc = comm_send(s->comm, buf, doubles, f->gbase + p, st);
    } else {
        rc = comm_recv(s->comm, buf, doubles, f->gbase + root, st);
    }
    if (comm_group_end(s->comm)) return -1;
    return rc;
}
// every block j0 <= j < j1 of a row-blocked array (row = `width` complex entries, block rows bstart[j]..bstart[j+1],
// base = pointer to row bstart[j0]) goes from its owner to the whole group, all transfers in one group
static int dist_allgather_blocks(NdSolver* s, const NdDistFront* f, cplx* base, int j0, int j1, size_t width, cudaStream_t st) {
    if (j0 >= j1 || width == 0) return 0;
    if (comm_group_begin(s->comm)) return -1;
    int rc = 0;
    for (int j = j0; j < j1 && !rc; ++j) {
        cplx* p = base + (size_t)(f->bstart[j] - f->bstart[j0]) * width;
        const size_t cnt = 2 * (size_t)(f->bstart[j + 1] - f->bstart[j]) * width;
        if (f->bowner[j] == f->grank) {
            for (int q = 0; q < f->gsize && !rc; ++q)
                if (q != f->grank) rc = comm_send(s->comm, p, cnt, f->gbase + q, st);
        } else {
            rc = comm_recv(s->comm, p, cnt, f->gbase + f->bowner[j], st);
        }
    }
    if (comm_group_end(s->comm)) return -1;
    return rc;
}

// child: the distributed front below (its F holds the Schur rows), or nullptr for a local child whose Schur block
// starts at cbase with leading dimension cld
static int dist_factor(NdSolver* s, NdDistFront* f, const NdDistFront* child, const cplx* cbase, long long cld,
                       cudaStream_t st) {
    const int g = f->gsize, me = f->grank, n = f->n;
    int max_rows = 0;
    for (int v : f->nloc_of) max_rows = std::max(max_rows, v);
    if (!f->F && f->nloc > 0) {
        FDFD_CHECK(cudaMalloc(&f->F, sizeof(cplx) * (size_t)f->nloc * n));
        FDFD_CHECK(cudaMemsetAsync(f->F, 0, sizeof(cplx) * (size_t)f->nloc * n, st));   // the never-read upper part stays finite
    }
    if (dist_grow(&s->dist_send, &s->dist_send_cap, (size_t)max_rows * n)) return -1;
    if (dist_grow(&s->dist_recv, &s->dist_recv_cap, (size_t)std::max(f->nloc, 1) * n)) return -1;
    if (dist_grow(&s->dist_panel, &s->dist_panel_cap, (size_t)std::max(n - f->bstart[1], 1) * f->kmax_step)) return -1;
    const long long* croff = child ? child->d_ring_off : nullptr;
    if (child) cbase = child->F;
    // ---- assembly: personalised all-to-all of packed contributions
    {
        PhaseScope ph(PH_ASSEMBLE, st);
        for (int t = 0; t < g; ++t) {
            const int d = (me + t) % g, src = (me - t + g) % g;
            const int rows_d = f->nloc_of[d];
            cplx* out = t == 0 ? f->F : s->dist_send;
            if (rows_d > 0) {
                dist_pack_kernel<<<rows_d, 256, 0, st>>>(out, f->d_rows_of[d], f->d_inv[f->cidx], cbase, croff, cld, n);
                ++g_fdfd_launches;
                FDFD_CHECK(cudaGetLastError());
            }
            if (t == 0) continue;
            if (comm_group_begin(s->comm)) return -1;
            int rc = 0;
            if (rows_d > 0) rc = comm_send(s->comm, s->dist_send, 2 * (size_t)rows_d * n, f->gbase + d, st);
            if (!rc && f->nloc > 0) rc = comm_recv(s->comm, s->dist_recv, 2 * (size_t)f->nloc * n, f->gbase + src, st);
            if (comm_group_end(s->comm) || rc) return -1;
            if (f->nloc > 0) {
                dist_add_kernel<<<f->nloc, 256, 0, st>>>(f->F, s->dist_recv, f->d_lslot, n);
                ++g_fdfd_launches;
                FDFD_CHECK(cudaGetLastError());
            }
        }
    }
    // ---- elimination steps
    bool la_pending = false;        // this rank already inverted (is inverting) the coming pivot block on the side stream
    for (int sidx = 0; sidx < f->nsteps; ++sidx) {
        g_phase_timing.level = f->level0 + sidx;
        const int col0 = f->bstart[sidx], b1 = f->bstart[sidx + 1], k = b1 - col0, owner = f->bowner[sidx];
        const int mbelow = n - b1, r0 = f->r0[sidx], mloc = f->nloc - r0;
        if (!f->Einv[sidx]) FDFD_CHECK(cudaMalloc(&f->Einv[sidx], sizeof(cplx) * (size_t)k * k));
        if (me == owner && la_pending) {
            FDFD_CHECK(cudaStreamWaitEvent(st, s->la_done, 0));
            la_pending = false;
        } else if (me == owner) {
            const cplx* piv = f->F + (size_t)f->lrow0[sidx] * n + col0;
            if (k <= 64) {
                PhaseScope ph(PH_PIVOT, st);
                launch_tile_inverse(piv, 0, n, k, f->Einv[sidx], 0, k, s->d_info, 1, 1, st);
            } else {
                {
                    PhaseScope ph(PH_EXTRACT, st);
                    int chunks = chunks_for((long long)k * k, 1);
                    sym_expand_kernel<<<(unsigned)chunks, 256, 0, st>>>(piv, f->Einv[sidx], k, n, 0, chunks);
                    ++g_fdfd_launches;
                }
                if (sym_invert_batch(s, f->Einv[sidx], (long long)k * k, k, k, 1, s->fws_W, st)) return -1;
            }
            FDFD_CHECK(cudaGetLastError());
        }
        {
            PhaseScope ph(PH_COPY, st);
            if (dist_bcast(s, f, f->Einv[sidx], 2 * (size_t)k * k, owner, st)) return -1;
        }
        s->factor_bytes += sizeof(cplx) * (size_t)k * k;
        if (mbelow == 0) continue;
        GemmBatch gb;
        gb.batch = 1;
        if (mloc > 0) {
            if (!f->G[sidx]) FDFD_CHECK(cudaMalloc(&f->G[sidx], sizeof(cplx) * (size_t)mloc * k));
            // G = F_RE Einv on my rows below the pivot block
            gb.transb = 1; gb.lower = 0; gb.mode = 0;
            gb.A = f->F + (size_t)r0 * n + col0; gb.sA = 0; gb.lda = n;
            gb.B = f->Einv[sidx]; gb.sB = 0; gb.ldb = k;
            gb.C = f->G[sidx]; gb.sC = 0; gb.ldc = k;
            gb.M = mloc; gb.N = k; gb.K = k;
            PhaseScope ph(PH_GGEMM, st);
            if (zgemm_batched(gb, st)) return -1;
            s->factor_flops += 8.0 * (double)mloc * k * k;
            s->factor_bytes += sizeof(cplx) * (size_t)mloc * k;
        }
        {
            // the F_RE panel of the whole front, in global row order
            PhaseScope ph(PH_COPY, st);
            for (int j = sidx + 1; j < f->nblk; ++j)
                if (f->bowner[j] == me)
                    FDFD_CHECK(cudaMemcpy2DAsync(s->dist_panel + (size_t)(f->bstart[j] - b1) * k, sizeof(cplx) * k,
                                                 f->F + (size_t)f->lrow0[j] * n + col0, sizeof(cplx) * n, sizeof(cplx) * k,
                                                 f->bstart[j + 1] - f->bstart[j], cudaMemcpyDeviceToDevice, st));
            if (dist_allgather_blocks(s, f, s->dist_panel, sidx + 1, f->nblk, k, st)) return -1;
        }
        {
            // S -= G F_RE^T on my block rows: columns from the first remaining slot up to the end of the block itself
            PhaseScope ph(PH_SCHUR, st);
            auto update_block = [&](int j) -> int {
                const int rows_j = f->bstart[j + 1] - f->bstart[j], ncols = f->bstart[j + 1] - b1;
                gb.transb = 1; gb.lower = 0; gb.mode = 1;
                gb.A = f->G[sidx] + (size_t)(f->lrow0[j] - r0) * k; gb.sA = 0; gb.lda = k;
                gb.B = s->dist_panel; gb.sB = 0; gb.ldb = k;
                gb.C = f->F + (size_t)f->lrow0[j] * n + b1; gb.sC = 0; gb.ldc = n;
                gb.M = rows_j; gb.N = ncols; gb.K = k;
                s->factor_flops += 8.0 * (double)rows_j * ncols * k;
                return zgemm_batched(gb, st);
            };
            // LOOK-AHEAD: if the next pivot block is mine, its rows are updated first and it is inverted on the side
            // stream while the other block rows (mine and everybody else's) are still being updated
            const int nxt = sidx + 1;
            const int k1 = nxt < f->nsteps ? f->bstart[nxt + 1] - f->bstart[nxt] : 0;
            const bool la = g_lookahead_enabled && nxt < f->nsteps && f->bowner[nxt] == me && k1 > 64;
            if (la) {
                if (update_block(nxt)) return -1;
                if (!f->Einv[nxt]) FDFD_CHECK(cudaMalloc(&f->Einv[nxt], sizeof(cplx) * (size_t)k1 * k1));
                FDFD_CHECK(cudaEventRecord(s->la_ready, st));
                FDFD_CHECK(cudaStreamWaitEvent(s->la_stream, s->la_ready, 0));
                const cplx* piv = f->F + (size_t)f->lrow0[nxt] * n + f->bstart[nxt];
                int chunks = chunks_for((long long)k1 * k1, 1);
                sym_expand_kernel<<<(unsigned)chunks, 256, 0, s->la_stream>>>(piv, f->Einv[nxt], k1, n, 0, chunks);
                ++g_fdfd_launches;
                const bool timing = g_phase_timing.on;
                g_phase_timing.on = false;
                // wide groups: a rank's share of the update is shorter than the inversion, which is then the critical
                // path and has the machine almost to itself -> the short-chain block Gauss-Jordan form
                g_gj_on_lookahead = g_dist_gj_group > 0 && f->gsize >= g_dist_gj_group;
                int rc = sym_invert_batch(s, f->Einv[nxt], (long long)k1 * k1, k1, k1, 1, s->fws_W, s->la_stream);
                g_gj_on_lookahead = 0;
                g_phase_timing.on = timing;
                if (rc) return -1;
                FDFD_CHECK(cudaEventRecord(s->la_done, s->la_stream));
                la_pending = true;
                g_zgemm_max_ctas = 148 - 8;
            }
            int rc = 0;
            // (no helper launch here, unlike the single-GPU chain levels: a rank updates SEVERAL blocks per step, and the
            // join of the first update's helper would hold the main stream until the whole look-ahead chain is through;
            // measured on 8 GPUs: 116 -> 129 ms)
            for (int j = sidx + 1; j < f->nblk && !rc; ++j) {
                if (f->bowner[j] != me || (la && j == nxt)) continue;
                rc = update_block(j);
            }
            g_zgemm_max_ctas = 148;
            if (rc) return -1;
        }
    }
    return 0;
}

static int dist_solve_workspace(NdDistFront* f) {
    if (f->vec) return 0;
    const size_t nr = DIST_NR_MAX;
    FDFD_CHECK(cudaMalloc(&f->vec, sizeof(cplx) * (size_t)f->n * nr));
    FDFD_CHECK(cudaMalloc(&f->yE, sizeof(cplx) * (size_t)std::max(f->kfull, 1) * nr));
    FDFD_CHECK(cudaMalloc(&f->oring, sizeof(cplx) * (size_t)std::max(f->mc[0], f->mc[1]) * nr));
    // gather buffer: [gsize][kmax_step][nr] for the summed partial products, then the row-split partials behind it
    FDFD_CHECK(cudaMalloc(&f->gat, sizeof(cplx) * (size_t)(f->gsize + 64) * f->kmax_step * nr));
    return 0;
}

// my_ring: the ring right-hand side of MY child ([mc][NR], replicated in the child's group)
template <int NR>
static int dist_forward(NdSolver* s, NdDistFront* f, const cplx* my_ring, cudaStream_t st) {
    if (dist_solve_workspace(f)) return -1;
    const int g = f->gsize, me = f->grank, n = f->n;
    const int partner = f->gbase + (me + g / 2) % g;
    if (comm_group_begin(s->comm)) return -1;
    int rc = comm_send(s->comm, my_ring, 2 * (size_t)f->mc[f->cidx] * NR, partner, st);
    if (!rc) rc = comm_recv(s->comm, f->oring, 2 * (size_t)f->mc[1 - f->cidx] * NR, partner, st);
    if (comm_group_end(s->comm) || rc) return -1;
    const cplx *ring1 = f->cidx == 0 ? my_ring : f->oring, *ring2 = f->cidx == 0 ? f->oring : my_ring;
    { dist_gather_kernel<NR><<<ceil_div(n, 128), 128, 0, st>>>(f->vec, ring1, ring2, f->d_inv[0], f->d_inv[1], n); ++g_fdfd_launches; }
    FDFD_CHECK(cudaGetLastError());
    for (int sidx = 0; sidx < f->nsteps; ++sidx) {
        const int col0 = f->bstart[sidx], k = f->bstart[sidx + 1] - col0;
        const int r0 = f->r0[sidx], mloc = f->bstart[sidx + 1] < n ? f->nloc - r0 : 0;
        if (dist_bcast(s, f, f->vec + (size_t)col0 * NR, 2 * (size_t)k * NR, f->bowner[sidx], st)) return -1;
        { dist_fwd_step_kernel<NR><<<ceil_div((long long)(k + mloc) * 32, 256), 256, 0, st>>>(
              f->Einv[sidx], f->G[sidx], f->vec, f->yE, f->d_lslot, r0, mloc, k, col0); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    // the ring handed to the parent: every rank gets every block
    return dist_allgather_blocks(s, f, f->vec + (size_t)f->kfull * NR, f->nsteps, f->nblk, NR, st);
}

template <int NR>
static int dist_backward(NdSolver* s, NdDistFront* f, const NdDistFront* parent, cudaStream_t st) {
    const int g = f->gsize, me = f->grank, n = f->n;
    if (parent && f->m > 0) {
        // (the parent's vector is replicated over its group, which contains mine: no communication)
        dist_pick_kernel<NR><<<ceil_div(f->m, 128), 128, 0, st>>>(f->vec + (size_t)f->kfull * NR, parent->vec, parent->d_cmap_mine, f->m);
        ++g_fdfd_launches;
        FDFD_CHECK(cudaGetLastError());
    }
    for (int sidx = f->nsteps - 1; sidx >= 0; --sidx) {
        const int col0 = f->bstart[sidx], k = f->bstart[sidx + 1] - col0;
        const int r0 = f->r0[sidx], mloc = f->bstart[sidx + 1] < n ? f->nloc - r0 : 0;
        cplx* mine = f->gat + (size_t)me * k * NR;
        if (mloc > 0) {
            int nsplit = std::min(64, std::max(1, mloc / 64));
            const int rows_per = ceil_div(mloc, nsplit);
            nsplit = ceil_div(mloc, rows_per);
            cplx* tmp = f->gat + (size_t)g * f->kmax_step * NR;
            dim3 grid(ceil_div(k, 32), nsplit);
            { dist_bwd_partial_kernel<NR><<<grid, 128, 0, st>>>(f->G[sidx], f->vec, f->d_lslot, r0, mloc, k, rows_per, tmp); ++g_fdfd_launches; }
            { dist_bwd_reduce_kernel<NR><<<ceil_div(k * NR, 128), 128, 0, st>>>(tmp, mine, k, nsplit); ++g_fdfd_launches; }
            FDFD_CHECK(cudaGetLastError());
        } else {
            FDFD_CHECK(cudaMemsetAsync(mine, 0, sizeof(cplx) * (size_t)k * NR, st));
        }
        if (comm_group_begin(s->comm)) return -1;
        int rc = 0;
        for (int p = 0; p < g && !rc; ++p) {
            if (p == me) continue;
            rc = comm_send(s->comm, mine, 2 * (size_t)k * NR, f->gbase + p, st);
            if (!rc) rc = comm_recv(s->comm, f->gat + (size_t)p * k * NR, 2 * (size_t)k * NR, f->gbase + p, st);
        }
        if (comm_group_end(s->comm) || rc) return -1;
        { dist_bwd_finish_kernel<NR><<<ceil_div(k * NR, 128), 128, 0, st>>>(f->vec, f->yE, f->gat, g, k, col0); ++g_fdfd_launches; }
        FDFD_CHECK(cudaGetLastError());
    }
    return 0;
}
