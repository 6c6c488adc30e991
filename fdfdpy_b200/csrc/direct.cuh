// Structured direct solver: batched multifrontal nested dissection on the periodic grid.
// Plan semantics are documented in fdfdpy_b200/ndplan.py; this side owns the numerics.
#pragma once
#include <vector>
#include "comm.cuh"
#include "operator.cuh"

struct NdLevelDesc {          // host view handed over the C ABI, all arrays int32 host pointers
    int kind;                 // 0 leaf, 1 merge
    int nb, kmax, mmax, ncls, child_mmax;
    const int* cls;           // [nb]
    const int* k_cls;         // [ncls]
    const int* ch1;           // [nb]      (merge)
    const int* ch2;           // [nb]      (merge)
    const int* c1map;         // [ncls][child_mmax] (merge)
    const int* c2map;         // [ncls][child_mmax] (merge)
    const int* x0;            // [nb]      (leaf)
    const int* y0;            // [nb]      (leaf)
    const int* slot_lx;       // [ncls][nmax] (leaf)
    const int* slot_ly;
    const int* slot_right;
    const int* slot_up;
    int send_to, recv_from;   // sharded tree (ndplan.shard_plan): peer ranks of this level's exchange, -1 = none
};

struct NdLevel {
    int kind, nb, kmax, mmax, nmax, ncls, child_mmax;
    int *cls, *k_cls, *ch1, *ch2, *c1map, *c2map, *inv1, *inv2;
    int *x0, *y0, *slot_lx, *slot_ly, *slot_right, *slot_up;
    cplx* Einv;   // [nb][kmax][kmax]   F_EE^-1 of the row-scaled (symmetric) front, full storage
    cplx* G;      // [nb][mmax][kmax]   F_RE F_EE^-1
    cplx* yE;     // solve workspace [nb][kmax][nrhs]
    size_t ye_off;
    int send_to, recv_from;
    int inplace;  // chain level factorised in place on the previous level's Schur blocks (no assembly pass)
};

// A front of a level shared by several ranks, stored and factorised DISTRIBUTED by block rows (ndplan.DistFront).
struct NdDistFrontDesc {      // host view handed over the C ABI
    int level0, nsteps;       // plan levels [level0, level0 + nsteps) are this front's elimination steps
    int gbase, gsize;         // ranks gbase .. gbase + gsize - 1 share it
    int n, nblk;              // compact front size, number of row blocks
    const int* bstart;        // [nblk + 1] first slot of every block (blocks 0 .. nsteps-1 are the pivot blocks)
    const int* bowner;        // [nblk]     group rank that owns the block's rows
    int mc1, mc2;             // ring sizes of the two children
    const int* inv1;          // [n] front slot -> position in child 1's ring (-1: none)
    const int* inv2;
};

struct NdDistFront {
    int level0, nsteps, gbase, gsize, grank, cidx, n, nblk, kfull, m, kmax_step;
    std::vector<int> bstart, bowner, lrow0;   // lrow0[j]: first local row of block j (-1: not mine)
    std::vector<int> nloc_of;                 // local row count of every group rank
    int nloc;
    int mc[2];
    int *d_inv[2];                            // [n] each
    int* d_cmap_mine;                         // [mc[cidx]] my child's ring position -> front slot
    std::vector<int*> d_rows_of;              // per group rank: the front slots of its local rows
    int* d_lslot;                             // = d_rows_of[grank]
    long long* d_ring_off;                    // [m] element offset of ring row a inside F (incl. the column origin), -1: not mine
    std::vector<int> r0;                      // per step: first local row below the pivot block
    std::vector<cplx*> Einv, G;               // per step: Einv [k][k] (replicated), G [(nloc - r0)][k]
    cplx* F;                                  // my rows of the front, [nloc][n], lower part valid
    cplx *vec, *yE, *oring, *gat;             // solve workspace: front vector [n][8], yE [kfull][8], other child's ring, gather buffer
};

struct NdSolver {
    int nx, ny;
    int tile;                         // (kept for ABI compatibility; the block inversion uses 64-wide base tiles)
    std::vector<NdLevel> levels;
    bool factored;
    const void* fact_op;              // operator and assembly version the factors belong to (fdfd_factor_solve_fields_host)
    unsigned long long fact_version;
    size_t factor_bytes;
    double factor_flops;              // real flops of the last factorisation (8 per complex MAC)
    int* d_info;                      // device flag: non-zero if a pivot tile was singular
    // factorisation arena (allocated once): ping-pong front batches + block-inversion workspace stack
    cplx *fws, *fws_F[2], *fws_W;
    size_t fws_cap;
    // solve workspace (grown on demand)
    cplx *ws_a, *ws_b, *ws_ring_a, *ws_ring_b, *ws_ye;
    size_t ws_vec_cap, ws_ring_cap, ws_ye_cap;
    cplx* ws_refine;                  // iterative-refinement residual / correction vectors
    size_t ws_refine_cap;
    cplx* ws_bsplit;                  // partial sums of the ring-split backward product (top levels)
    size_t ws_bsplit_cap;
    // sharded tree: communicator (not owned) and the packed Schur block in flight between two ranks
    FdfdComm* comm;
    cplx* xchg;
    size_t xchg_cap;
    // look-ahead on chain levels: the next pivot block is inverted on a side stream under the current Schur update
    cudaStream_t la_stream;
    cudaEvent_t la_ready, la_done;
    cudaEvent_t zg_fork, zg_join;     // helper launch of a Schur update on the look-ahead stream (zgemm.cuh)
    // distributed fronts of the shared levels, in elimination order (front j + 1 is the parent of front j)
    std::vector<NdDistFront*> dist;
    cplx *dist_send, *dist_recv, *dist_panel;
    size_t dist_send_cap, dist_recv_cap, dist_panel_cap;
};

int nd_create(NdSolver** out, int nx, int ny, int tile);
int nd_add_level(NdSolver* s, const NdLevelDesc* d);
int nd_add_dist_front(NdSolver* s, const NdDistFrontDesc* d);   // after fdfd_direct_set_comm and all levels
void nd_destroy(NdSolver* s);
int nd_factor(NdSolver* s, const FdfdOp* op, bool defer_check = false);
int nd_factor_check(NdSolver* s, const FdfdOp* op);     // after a deferred factorisation: sync + singular-pivot flag
// d_b, d_x: [nrhs][nx*ny] device vectors
int nd_solve(NdSolver* s, const FdfdOp* op, const cplx* d_b, cplx* d_x, int nrhs);
