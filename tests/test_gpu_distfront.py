"""The one-grid-on-several-ranks paths on a SINGLE GPU: world = 2, 4, 8 ranks run as threads of this process over
the library's in-process communicator (same call surface as the NCCL one, transfers are device copies behind a
rendezvous).  Everything above the transport is the code the NCCL runs execute: the sharded elimination tree with its
DISTRIBUTED top fronts (block-row ownership, personalised all-to-all assembly, panel all-gather, per-rank Schur
updates, replicated substitution vectors), the older owner-per-front scheme, and the slab stencil / Krylov loop.
Parity is against the CPU oracle at north_star's bars (fields 1e-8, residual 1e-10)."""
import os

import numpy as np
import pytest

from oracle import fdfd_oracle as orc

pytestmark = pytest.mark.gpu

OMEGA = 2 * np.pi * 200e12


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


class env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _sharded_solve(world, eps, npml, pol, b, dl=0.04):
    from fdfdpy_b200 import core
    from fdfdpy_b200.distributed import run_ranks

    def rank_body(comm):
        op = core.MaxwellOperator(OMEGA, eps, dl, npml, pol, 1e-6)
        d = core.DirectSolver(op, comm=comm)
        x = np.array(d.solve(b)).reshape(b.shape)
        x2 = np.array(d.solve(b[0] if b.ndim == 3 else b)).reshape(eps.shape)      # cached factors, second call
        return x, x2, d.last_relres, len(getattr(d.levels, "dist", ())), d.stats()["factor_bytes"]

    return run_ranks(world, rank_body)


@pytest.mark.parametrize("world,pol,shape,knobs", [
    (2, "Ez", (200, 160), {}),
    (2, "Hz", (96, 120), {}),
    (4, "Ez", (200, 160), {}),
    (8, "Ez", (200, 160), {}),
    # separators eliminated piece by piece (chain steps) and rings cut into several row blocks per rank
    (2, "Ez", (200, 160), {"FDFD_SPLIT_MIN": 40, "FDFD_SPLIT_PARTS": -24, "FDFD_DIST_RB": 32}),
    (4, "Hz", (200, 160), {"FDFD_SPLIT_MIN": 40, "FDFD_SPLIT_PARTS": -24, "FDFD_DIST_RB": 32}),
    (8, "Ez", (203, 157), {"FDFD_SPLIT_MIN": 30, "FDFD_SPLIT_PARTS": -16, "FDFD_DIST_RB": 24}),   # ragged: shape classes differ
    (4, "Ez", (131, 250), {"FDFD_SPLIT_MIN": 60, "FDFD_SPLIT_PARTS": 3, "FDFD_DIST_RB": 1000}),
])
def test_distributed_fronts_vs_oracle(world, pol, shape, knobs):
    rng = np.random.default_rng(1)
    nx, ny = shape
    eps = 1 + 5 * (rng.random((nx, ny)) > 0.5)
    npml = [10, 8]
    b = rng.standard_normal((3, nx, ny)) + 1j * rng.standard_normal((3, nx, ny))
    with env(**knobs):
        res = _sharded_solve(world, eps, npml, pol, b)
    A = orc.construct_A(OMEGA, eps, 0.04, npml, pol, 1e-6)
    ref = np.stack([orc.sparse_solve(A, b[j]).reshape(nx, ny) for j in range(3)])
    for r, (x, x2, relres, ndist, fbytes) in enumerate(res):
        assert ndist == int(np.log2(world)), "every shared merge must be a distributed front"
        assert relres < 1e-10, (r, relres)
        assert relerr(x, ref) < 1e-8, r
        assert relerr(x2, ref[0]) < 1e-8, r
        assert np.array_equal(x, res[0][0]), "replicated results must agree bit for bit across the ranks"
    # the factors are split, not replicated: no rank holds more than ~2/world of the single-GPU factor bytes at 8 ranks
    total = sum(r[4] for r in res)
    assert max(r[4] for r in res) < 0.75 * total if world >= 4 else True


def test_distributed_fronts_with_lookahead_size():
    """A grid whose top separators are cut into pieces of several hundred nodes (the production regime: recursive
    block inversion of the pivot blocks, look-ahead on the side stream, persistent GEMM kernel), 2 and 4 ranks,
    against the single-rank solver and the residual contract."""
    from fdfdpy_b200 import core
    rng = np.random.default_rng(5)
    nx, ny = 640, 600
    eps = 1 + 5 * (rng.random((nx, ny)) > 0.6)
    b = rng.standard_normal((2, nx, ny)) + 1j * rng.standard_normal((2, nx, ny))
    op = core.MaxwellOperator(OMEGA, eps, 0.03, [10, 12], "Ez", 1e-6)
    ref = np.array(op.direct().solve(b)).reshape(b.shape)
    assert op.direct().last_relres < 1e-10
    for world in (2, 4):
        res = _sharded_solve(world, eps, [10, 12], "Ez", b, dl=0.03)
        for x, x2, relres, ndist, _ in res:
            assert relres < 1e-10
            assert relerr(x, ref) < 1e-9, world
            assert np.array_equal(x, res[0][0])
    with env(FDFD_LOOKAHEAD=0):
        res0 = _sharded_solve(2, eps, [10, 12], "Ez", b, dl=0.03)
    assert relerr(res0[0][0], ref) < 1e-9


def test_owner_per_front_scheme_still_works():
    """FDFD_DIST_FRONTS=0: the shared fronts live on the lowest rank of their group (round 1's scheme, one Schur block
    per level over the wire)."""
    rng = np.random.default_rng(2)
    nx, ny = 120, 96
    eps = 1 + 5 * rng.random((nx, ny))
    b = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    with env(FDFD_DIST_FRONTS=0):
        res = _sharded_solve(4, eps, [8, 8], "Ez", b)
    ref = orc.sparse_solve(orc.construct_A(OMEGA, eps, 0.04, [8, 8], "Ez", 1e-6), b).reshape(nx, ny)
    for x, x2, relres, ndist, _ in res:
        assert ndist == 0
        assert relres < 1e-10 and relerr(x, ref) < 1e-8


def test_distributed_fronts_multi_rhs_chunks():
    """27 right-hand sides = chunks of 16 + 8 + 2 + 1 through the distributed substitution."""
    rng = np.random.default_rng(3)
    nx, ny = 112, 104
    eps = 1 + 5 * rng.random((nx, ny))
    b = rng.standard_normal((27, nx, ny)) + 1j * rng.standard_normal((27, nx, ny))
    with env(FDFD_SPLIT_MIN=40, FDFD_SPLIT_PARTS=-20, FDFD_DIST_RB=28):
        res = _sharded_solve(4, eps, [8, 8], "Hz", b)
    A = orc.construct_A(OMEGA, eps, 0.04, [8, 8], "Hz", 1e-6)
    for j in (0, 7, 15, 16, 23, 24, 25, 26):
        ref = orc.sparse_solve(A, b[j]).reshape(nx, ny)
        assert relerr(res[0][0][j], ref) < 1e-8, j
        assert relerr(res[3][0][j], ref) < 1e-8, j


@pytest.mark.parametrize("world,pol", [(2, "Ez"), (3, "Hz"), (4, "Ez")])
def test_slab_operator_in_process_ranks(world, pol):
    """Slab decomposition of the matrix-free stencils (halo exchange, all-reduced inner products) over the in-process
    transport: apply parity 1e-13, distributed BiCGSTAB / COCG to the oracle's solution 1e-8."""
    from fdfdpy_b200.distributed import SlabOperator, run_ranks
    rng = np.random.default_rng(7)
    nx, ny = 41, 36
    eps = 1 + 2 * rng.random((nx, ny))
    npml = [8, 8]
    xv = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
    A = orc.construct_A(OMEGA, eps, 0.05, npml, pol, 1e-6)
    ref = A.dot(xv.ravel()).reshape(nx, ny)
    b = np.zeros((nx, ny), dtype=complex)
    b[nx // 2, ny // 2] = 1j * OMEGA
    sol = orc.sparse_solve(A, b).reshape(nx, ny)

    def rank_body(comm):
        slab = SlabOperator(OMEGA, eps, 0.05, npml, pol, 1e-6, comm=comm)
        sl = slice(slab.x0, slab.x1)
        out = {"rows": (slab.x0, slab.x1), "planes": slab.dot(xv[sl]), "fused": slab.dot(xv[sl], fused=True)}
        for method in ("bicgstab", "cocg"):
            xs, info = slab.krylov(b[sl], method=method, tol=1e-12, maxiter=20000, check_every=20)
            out[method] = (xs, info)
        return out

    res = run_ranks(world, rank_body)
    for method in ("planes", "fused"):
        full = np.concatenate([r[method] for r in res])
        assert relerr(full, ref) < 1e-13, method
    for method in ("bicgstab", "cocg"):
        full = np.concatenate([r[method][0] for r in res])
        assert all(r[method][1]["relres"] < 1e-10 for r in res), [r[method][1] for r in res]
        assert relerr(full, sol) < 1e-8, method


@pytest.mark.parametrize("world,pol,overlap", [(1, "Ez", 4), (2, "Ez", 4), (4, "Ez", 4), (4, "Hz", 3), (3, "Ez", 2)])
def test_slab_schwarz_preconditioner(world, pol, overlap):
    """Restricted additive Schwarz on the slabs (per-slab direct factors, artificial PML at the cut faces): the
    preconditioned distributed BiCGSTAB reaches north_star's bars (fields 1e-8 against the oracle's direct solve,
    residual 1e-10) on a strongly scattering grid in a small fraction of the unpreconditioned iterations."""
    from fdfdpy_b200.distributed import SlabOperator, run_ranks
    rng = np.random.default_rng(11)
    nx, ny = 232, 96
    eps = 1 + 11 * (rng.random((nx // 8, ny // 8)) > 0.6).repeat(8, 0).repeat(8, 1)
    npml = [10, 10]
    dl = 0.04
    A = orc.construct_A(OMEGA, eps, dl, npml, pol, 1e-6)
    b = np.zeros((nx, ny), dtype=complex)
    b[nx // 2, ny // 2] = 1j * OMEGA
    b[nx // 5, ny // 3] = -0.5j * OMEGA
    sol = orc.sparse_solve(A, b).reshape(nx, ny)

    def rank_body(comm):
        slab = SlabOperator(OMEGA, eps, dl, npml, pol, 1e-6, comm=comm if world > 1 else None)
        sl = slice(slab.x0, slab.x1)
        d = slab.setup_schwarz(eps, overlap=overlap, npml_sub=10)
        xs, info = slab.krylov(b[sl], method="bicgstab", tol=1e-11, maxiter=400, check_every=2)
        xg, infog = slab.krylov(b[sl], method="gmres", tol=1e-11, maxiter=400, restart=60)
        slab.drop_schwarz()
        _, plain = slab.krylov(b[sl], method="bicgstab", tol=1e-11, maxiter=400, check_every=50)
        return xs, info, plain, d.stats()["factor_bytes"], xg, infog

    res = run_ranks(world, rank_body)
    full = np.concatenate([r[0] for r in res])
    info = res[0][1]
    assert all(r[1]["relres"] < 1e-10 and r[1]["converged"] for r in res), [r[1] for r in res]
    assert relerr(full, sol) < 1e-8
    assert info["iters"] <= 40 * max(world, 2), info      # measured: 22 / 39 / 86 iterations for 2 / 4 / 3 (overlap 2) slabs
    assert not res[0][2]["converged"] or res[0][2]["iters"] > 4 * info["iters"], (info, res[0][2])
    # GMRES on the same preconditioner: one application per iteration, so at most BiCGSTAB's application count
    infog = res[0][5]
    assert all(r[5]["relres"] < 1e-10 and r[5]["converged"] for r in res), [r[5] for r in res]
    assert relerr(np.concatenate([r[4] for r in res]), sol) < 1e-8
    assert infog["iters"] <= 2 * info["iters"], (infog, info)
