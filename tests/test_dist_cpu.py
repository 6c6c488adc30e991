"""world_size-2/-4 gloo tests (CPU) of the multi-GPU host logic: the sharded elimination plan
(ndplan.shard_plan) executed by the numpy model with the same point-to-point exchanges the CUDA
library issues over NCCL, and the slab decomposition of the stencil with its halo exchange."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class GlooComm:
    def __init__(self, dist, torch):
        self.dist, self.torch = dist, torch

    def send(self, arr, dst):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.float64).copy())
        self.dist.send(t, dst)

    def recv(self, shape, src):
        out = np.empty(shape, dtype=np.complex128)
        t = self.torch.from_numpy(out.view(np.float64))
        self.dist.recv(t, src)
        return out

    def sendrecv(self, sarr, dst, rshape, src):
        """Both directions at once (non-blocking send), like a grouped NCCL send/recv pair."""
        req = None
        if sarr is not None:
            t = self.torch.from_numpy(np.ascontiguousarray(sarr).view(np.float64).copy())
            req = self.dist.isend(t, dst)
        out = self.recv(rshape, src) if rshape is not None else None
        if req is not None:
            req.wait()
        return out

    def allreduce(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.float64).copy())
        self.dist.all_reduce(t)
        return t.numpy().view(np.complex128)


def _worker_direct(rank, world, port, shape, npml, pol, q, split=None, distribute=False, rb=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        from fdfdpy_b200.ndplan import build_plan, shard_plan
        from oracle import fdfd_oracle as orc
        from tests.nd_model import factor, solve, row_scale
        nx, ny = shape
        rng = np.random.default_rng(0)
        eps = 1 + 5 * rng.random((nx, ny))
        omega = 2 * np.pi * 200e12
        planes = orc.stencil_planes(omega, eps, 0.04, npml, pol, 1e-6)
        isxf, _, isyf, _ = orc.pml_inverse_factors(omega, 1e-6, (nx, ny), npml, 0.04)
        d = row_scale(isxf, isyf)
        full = build_plan(nx, ny) if split is None else build_plan(nx, ny, split_min=split[0], split_parts=split[1])
        levels = shard_plan(full, world, rank, distribute=distribute, rb=rb)
        comm = GlooComm(dist, torch)
        store = factor(levels, planes, nx, ny, d, tile=8, comm=comm)
        b = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
        u = solve(levels, store, b, nx, ny, d, comm=comm)
        ref = orc.sparse_solve(orc.planes_to_csr(planes), b).reshape(nx, ny)
        q.put((rank, float(np.linalg.norm(u - ref) / np.linalg.norm(ref)),
               int(sum(lv.nb * int(lv.kmax) for lv in levels))))
    except Exception as e:                      # surface the failure instead of hanging the peer
        q.put((rank, repr(e), 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,npml,pol,split", [(2, (24, 20), [3, 3], "Ez", None), (2, (19, 33), [0, 4], "Hz", None),
                                                        (4, (32, 28), [3, 3], "Ez", None),
                                                        (4, (32, 28), [3, 3], "Ez", (6, 3))])
def test_sharded_elimination_tree_gloo(world, shape, npml, pol, split):
    """split: separators eliminated piece by piece (chain levels) also on the levels shared between ranks"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_direct, args=(r, world, port, shape, npml, pol, q, split)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, _ in res:
        assert isinstance(err, float), (rank, err)
        assert err < 1e-11, (rank, err)


@pytest.mark.parametrize("world,shape,npml,pol,split,rb", [(2, (24, 20), [3, 3], "Ez", None, None),
                                                           (2, (19, 33), [0, 4], "Hz", (6, 3), 5),
                                                           (4, (32, 28), [3, 3], "Ez", (6, -4), 6),
                                                           (4, (37, 30), [3, 3], "Hz", (8, 2), 4)])
def test_distributed_fronts_gloo(world, shape, npml, pol, split, rb):
    """The DISTRIBUTED top fronts (ndplan.DistFront: block rows dealt to the ranks of a group, all-to-all assembly,
    pivot owner broadcasts Einv, panel all-gather, replicated substitution vectors) executed by the numpy model over
    gloo point-to-point messages, against the oracle's sparse solve."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_direct, args=(r, world, port, shape, npml, pol, q, split, True, rb)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, _ in res:
        assert isinstance(err, float), (rank, err)
        assert err < 1e-11, (rank, err)


def test_dist_front_plan_is_consistent():
    """Host logic of the distributed fronts: compact slots, block structure, child maps, the same front seen by
    every rank of its group."""
    from fdfdpy_b200.ndplan import build_plan, shard_plan
    for (nx, ny, sm, sp, rb) in [(96, 80, None, None, None), (203, 157, 30, -16, 24), (4096, 4096, None, None, None)]:
        levels = build_plan(nx, ny, split_min=sm, split_parts=sp)
        for world in (2, 4, 8):
            plans = [shard_plan(levels, world, r, rb=rb) for r in range(world)]
            assert all(len(p.dist) == int(np.log2(world)) for p in plans)
            for j in range(int(np.log2(world))):
                for r in range(world):
                    df = plans[r].dist[j]
                    assert df.gsize == 2 ** (j + 1) and df.gbase == (r // df.gsize) * df.gsize
                    twin = plans[df.gbase].dist[j]                    # every rank of the group describes the same front
                    for k in ("level0", "nsteps", "n", "m", "kfull"):
                        assert getattr(df, k) == getattr(twin, k)
                    assert np.array_equal(df.bstart, twin.bstart) and np.array_equal(df.bowner, twin.bowner)
                    assert df.bstart[0] == 0 and df.bstart[-1] == df.n and np.all(np.diff(df.bstart) > 0)
                    assert df.bstart[df.nsteps] == df.kfull
                    assert set(df.bowner.tolist()) <= set(range(df.gsize))
                    for c in (0, 1):
                        inv = df.inv[c]
                        hit = np.sort(inv[inv >= 0])
                        assert np.array_equal(hit, np.arange(df.mc[c]))      # every child ring node lands exactly once
                    assert np.all((df.inv[0] >= 0) | (df.inv[1] >= 0))       # every front slot comes from a child
                    # nodes reached by BOTH children are exactly the separator being closed plus shared corners
                    if j > 0:
                        assert plans[r].dist[j - 1].m == df.mc[1 if (r - df.gbase) >= df.gsize // 2 else 0]
                    assert all(plans[r][df.level0 + s].nb == 0 for s in range(df.nsteps))
                # the root front ends with an empty ring
            assert plans[0].dist[-1].m == 0


def test_shard_plan_partitions_the_tree():
    from fdfdpy_b200.ndplan import build_plan, shard_plan
    levels = build_plan(96, 80)
    for world in (1, 2, 4, 8):
        shards = [shard_plan(levels, world, r, distribute=False) for r in range(world)]
        for l, lv in enumerate(levels):
            gids = np.concatenate([s[l].gids for s in shards])
            assert sorted(gids.tolist()) == list(range(lv.nb))          # every front owned exactly once
            sends = sorted((r, s[l].send_to) for r, s in enumerate(shards) if s[l].send_to >= 0)
            recvs = sorted((s[l].recv_from, r) for r, s in enumerate(shards) if s[l].recv_from >= 0)
            assert sends == recvs                                        # every send has its receive
        if world > 1:
            assert sum(1 for lv in shards[0] if lv.recv_from >= 0) == int(np.log2(world))
    with pytest.raises(ValueError):
        shard_plan(levels, 3, 0)


def _worker_slab(rank, world, port, shape, npml, pol, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        from fdfdpy_b200.distributed import slab_rows
        from oracle import fdfd_oracle as orc
        nx, ny = shape
        rng = np.random.default_rng(2)
        eps = 1 + 5 * rng.random((nx, ny))
        x = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
        omega = 2 * np.pi * 200e12
        c0, cxm, cxp, cym, cyp = orc.stencil_planes(omega, eps, 0.04, npml, pol, 1e-6)
        x0, x1 = slab_rows(nx, world, rank)
        # the slab in the library's extended layout: one halo row on each side, filled by the neighbours
        ext = np.zeros((x1 - x0 + 2, ny), dtype=np.complex128)
        ext[1:-1] = x[x0:x1]
        lower, upper = (rank - 1) % world, (rank + 1) % world
        comm = GlooComm(dist, torch)
        # same order as op_halo_exchange: first row down / upper halo in, then last row up / lower halo in
        reqs = [dist.isend(torch.from_numpy(ext[1].view(np.float64).copy()), lower)]
        ext[-1] = comm.recv((ny,), upper)
        reqs.append(dist.isend(torch.from_numpy(ext[-2].view(np.float64).copy()), upper))
        ext[0] = comm.recv((ny,), lower)
        for r in reqs:
            r.wait()
        sl = slice(x0, x1)
        y = (c0[sl] * ext[1:-1] + cxm[sl] * ext[:-2] + cxp[sl] * ext[2:] +
             cym[sl] * np.roll(ext[1:-1], 1, axis=1) + cyp[sl] * np.roll(ext[1:-1], -1, axis=1))
        ref = orc.planes_to_csr((c0, cxm, cxp, cym, cyp)).dot(x.ravel()).reshape(nx, ny)[sl]
        # inner product summed over the ranks equals the global one
        part = np.array([np.vdot(x[sl], y)])
        tot = comm.allreduce(part)[0]
        gref = np.vdot(x, orc.planes_to_csr((c0, cxm, cxp, cym, cyp)).dot(x.ravel()).reshape(nx, ny))
        q.put((rank, float(np.linalg.norm(y - ref) / np.linalg.norm(ref)), float(abs(tot - gref) / abs(gref))))
    except Exception as e:
        q.put((rank, repr(e), 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,pol", [(2, (21, 16), "Ez"), (3, (20, 12), "Hz")])
def test_slab_halo_exchange_gloo(world, shape, pol):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_slab, args=(r, world, port, shape, [3, 3], pol, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, derr in res:
        assert isinstance(err, float), (rank, err)
        assert err < 1e-13 and derr < 1e-12, (rank, err, derr)


def test_slab_rows_cover_the_grid():
    from fdfdpy_b200.distributed import slab_rows
    for gnx in (7, 64, 100, 4096):
        for world in (1, 2, 3, 8):
            rows = [slab_rows(gnx, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == gnx
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1


def test_schwarz_model_converges_and_matches_direct():
    """The restricted additive Schwarz preconditioner of the slab path, as a numpy model (tests/schwarz_model.py: the
    subdomains, overlap, artificial PML and stretch factors of csrc/operator.cu / csrc/krylov.cu): preconditioned
    GMRES reaches the oracle's direct solution on a waveguide device in a few iterations per slab."""
    from oracle import fdfd_oracle as orc
    from tests import schwarz_model as sm
    omega, dl, L0, npml = 2 * np.pi * 200e12, 0.02, 1e-6, [10, 10]
    nx, ny = 192, 96
    eps = sm.device_eps(nx, ny)
    b = np.zeros((nx, ny), complex)
    b[nx // 5, ny // 2] = 1j * omega
    x, its, rr, A = sm.solve(omega, eps, dl, npml, L0, b, slabs=3, overlap=4, npml_s=10, tol=1e-11)
    A0 = orc.construct_A(omega, eps, dl, npml, "Ez", L0)
    assert abs(A - A0).max() <= 1e-12 * abs(A0).max()          # the model's global operator IS the oracle's
    ref = orc.sparse_solve(A0, b).reshape(nx, ny)
    assert rr < 1e-10 and its <= 30, (its, rr)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-8


def _worker_schwarz(rank, world, port, shape, overlap, npml_s, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        from tests import schwarz_model as sm
        omega, dl, L0, npml = 2 * np.pi * 200e12, 0.03, 1e-6, [6, 5]
        nx, ny = shape
        rng = np.random.default_rng(3)
        eps = 1 + 8 * (rng.random((nx, ny)) > 0.7)
        r = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
        _, M = sm.build(omega, eps, dl, npml, L0, world, overlap, npml_s)        # the whole preconditioner, one process
        ref = M(r.ravel()).reshape(nx, ny)
        sub = sm.build_rank(omega, eps, dl, npml, L0, world, rank, overlap, npml_s)
        z = sm.rank_apply(sub, r[sub["x0"]:sub["x1"]], GlooComm(dist, torch), rank, world, overlap, npml_s)
        q.put((rank, float(np.linalg.norm(z - ref[sub["x0"]:sub["x1"]]) / np.linalg.norm(ref[sub["x0"]:sub["x1"]]))))
    except Exception as e:
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,overlap,npml_s", [(2, (40, 24), 3, 5), (3, (47, 20), 2, 4), (4, (64, 18), 4, 6)])
def test_schwarz_apply_rank_local_gloo(world, shape, overlap, npml_s):
    """One application of the Schwarz preconditioner computed rank by rank (each rank: its rows, its subdomain
    factors, the overlap rows exchanged with both neighbours over gloo in the order csrc/krylov.cu schwarz_apply
    uses over NCCL, world 2 included where both neighbours are the same peer) equals the single-process model."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_schwarz, args=(r, world, port, shape, overlap, npml_s, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err in res:
        assert isinstance(err, float), (rank, err)
        assert err < 1e-12, (rank, err)
