"""One process per GPU: the communicator the sharded paths run on (NCCL over NVLink / NVSwitch,
created inside libfdfd_b200.so; see include/fdfd_b200.h "one grid split over several GPUs").

The reference is single-process; this is the multi-GPU extension of its hot path:

* ``DirectSolver(op, comm=...)``   one grid's elimination tree split over the ranks (subtree per
  GPU, binary reduction tree above; ndplan.shard_plan);
* ``SlabOperator``  the matrix-free stencil and the Krylov solvers on slabs with halo exchange;
  ``SlabOperator.setup_schwarz`` adds the restricted additive Schwarz preconditioner (per-slab direct
  factors with PML transmission) that turns the slab path into a solver for large grids.

The 128-byte NCCL id has to reach every rank once; ``Communicator.from_torch()`` uses an
initialised ``torch.distributed`` process group for that (torch is plumbing here, nothing else),
``Communicator(rank, world, uid)`` takes an id delivered any other way.
"""
import ctypes as C
import glob
import os
import site

import numpy as np

from . import _lib
from ._lib import check


def _find_nccl():
    """The NCCL build torch ships (so both users share one copy when torch is loaded), else the system one."""
    cands = []
    for sp in site.getsitepackages() + [site.getusersitepackages()]:
        cands += glob.glob(os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2"))
    return cands[0] if cands else None


class Communicator:
    def __init__(self, rank, world, uid):
        self.lib = _lib.load()
        _lib.require_gpu()
        path = os.environ.get("FDFD_NCCL_LIB") or _find_nccl()
        check(self.lib.fdfd_comm_load(path.encode() if path else None))
        self.rank, self.world = int(rank), int(world)
        uid = np.ascontiguousarray(np.frombuffer(bytes(uid), dtype=np.uint8))
        if uid.size != 128:
            raise ValueError("the communicator id is 128 bytes")
        self.h = C.c_void_p()
        check(self.lib.fdfd_comm_create(C.byref(self.h), _lib.ptr(uid), self.rank, self.world))

    @staticmethod
    def unique_id():
        lib = _lib.load()
        path = os.environ.get("FDFD_NCCL_LIB") or _find_nccl()
        check(lib.fdfd_comm_load(path.encode() if path else None))
        uid = np.zeros(128, dtype=np.uint8)
        check(lib.fdfd_comm_unique_id(_lib.ptr(uid)))
        return uid.tobytes()

    @classmethod
    def from_torch(cls, group=None):
        """Rank / world / id exchange through a torch.distributed process group (default: the global one)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(rank, world, box[0])

    @classmethod
    def local_group(cls, world):
        """``world`` in-process communicators (one per host thread; see ``run_ranks``): the multi-rank code paths
        on a single GPU, transfers are device copies behind a rendezvous instead of NCCL."""
        lib = _lib.load()
        _lib.require_gpu()
        arr = (C.c_void_p * world)()
        check(lib.fdfd_comm_create_local(arr, int(world)))
        out = []
        for r in range(world):
            c = cls.__new__(cls)
            c.lib, c.rank, c.world, c.h = lib, r, int(world), C.c_void_p(arr[r])
            out.append(c)
        return out

    def abort(self):
        self.lib.fdfd_comm_abort(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.fdfd_comm_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def run_ranks(world, fn):
    """Run ``fn(comm)`` for ``world`` in-process ranks, one thread each (ctypes releases the GIL inside the
    library, so the ranks really overlap); returns the list of results, re-raises the first failure."""
    import threading
    comms = Communicator.local_group(world)
    results, errors = [None] * world, [None] * world

    def body(r):
        try:
            results[r] = fn(comms[r])
        except BaseException as e:          # noqa: B902 - the other ranks must be released whatever went wrong
            errors[r] = e
            comms[r].abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and "aborted" not in str(e):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results


def slab_rows(gnx, world, rank):
    """Rows [x0, x1) of rank ``rank``: contiguous blocks whose sizes differ by at most one."""
    base, extra = divmod(int(gnx), int(world))
    x0 = rank * base + min(rank, extra)
    return x0, x0 + base + (1 if rank < extra else 0)


class SlabOperator:
    """This rank's slab of the Maxwell operator (rows x0..x1 of the reference's (Nx, Ny) arrays) with
    the matrix-free stencil and the Krylov solvers running across the ranks: halo rows travel over
    NCCL send/recv, inner products are all-reduced on the device.

    ``eps_r`` is the WHOLE permittivity array (every rank passes the same one; only its slab and the
    two neighbouring rows go to the device).  Vectors are per-rank slabs of shape (x1 - x0, Ny).
    ``comm=None`` builds a single slab that wraps onto itself (world size 1)."""

    def __init__(self, omega, eps_r, dl, NPML, pol, L0, comm=None, rows=None):
        from .core import POL
        self.lib = _lib.load()
        _lib.require_gpu()
        eps_r = np.asarray(eps_r)
        self.gnx, self.ny = eps_r.shape
        self.comm = comm
        world, rank = (comm.world, comm.rank) if comm is not None else (1, 0)
        self.x0, self.x1 = rows if rows is not None else slab_rows(self.gnx, world, rank)
        self.nxl = self.x1 - self.x0
        self.pol = pol
        self.omega, self.dl, self.L0, self.npml = float(omega), float(dl), float(L0), [int(NPML[0]), int(NPML[1])]
        self._schwarz = None
        self.h = C.c_void_p()
        check(self.lib.fdfd_slab_op_create(C.byref(self.h), comm.h if comm is not None else None, self.gnx, self.ny,
                                           self.x0, self.nxl, float(omega), float(dl), int(NPML[0]), int(NPML[1]),
                                           POL[pol], float(L0)))
        self.assemble(eps_r)

    def _ext_rows(self):
        return np.arange(self.x0 - 1, self.x1 + 1) % self.gnx

    def assemble(self, eps_r, averaging=True):
        ext = _lib.as_c128(np.asarray(eps_r)[self._ext_rows()])
        check(self.lib.fdfd_op_assemble_host(self.h, _lib.ptr(ext), None, int(bool(averaging))))

    def to_ext(self, v):
        v = np.asarray(v)
        if v.shape != (self.nxl, self.ny):
            raise ValueError("slab vectors have shape {}".format((self.nxl, self.ny)))
        ext = np.zeros((self.nxl + 2, self.ny), dtype=np.complex128)
        ext[1:-1] = v
        return ext

    def dot(self, x, fused=False):
        """y = A x on this rank's rows (collective: every rank must call it)."""
        xe = self.to_ext(x)
        ye = np.zeros_like(xe)
        check(self.lib.fdfd_op_apply_host(self.h, _lib.ptr(xe), _lib.ptr(ye), 1, int(fused)))
        return ye[1:-1]

    def krylov(self, b, method="bicgstab", x0=None, tol=1e-10, maxiter=20000, fused=True, check_every=10, restart=50):
        """Distributed BiCGSTAB / COCG / GMRES(restart) on the slabs (collective).  With ``setup_schwarz`` done,
        BiCGSTAB and GMRES are right-preconditioned by the Schwarz preconditioner; GMRES is the robust choice there
        (one preconditioner application per iteration, no breakdowns)."""
        be = self.to_ext(b)
        xe = self.to_ext(np.zeros_like(be[1:-1]) if x0 is None else x0)
        it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
        check(self.lib.fdfd_krylov_solve_host(self.h, None, _lib.ptr(be), _lib.ptr(xe),
                                              {"bicgstab": 0, "cocg": 1, "gmres": 2}[method],
                                              float(tol), int(maxiter), int(bool(fused)),
                                              int(restart if method == "gmres" else check_every),
                                              None, 0, C.byref(it), C.byref(rr), C.byref(conv)))
        return xe[1:-1].copy(), dict(iters=it.value, relres=rr.value, converged=bool(conv.value))

    # ---- restricted additive Schwarz preconditioner: what makes the slab path a solver at scale
    def setup_schwarz(self, eps_r, overlap=4, npml_sub=12, averaging=True):
        """Build, assemble and factorise this rank's subdomain (the slab + ``overlap`` rows of each neighbour +
        ``npml_sub`` rows of artificial PML on both sides, as a local torus) and attach it as the right
        preconditioner of ``krylov(method='bicgstab')``.  ``eps_r`` is the whole permittivity array again.
        Local work only (no collective).  Returns the subdomain's DirectSolver (factor bytes, timings)."""
        from .core import DirectSolver, MaxwellOperator
        self.drop_schwarz()
        ext = int(overlap) + int(npml_sub)
        h = C.c_void_p()
        check(self.lib.fdfd_schwarz_sub_create(C.byref(h), self.h, int(overlap), int(npml_sub)))
        sub = MaxwellOperator._adopt(h, self.nxl + 2 * ext, self.ny, self.omega, self.dl, [0, self.npml[1]], self.pol, self.L0)
        rows = np.arange(self.x0 - ext, self.x1 + ext) % self.gnx
        sub.assemble(np.asarray(eps_r)[rows], averaging=averaging)
        d = DirectSolver(sub)
        d.factor()
        check(self.lib.fdfd_slab_set_schwarz(self.h, sub.h, d.h, int(overlap), int(npml_sub)))
        self._schwarz = (sub, d)
        return d

    def drop_schwarz(self):
        if getattr(self, "_schwarz", None) is not None:
            check(self.lib.fdfd_slab_set_schwarz(self.h, None, None, 0, 0))
            self._schwarz = None

    def __del__(self):
        try:
            self.drop_schwarz()
        except Exception:
            pass
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.fdfd_op_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass
