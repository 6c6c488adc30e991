"""Python handles over the C ABI: the device-resident Maxwell operator and the direct solver.

These are the objects the reference-shaped API in simulation.py / linalg.py is built from.
Nothing here computes on the CPU: arrays are marshalled to libfdfd_b200.so and back.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_c128, as_i32, check, pinned_empty, ptr
from .ndplan import build_plan

POL = {"Ez": 0, "Hz": 1}
_plan_cache = {}


def get_plan(nx, ny, sharded=False):
    """Elimination plan of an nx x ny grid (cached).  FDFD_SPLIT_MIN / FDFD_SPLIT_PARTS / FDFD_SPLIT_MAX_STEPS override
    the separator splitting of ndplan.build_plan for A/B measurements.  ``sharded``: the plan of a tree split over
    several GPUs eliminates the top separators in up to 16 pieces instead of 8 (see ndplan.SPLIT_MAX_STEPS)."""
    import os
    sm, sp = os.environ.get("FDFD_SPLIT_MIN"), os.environ.get("FDFD_SPLIT_PARTS")
    ms = os.environ.get("FDFD_SPLIT_MAX_STEPS")
    max_steps = int(ms) if ms else (16 if sharded else None)
    key = (int(nx), int(ny), sm, sp, max_steps)
    if key not in _plan_cache:
        _plan_cache[key] = build_plan(int(nx), int(ny), split_min=int(sm) if sm else None,
                                      split_parts=int(sp) if sp else None, split_max_steps=max_steps)
    return _plan_cache[key]


class MaxwellOperator:
    """A = Dxf mu^-1 Dxb + Dyf mu^-1 Dyb + w^2 eps on the device (reference: linalg.py:39 construct_A).

    Behaves like the scipy matrix the reference keeps in ``Simulation.A`` as far as the hot path
    needs: ``shape``, ``dot``; ``to_scipy()`` exports it as CSR for inspection.
    """

    def __init__(self, omega, eps_r, dl, NPML, pol, L0, averaging=True, eps_nl=None):
        _lib.require_gpu()
        self.lib = _lib.load()
        if pol not in POL:
            raise ValueError("something went wrong and pol is not one of Ez, Hz, instead was given {}".format(pol))
        eps_r = np.asarray(eps_r)
        if eps_r.ndim != 2:
            raise ValueError("eps_r must be a 2-D array")
        self.nx, self.ny = eps_r.shape
        self.shape = (self.nx * self.ny, self.nx * self.ny)
        self.pol = pol
        self.omega, self.dl, self.L0 = float(omega), float(dl), float(L0)
        self.NPML = [int(NPML[0]), int(NPML[1])]
        self.h = C.c_void_p()
        check(self.lib.fdfd_op_create(C.byref(self.h), self.nx, self.ny, self.omega, self.dl, self.NPML[0],
                                      self.NPML[1], POL[pol], self.L0))
        self._direct = None
        self.preconditioner = None      # optional DirectSolver of a nearby operator (same grid)
        self.assemble(eps_r, eps_nl, averaging)

    @classmethod
    def _adopt(cls, handle, nx, ny, omega, dl, NPML, pol, L0):
        """Wrap an operator handle the library created itself (fdfd_schwarz_sub_create); not assembled yet."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.nx, self.ny = int(nx), int(ny)
        self.shape = (self.nx * self.ny, self.nx * self.ny)
        self.pol = pol
        self.omega, self.dl, self.L0 = float(omega), float(dl), float(L0)
        self.NPML = [int(NPML[0]), int(NPML[1])]
        self.h = handle
        self._direct = None
        self.preconditioner = None
        self.has_nl = False
        return self

    def assemble(self, eps_r, eps_nl=None, averaging=True):
        eps_r = np.asarray(eps_r)
        if eps_r.shape != (self.nx, self.ny):
            raise ValueError("eps_r shape changed; build a new operator")
        if eps_nl is None and np.isrealobj(eps_r):
            # real permittivity goes over PCIe as float64 and is widened on the device
            er = np.ascontiguousarray(eps_r, dtype=np.float64)
            check(self.lib.fdfd_op_assemble_host_f64(self.h, ptr(er), int(bool(averaging))))
            en = None
        else:
            er = as_c128(eps_r)
            en = None if eps_nl is None else as_c128(np.broadcast_to(eps_nl, er.shape))
            check(self.lib.fdfd_op_assemble_host(self.h, ptr(er), ptr(en), int(bool(averaging))))
        self.has_nl = en is not None
        if self._direct is not None:
            self._direct.factored = False

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.fdfd_op_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    # ---- matrix-like surface
    def dot(self, x, fused=False):
        if np.asarray(x).dtype == np.complex64:          # complex64 storage, fp64 arithmetic
            x = np.ascontiguousarray(x)
            if x.size != self.nx * self.ny:
                raise ValueError("complex64 apply takes one vector")
            y = np.empty_like(x)
            check(self.lib.fdfd_op_apply_host_c64(self.h, ptr(x), ptr(y), int(bool(fused))))
            return y
        x = as_c128(x)
        nvec = x.size // (self.nx * self.ny)
        y = pinned_empty(x.shape)
        check(self.lib.fdfd_op_apply_host(self.h, ptr(x), ptr(y), nvec, int(fused)))
        return y

    def planes(self):
        out = np.empty((5, self.nx, self.ny), dtype=np.complex128)
        check(self.lib.fdfd_op_get_planes_host(self.h, ptr(out)))
        return out

    def eps_flags(self):
        """Device-side flags of the assembled permittivity: bit 0 = complex entries, bit 1 = negative real parts."""
        f = C.c_int(0)
        check(self.lib.fdfd_op_eps_flags(self.h, C.byref(f)))
        return f.value

    def sfactors(self):
        arrs = [np.empty(n, dtype=np.complex128) for n in (self.nx, self.nx, self.ny, self.ny)]
        check(self.lib.fdfd_op_get_sfactors_host(self.h, *[ptr(a) for a in arrs]))
        return tuple(arrs)

    def to_scipy(self, matrix_format="csr"):
        """Export A as a scipy sparse matrix (format conversion of the device planes, no solve)."""
        import scipy.sparse as sp
        c0, cxm, cxp, cym, cyp = self.planes()
        nx, ny = self.nx, self.ny
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        row = (ii * ny + jj).ravel()
        cols = [row, (((ii - 1) % nx) * ny + jj).ravel(), (((ii + 1) % nx) * ny + jj).ravel(),
                (ii * ny + (jj - 1) % ny).ravel(), (ii * ny + (jj + 1) % ny).ravel()]
        vals = np.concatenate([p.ravel() for p in (c0, cxm, cxp, cym, cyp)])
        A = sp.coo_matrix((vals, (np.tile(row, 5), np.concatenate(cols))), shape=self.shape)
        return A.asformat(matrix_format)

    def derive_fields(self, X, averaging=None):
        """In-plane fields; ``averaging`` overrides the operator's Hz edge-averaging flag."""
        X = as_c128(X)
        f1, f2 = pinned_empty(X.shape), pinned_empty(X.shape)
        av = -1 if averaging is None else int(bool(averaging))
        check(self.lib.fdfd_op_derive_fields_host(self.h, ptr(X), ptr(f1), ptr(f2), av))
        return f1.reshape(self.nx, self.ny), f2.reshape(self.nx, self.ny)

    # ---- solvers
    def direct(self, tile=64):
        if self._direct is None:
            self._direct = DirectSolver(self, tile=tile)
        return self._direct

    def solve(self, b, max_refine=3, tol=1e-12):
        """Direct solve with the cached factorisation (factorises on first use)."""
        return self.direct().solve(b, max_refine=max_refine, tol=tol)

    def krylov(self, b, method="bicgstab", x0=None, tol=1e-10, maxiter=20000, fused=True, check_every=10,
               precondition=False, c12=None, real_inner=False, restart=50):
        """Krylov solve of A x (+ c12 conj(x)) = b.  ``precondition=True`` uses whatever factorisation
        the direct-solver handle currently caches (it may belong to a nearby operator).
        A complex64 ``b`` selects complex64 vector storage (fp64 arithmetic and scalars).
        ``method='gmres'``: right-preconditioned GMRES(``restart``), one preconditioner application per iteration."""
        if np.asarray(b).dtype == np.complex64:
            if precondition or c12 is not None:
                raise ValueError("the complex64 solver takes neither a preconditioner nor an anti-linear term")
            b = np.ascontiguousarray(b)
            x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0, dtype=np.complex64).copy()
            it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
            check(self.lib.fdfd_krylov_solve_host_c64(self.h, ptr(b), ptr(x), {"bicgstab": 0, "cocg": 1}[method],
                                                      float(tol), int(maxiter), int(bool(fused)),
                                                      int(check_every), C.byref(it), C.byref(rr), C.byref(conv)))
            return x.reshape(b.shape), dict(iters=it.value, relres=rr.value, converged=bool(conv.value))
        b = as_c128(b)
        x = np.zeros_like(b) if x0 is None else as_c128(x0).copy()
        it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
        pre = None
        if precondition:
            d = self.preconditioner if self.preconditioner is not None else self.direct()
            if not d.has_factors:
                d.factor()
            pre = d.h
        c12a = None if c12 is None else as_c128(c12)
        check(self.lib.fdfd_krylov_solve_host(self.h, pre, ptr(b), ptr(x), {"bicgstab": 0, "cocg": 1, "gmres": 2}[method],
                                              float(tol), int(maxiter), int(fused),
                                              int(restart if method == "gmres" else check_every), ptr(c12a),
                                              int(bool(real_inner)), C.byref(it), C.byref(rr), C.byref(conv)))
        return x.reshape(b.shape), dict(iters=it.value, relres=rr.value, converged=bool(conv.value))


class DirectSolver:
    """Structured direct solver handle (reference: linalg.py:123 solver_direct / pardisoSolver)."""

    def __init__(self, op, tile=64, comm=None):
        """``comm`` (fdfdpy_b200.distributed.Communicator): this rank's shard of ONE grid's elimination
        tree; operator and right-hand sides are replicated, the factors are split over the ranks."""
        self.op = op
        self.lib = op.lib
        self.h = C.c_void_p()
        check(self.lib.fdfd_direct_create(C.byref(self.h), op.nx, op.ny, int(tile)))
        self.levels = get_plan(op.nx, op.ny, sharded=comm is not None and comm.world > 1)
        self.comm = comm
        if comm is not None and comm.world > 1:
            import os
            from .ndplan import shard_plan
            # FDFD_DIST_FRONTS=0: shared fronts live on the lowest rank of their group (the older scheme);
            # FDFD_DIST_RB: rows per ring block of a distributed front
            rb = os.environ.get("FDFD_DIST_RB")
            self.levels = shard_plan(self.levels, comm.world, comm.rank,
                                     distribute=os.environ.get("FDFD_DIST_FRONTS", "1") != "0",
                                     rb=int(rb) if rb else None)
            check(self.lib.fdfd_direct_set_comm(self.h, comm.h))
        keep = []
        for lv in self.levels:
            d = _lib.LevelDesc()
            d.kind = 0 if lv.kind == "leaf" else 1
            d.nb, d.kmax, d.mmax, d.ncls = lv.nb, lv.kmax, lv.mmax, lv.ncls
            d.child_mmax = getattr(lv, "child_mmax", 0)
            d.send_to, d.recv_from = getattr(lv, "send_to", -1), getattr(lv, "recv_from", -1)
            names = ["cls", "k_cls"] + (["x0", "y0", "slot_lx", "slot_ly", "slot_right", "slot_up"]
                                        if lv.kind == "leaf" else ["ch1", "ch2", "c1map", "c2map"])
            for nme in names:
                arr = as_i32(getattr(lv, nme))
                keep.append(arr)
                setattr(d, nme, arr.ctypes.data_as(C.POINTER(C.c_int)))
            check(self.lib.fdfd_direct_add_level(self.h, C.byref(d)))
        for df in getattr(self.levels, "dist", ()):
            d = _lib.DistFrontDesc()
            d.level0, d.nsteps, d.gbase, d.gsize, d.n, d.nblk = df.level0, df.nsteps, df.gbase, df.gsize, df.n, len(df.bowner)
            d.mc1, d.mc2 = df.mc
            for nme, arr in (("bstart", df.bstart), ("bowner", df.bowner), ("inv1", df.inv[0]), ("inv2", df.inv[1])):
                arr = as_i32(arr)
                keep.append(arr)
                setattr(d, nme, arr.ctypes.data_as(C.POINTER(C.c_int)))
            check(self.lib.fdfd_direct_add_dist_front(self.h, C.byref(d)))
        del keep
        self.factored = False       # cached factors belong to the operator's CURRENT planes
        self.has_factors = False    # some factorisation is cached (possibly of an earlier operator state)
        self.last_relres = None
        self.last_refine_steps = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.fdfd_direct_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def factor(self):
        check(self.lib.fdfd_direct_factor(self.h, self.op.h))
        self.factored = True
        self.has_factors = True

    def stats(self):
        fb, ff = C.c_double(0), C.c_double(0)
        check(self.lib.fdfd_direct_stats(self.h, C.byref(fb), C.byref(ff)))
        return dict(factor_bytes=fb.value, factor_flops=ff.value)

    def solve(self, b, max_refine=3, tol=1e-12):
        if not self.factored:
            self.factor()
        b = as_c128(b)
        n = self.op.nx * self.op.ny
        nrhs = b.size // n
        x = pinned_empty(b.shape)
        rr, steps = C.c_double(0), C.c_int(0)
        check(self.lib.fdfd_direct_solve_host(self.h, self.op.h, ptr(b), ptr(x), nrhs, int(max_refine), float(tol),
                                              C.byref(rr), C.byref(steps)))
        self.last_relres, self.last_refine_steps = rr.value, steps.value
        return x

    def solve_fields(self, src, scale, averaging=None, max_refine=3, tol=1e-12):
        """x = A^-1 (scale * src) and the two in-plane fields in ONE library call
        (simulation.py:113-178): src crosses PCIe once (as float64 when it is real), x never comes
        back up for the derived fields, and the three results land in page-locked arrays."""
        src = np.asarray(src)
        shape = (self.op.nx, self.op.ny)
        if src.size != shape[0] * shape[1]:
            raise ValueError("src must have the grid's shape")
        real = np.isrealobj(src)
        s = np.ascontiguousarray(src, dtype=np.float64 if real else np.complex128)
        x, f1, f2 = pinned_empty(shape), pinned_empty(shape), pinned_empty(shape)
        rr, steps, fms = C.c_double(0), C.c_int(0), C.c_double(0)
        av = -1 if averaging is None else int(bool(averaging))
        scale = complex(scale)
        # a missing factorisation is done INSIDE the call (queued first; src crosses PCIe while it runs)
        self.last_factor_ms = None
        factoring = not self.factored
        try:
            check(self.lib.fdfd_factor_solve_fields_host(self.h, self.op.h, ptr(s), int(real), scale.real, scale.imag,
                                                         ptr(x), ptr(f1), ptr(f2), av, int(max_refine), float(tol),
                                                         C.byref(rr), C.byref(steps), C.byref(fms)))
        except Exception:
            if factoring:
                self.factored = self.has_factors = False
            raise
        if factoring:
            self.factored = self.has_factors = True
            self.last_factor_ms = fms.value
        self.last_relres, self.last_refine_steps = rr.value, steps.value
        return x, f1, f2


def mode_solve(eps_line, omega, dl, pol, L0, neff, order=1, averaged=False):
    """``order`` eigenpairs of the 1-D waveguide operator nearest (w sqrt(mu0' eps0') neff)^2,
    closest first (reference: source/mode.py:91-92 -> solver_eigs)."""
    _lib.require_gpu()
    lib = _lib.load()
    eps_line = np.ascontiguousarray(np.real(eps_line), dtype=np.float64).reshape(-1)
    n = eps_line.size
    vals = np.empty(order, dtype=np.float64)
    vecs = np.empty((order, n), dtype=np.float64)
    check(lib.fdfd_mode_solve_host(ptr(eps_line), n, float(omega), float(dl), POL[pol], float(L0), float(neff),
                                   int(order), int(bool(averaged)), ptr(vals), ptr(vecs)))
    return vals, vecs
