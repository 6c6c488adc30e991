"""Fused matrix-free stencils (Ez, and Hz with STENCIL_POL=Hz; STENCIL_LOSSY=1 makes eps complex): parity against the
planes kernel and GB/s for the row-march variants."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from fdfdpy_b200 import _lib, core  # noqa: E402

import os
lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [4096]
POL = os.environ.get("STENCIL_POL", "Ez")
for n in sizes:
    eps_in = bench.synthetic_eps(n)
    if os.environ.get("STENCIL_LOSSY"):
        eps_in = eps_in * (1 + 0.01j)
    op = core.MaxwellOperator(bench.OMEGA0, eps_in, bench.DL, bench.NPML, POL, bench.L0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    ref = op.dot(x, fused=False)
    d_x, d_y = C.c_void_p(), C.c_void_p()
    _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * n * n))
    _lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * n * n))
    _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(x), 16.0 * n * n))
    # Hz: (-4, -8) = one-shot kernel, 0 = marching kernel with automatic rows per CTA, > 0 = rows per CTA
    variants = [(2, 1), (4, 1), (8, 1)] if POL == "Ez" else [(-4, 1), (0, 1), (4, 1), (8, 1), (16, 1), (32, 1), (64, 1), (32, 0)]
    for rows, halo in variants:
        if POL == "Ez":
            _lib.check(lib.fdfd_stencil_set_variant(rows, 0))
        else:
            _lib.check(lib.fdfd_stencil_set_hz_variant(rows, halo))
        err = np.linalg.norm(op.dot(x, fused=True) - ref) / np.linalg.norm(ref)
        for _ in range(5):
            _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, 1))
        ms = C.c_double(0)
        reps = 50
        _lib.check(lib.fdfd_timer_start(op.h))
        for _ in range(reps):
            _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, 1))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
        print(f"n={n} {POL} rows={rows} halo={halo}: {ms.value / reps * 1e3:.1f} us  {48.0 * n * n * reps / ms.value / 1e6:.0f} GB/s  "
              f"rel diff vs planes kernel {err:.2e}", flush=True)
    _lib.check(lib.fdfd_stencil_set_variant(4, 0))
    _lib.check(lib.fdfd_stencil_set_hz_variant(0, 1))
    if POL != "Ez" or os.environ.get("STENCIL_ONLY"):
        lib.fdfd_free(d_x)
        lib.fdfd_free(d_y)
        del op
        continue
    # complex64 storage (24 B/cell) and one BiCGSTAB iteration in both storage types
    x32 = x.astype(np.complex64)
    err32 = np.linalg.norm(op.dot(x32, fused=True) - ref) / np.linalg.norm(ref)
    for rows in (2, 4, 8):
        _lib.check(lib.fdfd_stencil_set_variant(rows, 1))
        for _ in range(5):
            _lib.check(lib.fdfd_op_apply_dev_c64(op.h, d_x, d_y, 1))
        ms = C.c_double(0)
        _lib.check(lib.fdfd_timer_start(op.h))
        for _ in range(50):
            _lib.check(lib.fdfd_op_apply_dev_c64(op.h, d_x, d_y, 1))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
        print(f"n={n} complex64 fused rows={rows}: {ms.value / 50 * 1e3:.1f} us  {24.0 * n * n * 50 / ms.value / 1e6:.0f} GB/s  "
              f"rel diff vs fp64 planes kernel {err32:.2e}", flush=True)
    _lib.check(lib.fdfd_stencil_set_variant(4, 1))
    it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
    b = np.zeros((n, n), dtype=np.complex128)
    b[n // 2, n // 2] = 1j * bench.OMEGA0
    for name, fn, arr in (("complex128", lib.fdfd_krylov_solve_dev, b), ("complex64", lib.fdfd_krylov_solve_dev_c64, b.astype(np.complex64))):
        for rep in range(2):
            _lib.check(lib.fdfd_memcpy_h2d(d_y, _lib.ptr(arr), float(arr.nbytes)))
            _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(np.zeros_like(arr)), float(arr.nbytes)))
            _lib.check(lib.fdfd_timer_start(op.h))
            if name == "complex128":
                _lib.check(fn(op.h, None, d_y, d_x, 0, 1e-30, 200, 1, 200, None, 0, C.byref(it), C.byref(rr), C.byref(conv)))
            else:
                _lib.check(fn(op.h, d_y, d_x, 0, 1e-30, 200, 1, 200, C.byref(it), C.byref(rr), C.byref(conv)))
            _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
        print(f"n={n} BiCGSTAB {name}: {ms.value / max(it.value, 1):.3f} ms/iteration ({it.value} iterations, "
              f"relres after them {rr.value:.3e})", flush=True)
    lib.fdfd_free(d_x)
    lib.fdfd_free(d_y)
    del op
