"""Fused matrix-free Ez stencil: parity against the planes kernel and GB/s for the row-march variants."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from fdfdpy_b200 import _lib, core  # noqa: E402

lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [4096]
for n in sizes:
    op = core.MaxwellOperator(bench.OMEGA0, bench.synthetic_eps(n), bench.DL, bench.NPML, "Ez", bench.L0)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    ref = op.dot(x, fused=False)
    d_x, d_y = C.c_void_p(), C.c_void_p()
    _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * n * n))
    _lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * n * n))
    _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(x), 16.0 * n * n))
    for rows in (2, 4, 8):
        _lib.check(lib.fdfd_stencil_set_variant(rows))
        err = np.linalg.norm(op.dot(x, fused=True) - ref) / np.linalg.norm(ref)
        for _ in range(5):
            _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, 1))
        ms = C.c_double(0)
        reps = 50
        _lib.check(lib.fdfd_timer_start(op.h))
        for _ in range(reps):
            _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, 1))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
        print(f"n={n} rows={rows}: {ms.value / reps * 1e3:.1f} us  {48.0 * n * n * reps / ms.value / 1e6:.0f} GB/s  "
              f"rel diff vs planes kernel {err:.2e}", flush=True)
    _lib.check(lib.fdfd_stencil_set_variant(8))
    lib.fdfd_free(d_x)
    lib.fdfd_free(d_y)
    del op
