"""Launches the kernels that get an `ncu --set full` capture once each, at the bench workload's sizes:
the persistent DMMA zgemm on a top-level Schur update, the fused Ez stencil and the planes stencil at 4096^2."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from fdfdpy_b200 import _lib, core  # noqa: E402

lib = _lib.load()
n = 4096
op = core.MaxwellOperator(bench.OMEGA0, bench.synthetic_eps(n), bench.DL, bench.NPML, "Ez", bench.L0)
d_x, d_y = C.c_void_p(), C.c_void_p()
_lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * n * n))
_lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * n * n))
x = np.ones((n, n), dtype=np.complex128)
_lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(x), 16.0 * n * n))
for fused in (1, 0):
    for _ in range(2):
        _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, fused))
_lib.check(lib.fdfd_op_sync(op.h))
ms = C.c_double(0)
# level-19 Schur update of the 4096^2 tree: S (8192 x 8192, lower) -= G (8192 x 4094) F_RE^T
_lib.check(lib.fdfd_zgemm_bench(8192, 8192, 4094, 1, 1, 1, 1, 1, C.byref(ms)))
print("zgemm ms", ms.value)
