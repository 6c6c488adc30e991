"""Launches the kernels that get an `ncu --set full` capture once each, at the bench workload's sizes:
the persistent DMMA zgemm on a chain-step Schur update of the 4096^2 tree, the fused Ez / Hz stencils and the
planes stencil at 4096^2.  PROFILE_ONLY=zgemm|stencil restricts the run."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from fdfdpy_b200 import _lib, core  # noqa: E402

lib = _lib.load()
only = os.environ.get("PROFILE_ONLY", "")
n = 4096
if only in ("", "stencil"):
    for pol in ("Ez", "Hz"):
        op = core.MaxwellOperator(bench.OMEGA0, bench.synthetic_eps(n), bench.DL, bench.NPML, pol, bench.L0)
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * n * n))
        _lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * n * n))
        x = np.ones((n, n), dtype=np.complex128)
        _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(x), 16.0 * n * n))
        for fused in (1, 0):
            for _ in range(2):
                _lib.check(lib.fdfd_op_apply_dev(op.h, d_x, d_y, 1, fused))
        _lib.check(lib.fdfd_op_sync(op.h))
        lib.fdfd_free(d_x)
        lib.fdfd_free(d_y)
        del op
if only in ("", "zgemm"):
    ms = C.c_double(0)
    # a chain step of the top fronts of the 4096^2 tree: S (9727 x 9727, lower) -= G (9727 x 512) F_RE^T
    _lib.check(lib.fdfd_zgemm_bench(9727, 9727, 512, 1, 1, 1, 1, 2, C.byref(ms)))
    print("zgemm chain step ms", ms.value, "TFLOP/s (8 per cMAC)", 8.0 * 9727 * (9727 + 64) / 2 * 512 / ms.value / 1e9)
    # the same shape, full square (what cuBLAS would have to do)
    _lib.check(lib.fdfd_zgemm_bench(8192, 8192, 512, 1, 1, 1, 0, 2, C.byref(ms)))
    print("zgemm 8192x8192x512 ms", ms.value, "TFLOP/s (8 per cMAC)", 8.0 * 8192 * 8192 * 512 / ms.value / 1e9)
