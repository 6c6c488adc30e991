"""Phase timing of a 16-right-hand-side solve at 2048^2 (BASELINE config 4's inner loop)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from fdfdpy_b200 import _lib, core

lib = _lib.load()
nrhs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2048, 2048)
pol = sys.argv[4] if len(sys.argv) > 4 else "Hz"
n = max(nx, ny)
op = core.MaxwellOperator(bench.OMEGA0, bench.synthetic_eps(n)[:nx, :ny], bench.DL, bench.NPML, pol, bench.L0)
d = core.DirectSolver(op)
d.factor()
rng = np.random.default_rng(0)
b = np.zeros((nrhs, nx, ny), dtype=np.complex128)
for j in range(nrhs):
    b[j, rng.integers(nx // 4, 3 * nx // 4), rng.integers(ny // 4, 3 * ny // 4)] = 1j * bench.OMEGA0
nbytes = 16.0 * nx * ny * nrhs
d_b, d_x = C.c_void_p(), C.c_void_p()
_lib.check(lib.fdfd_malloc(C.byref(d_b), nbytes))
_lib.check(lib.fdfd_malloc(C.byref(d_x), nbytes))
_lib.check(lib.fdfd_memcpy_h2d(d_b, _lib.ptr(b), nbytes))
rr, st = C.c_double(0), C.c_int(0)
names = "assemble pivot panel rowgemm copy update expand solve_fwd solve_bwd stencil ggemm schur small".split()
for refine in (-1, 3):
    for rep in range(2):
        lib.fdfd_phase_timing(1)
        ms = C.c_double(0)
        _lib.check(lib.fdfd_timer_start(op.h))
        _lib.check(lib.fdfd_direct_solve_dev(d.h, op.h, d_b, d_x, nrhs, refine, 1e-12, C.byref(rr), C.byref(st)))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
        ph = np.zeros(13)
        lib.fdfd_phase_timing_read(_lib.ptr(ph))
        lib.fdfd_phase_timing(0)
    print(f"{nx}x{ny} {pol} factor {d.stats()['factor_bytes'] / 1e9:.2f} GB nrhs={nrhs} max_refine={refine}: {ms.value:.1f} ms  relres {rr.value:.1e} steps {st.value}  phases: " +
          " ".join(f"{k}={v:.1f}" for k, v in zip(names, ph) if v > 0), flush=True)
nl = len(d.levels)
lib.fdfd_phase_timing(1)
_lib.check(lib.fdfd_direct_solve_dev(d.h, op.h, d_b, d_x, nrhs, -1, 1e-12, C.byref(rr), C.byref(st)))
pl = np.zeros((nl, 13))
lib.fdfd_phase_timing_read_levels(_lib.ptr(pl), nl)
lib.fdfd_phase_timing(0)
for li, lv in enumerate(d.levels):
    print(f"   L{li:02d} nb={lv.nb:8d} k={lv.kmax:5d} m={lv.mmax:5d} fwd {pl[li][7]:7.2f} bwd {pl[li][8]:7.2f}")
