"""Where the end-to-end milliseconds of Simulation.solve_fields go (host timers around each piece)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from fdfdpy_b200 import Simulation, _lib

n = 4096
eps = bench.synthetic_eps(n)
src = bench.synthetic_src(n)
sim = Simulation(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
lib = _lib.load()
for it in range(4):
    t0 = time.perf_counter(); sim.eps_r = eps; t1 = time.perf_counter(); sim.src = src
    f = sim.solve_fields(); t2 = time.perf_counter()
    print("iter %d: eps setter %.1f ms, solve_fields %.1f ms [factor %.1f], total %.1f ms relres %.2e" % (
        it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, sim.timings.get('factor', 0) * 1e3, (t2 - t0) * 1e3,
        sim.last_solve['relres']), flush=True)
# pieces of the last call, repeated in isolation on the same handles
d = sim._op.direct()
t = time.perf_counter(); a = src.any(); print("src.any(): %.1f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter(); bufs = [_lib.pinned_empty((n, n)) for _ in range(3)]; print("3 pinned_empty (pool miss): %.1f ms" % ((time.perf_counter() - t) * 1e3))
del bufs
t = time.perf_counter(); bufs = [_lib.pinned_empty((n, n)) for _ in range(3)]; print("3 pinned_empty (pool hit): %.1f ms" % ((time.perf_counter() - t) * 1e3))
for rep in range(2):
    t = time.perf_counter(); x, f1, f2 = d.solve_fields(src, 1j * bench.OMEGA0); print("solve_fields call (factor cached): %.1f ms" % ((time.perf_counter() - t) * 1e3))
dbuf = C.c_void_p(); _lib.check(lib.fdfd_malloc(C.byref(dbuf), 16.0 * n * n))
for name, arr in (("pageable f64 134MB", src), ("pinned c128 268MB", x)):
    t = time.perf_counter(); _lib.check(lib.fdfd_memcpy_h2d(dbuf, _lib.ptr(arr), float(arr.nbytes))); dt = time.perf_counter() - t
    print("H2D %s: %.1f ms (%.1f GB/s)" % (name, dt * 1e3, arr.nbytes / dt / 1e9))
    t = time.perf_counter(); _lib.check(lib.fdfd_memcpy_d2h(_lib.ptr(arr), dbuf, float(arr.nbytes))); dt = time.perf_counter() - t
    print("D2H %s: %.1f ms (%.1f GB/s)" % (name, dt * 1e3, arr.nbytes / dt / 1e9))
