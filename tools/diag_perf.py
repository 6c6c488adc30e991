"""Quick GPU timing of the direct solver at a few sizes (diagnostic, not the benchmark)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from fdfdpy_b200 import core  # noqa: E402

OMEGA = 2 * np.pi * 200e12
sizes = [int(a) for a in sys.argv[1:]] or [512, 1024]
import os
tile = int(os.environ.get("FDFD_TILE", "64"))
if os.environ.get("ZGEMM_VARIANT"):
    from fdfdpy_b200 import _lib as _l
    _l.check(_l.load().fdfd_zgemm_set_variant(int(os.environ["ZGEMM_VARIANT"])))
for n in sizes:
    rng = np.random.default_rng(0)
    eps = 1 + 11 * (rng.random((n, n)) > 0.5)
    t0 = time.time()
    op = core.MaxwellOperator(OMEGA, eps, 0.02, [15, 15], "Ez", 1e-6)
    t1 = time.time()
    d = core.DirectSolver(op, tile=tile)
    t2 = time.time()
    d.factor()
    t3 = time.time()
    import ctypes as C
    from fdfdpy_b200 import _lib
    op.lib.fdfd_phase_timing(1)
    d.factor()
    t4 = time.time()
    ph = np.zeros(13)
    op.lib.fdfd_phase_timing_read(_lib.ptr(ph))
    op.lib.fdfd_phase_timing(0)
    names = "assemble pivot panel rowgemm copy update expand solve_fwd solve_bwd stencil ggemm schur small".split()
    print("   phases ms:", " ".join(f"{k}={v:.1f}" for k, v in zip(names, ph)), "sum=%.1f" % ph.sum(), flush=True)
    b = np.zeros((n, n), dtype=complex)
    b[n // 2, n // 2] = 1j * OMEGA
    x = d.solve(b, max_refine=0)
    t5 = time.time()
    op.lib.fdfd_phase_timing(1)
    d.factor()
    x = d.solve(b, max_refine=-1)
    nl = len(d.levels)
    pl = np.zeros((nl, 13))
    op.lib.fdfd_phase_timing_read_levels(_lib.ptr(pl), nl)
    op.lib.fdfd_phase_timing(0)
    print("   per level ms (" + " ".join(names) + "):")
    for li, lv in enumerate(d.levels):
        if os.environ.get("DIAG_BRIEF") and li < len(d.levels) - 14:
            continue
        print(f"   L{li:02d} nb={lv.nb:8d} k={lv.kmax:5d} m={lv.mmax:5d} | " + " ".join(f"{v:7.2f}" for v in pl[li]) +
              f" | sum {pl[li].sum():8.2f}", flush=True)
    print("   totals: " + " ".join(f"{k}={v:.1f}" for k, v in zip(names, pl.sum(0))), flush=True)
    t5 = time.time()
    x = d.solve(b, max_refine=3, tol=1e-12)
    t6 = time.time()
    st = d.stats()
    print(f"N={n} op {t1-t0:.3f}s plan {t2-t1:.3f}s factor1 {t3-t2:.3f}s factor2 {t4-t3:.3f}s "
          f"({st['factor_flops']/(t4-t3)/1e12:.2f} TF/s, {st['factor_bytes']/1e9:.2f} GB) "
          f"solve0 {t5-t4:.3f}s solve+refine {t6-t5:.3f}s relres {d.last_relres:.2e} steps {d.last_refine_steps}",
          flush=True)
    del d, op
