import sys, time, numpy as np
sys.path.insert(0, ".")
import bench
from fdfdpy_b200 import Simulation
n = 4096
eps = bench.synthetic_eps(n); src = bench.synthetic_src(n)
sim = Simulation(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
for it in range(3):
    t0 = time.perf_counter(); sim.eps_r = eps; t1 = time.perf_counter(); sim.src = src
    f = sim.solve_fields(); t2 = time.perf_counter()
    print("e2e iter %d: eps setter %.1f ms (factor incl.: no), solve_fields %.1f ms [factor %.1f], total %.1f ms relres %.2e" % (
        it, (t1-t0)*1e3, (t2-t1)*1e3, sim.timings.get('factor', 0)*1e3, (t2-t0)*1e3, sim.last_solve['relres']), flush=True)
