python bench.py --steps 4 --warmup 2 --workload sweep 2>&1 | tail -1 | cut -c 1-330
python bench.py --steps 8 --warmup 4 --workload sweep 2>&1 | tail -1 | cut -c 1-330
python - <<'PY'
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, ".")
import bench
from fdfdpy_b200 import _lib, core
lib=_lib.load()
n=2048
eps=bench.synthetic_eps(n)
op=core.MaxwellOperator(bench.OMEGA0, eps, bench.DL, bench.NPML, "Hz", bench.L0)
d=core.DirectSolver(op)
b=np.zeros((16,n,n),dtype=complex); b[:,1000,1000]=1
for it in range(3):
    t0=time.perf_counter(); d.factor(); lib.fdfd_op_sync(op.h); t1=time.perf_counter()
    x=d.solve(b); t2=time.perf_counter()
    print("factor %.1f ms, 16-rhs solve_host %.1f ms relres %.1e steps %d"%((t1-t0)*1e3,(t2-t1)*1e3,d.last_relres,d.last_refine_steps))
PY
