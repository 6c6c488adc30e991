"""Schwarz-preconditioned slab solve with in-process ranks on ONE GPU: iteration counts and times.
    python tools/schwarz_probe.py N WORLD [overlap npml_sub]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from fdfdpy_b200.distributed import SlabOperator, run_ranks  # noqa: E402

n, world = int(sys.argv[1]), int(sys.argv[2])
ov, npml_s = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (4, 12)
kind = sys.argv[5] if len(sys.argv) > 5 else "device"
eps, src = (bench.synthetic_device_eps(n), bench.synthetic_device_src(n)) if kind == "device" else (bench.synthetic_eps(n), bench.synthetic_src(n))


def body(comm):
    slab = SlabOperator(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0, comm=comm if world > 1 else None)
    t0 = time.perf_counter()
    d = slab.setup_schwarz(eps, overlap=ov, npml_sub=npml_s)
    t1 = time.perf_counter()
    xs, info = slab.krylov(1j * bench.OMEGA0 * src[slab.x0:slab.x1], method="bicgstab", tol=1e-10, maxiter=1000, check_every=5)
    t2 = time.perf_counter()
    xg, infog = slab.krylov(1j * bench.OMEGA0 * src[slab.x0:slab.x1], method="gmres", tol=1e-10, maxiter=1000, restart=80)
    t3 = time.perf_counter()
    return dict(setup_s=t1 - t0, bicgstab_s=t2 - t1, gmres_s=t3 - t2, factor_gb=d.stats()["factor_bytes"] / 1e9, bicgstab=info,
                gmres=infog, diff=float(np.linalg.norm(xs - xg) / np.linalg.norm(xs)))


for r in run_ranks(world, body)[:1]:
    print(f"{kind} n={n} world={world} overlap={ov} npml_sub={npml_s}: {r}", flush=True)
