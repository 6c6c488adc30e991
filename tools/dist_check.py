"""Multi-GPU check of the sharded paths, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--parity 200x160] [--size 4096] [--reps 2]

* parity: sharded direct solve vs the CPU oracle on a small grid (rank 0 checks, all ranks must agree);
* size  : factor + solve timing of ONE size x size grid split over the ranks (device time, max over ranks).
torch.distributed (gloo) is only the rendezvous that carries the 128-byte NCCL id and the timing reduction.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parity", default="200x160")
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--pol", default="Ez")
    ap.add_argument("--slab-parity", default="")
    ap.add_argument("--slab-size", type=int, default=0)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--phases", action="store_true", help="one extra factorisation with per-level phase timing, every rank prints its top levels")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    from fdfdpy_b200 import _lib, core
    from fdfdpy_b200.distributed import Communicator
    lib = _lib.load()
    _lib.check(lib.fdfd_set_device(local))
    comm = Communicator.from_torch()
    omega = 2 * np.pi * 200e12
    out = {"world": world}

    if args.parity:
        nx, ny = (int(v) for v in args.parity.split("x"))
        rng = np.random.default_rng(1)
        eps = 1 + 5 * (rng.random((nx, ny)) > 0.5)
        b = rng.standard_normal((2, nx, ny)) + 1j * rng.standard_normal((2, nx, ny))
        op = core.MaxwellOperator(omega, eps, 0.04, [10, 8], args.pol, 1e-6)
        d = core.DirectSolver(op, comm=comm)
        x = d.solve(b).reshape(2, nx, ny)
        res = [float(np.linalg.norm(b[j].ravel() - op.dot(x[j]).ravel()) / np.linalg.norm(b[j])) for j in range(2)]
        out["parity_relres"] = max(res)
        xs = [None] * world
        dist.all_gather_object(xs, x)
        out["ranks_agree"] = bool(all(np.array_equal(xs[0], xi) for xi in xs))
        if rank == 0:
            from oracle import fdfd_oracle as orc
            A = orc.construct_A(omega, eps, 0.04, [10, 8], args.pol, 1e-6)
            ref = np.stack([orc.sparse_solve(A, b[j]).reshape(nx, ny) for j in range(2)])
            out["parity_rel_l2_vs_oracle"] = float(np.linalg.norm(x - ref) / np.linalg.norm(ref))
        del d, op

    if args.size:
        import bench
        n = args.size
        eps = bench.synthetic_eps(n)
        src = bench.synthetic_src(n)
        op = core.MaxwellOperator(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0)
        d = core.DirectSolver(op, comm=comm)
        import ctypes as C
        times = []
        for rep in range(args.reps + 1):
            dist.barrier()
            ms_f, ms_s = C.c_double(0), C.c_double(0)
            _lib.check(lib.fdfd_timer_start(op.h))
            d.factor()
            _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms_f)))
            t0 = time.perf_counter()
            x, f1, f2 = d.solve_fields(src, 1j * bench.OMEGA0)
            ms_s.value = (time.perf_counter() - t0) * 1e3
            times.append((ms_f.value, ms_s.value, d.last_relres))
        tt = torch.tensor([times[-1][0], times[-1][1]], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        st = d.stats()
        out.update({"size": n, "factor_ms_max": float(tt[0]), "solve_fields_host_ms_max": float(tt[1]),
                    "relres": times[-1][2], "factor_bytes_this_rank": st["factor_bytes"],
                    "factor_ms_all_reps": [t[0] for t in times]})
        fb, tb = C.c_double(0), C.c_double(0)
        lib.fdfd_mem_info(C.byref(fb), C.byref(tb))
        out["hbm_used_gb_this_rank"] = (tb.value - fb.value) / 1e9
        if args.phases:
            names = "assemble pivot panel rowgemm copy update expand solve_fwd solve_bwd stencil ggemm schur small".split()
            nl = len(d.levels)
            dist.barrier()
            lib.fdfd_phase_timing(1)
            d.factor()
            pl = np.zeros((nl, 13))
            lib.fdfd_phase_timing_read_levels(_lib.ptr(pl), nl)
            lib.fdfd_phase_timing(0)
            keep = [i for i in (0, 1, 4, 5, 6, 10, 11, 12)]
            first_top = next((i for i, lv in enumerate(d.levels) if lv.nb <= 1 and lv.kmax >= 256), nl)
            lines = ["rank %d: local levels (nb > 1 or small) sum %.1f ms; phases " % (rank, pl[:first_top].sum()) +
                     " ".join("%s=%.1f" % (names[i], pl[:, i].sum()) for i in keep)]
            for li in range(first_top, nl):
                lv = d.levels[li]
                lines.append("   r%d L%02d nb=%d k=%d m=%d | " % (rank, li, lv.nb, lv.kmax, lv.mmax) +
                             " ".join("%s=%.2f" % (names[i], pl[li, i]) for i in keep) + " | sum %.2f" % pl[li].sum())
            allp = [None] * world
            dist.all_gather_object(allp, "\n".join(lines))
            if rank == 0:
                for r in (0, world // 2, world - 1):
                    print(allp[r], file=sys.stderr, flush=True)
    if args.slab_parity:
        from fdfdpy_b200.distributed import SlabOperator
        from oracle import fdfd_oracle as orc
        nx, ny = (int(v) for v in args.slab_parity.split("x"))
        rng = np.random.default_rng(7)
        eps = 1 + 2 * rng.random((nx, ny))
        npml = [8, 8]
        slab = SlabOperator(omega, eps, 0.05, npml, args.pol, 1e-6, comm=comm)
        sl = slice(slab.x0, slab.x1)
        xv = rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny))
        A = orc.construct_A(omega, eps, 0.05, npml, args.pol, 1e-6)
        ref = A.dot(xv.ravel()).reshape(nx, ny)[sl]
        y = slab.dot(xv[sl])
        out["slab_apply_rel_err"] = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
        if True:
            yf = slab.dot(xv[sl], fused=True)
            out["slab_apply_fused_rel_err"] = float(np.linalg.norm(yf - ref) / np.linalg.norm(ref))
        b = np.zeros((nx, ny), dtype=complex)
        b[nx // 2, ny // 2] = 1j * omega
        sol = orc.sparse_solve(A, b).reshape(nx, ny)[sl]
        for method in ("bicgstab", "cocg"):
            xs, info = slab.krylov(b[sl], method=method, tol=1e-12, maxiter=20000, check_every=20)
            nrm = torch.tensor([np.linalg.norm(xs - sol) ** 2, np.linalg.norm(sol) ** 2], dtype=torch.float64)
            dist.all_reduce(nrm)
            out["slab_" + method] = dict(info, rel_l2_vs_oracle=float(np.sqrt(nrm[0] / nrm[1])))
        del slab

    if args.slab_size:
        import ctypes as C
        import bench
        from fdfdpy_b200.distributed import SlabOperator
        n = args.slab_size
        eps = bench.synthetic_eps(n)
        slab = SlabOperator(bench.OMEGA0, eps, bench.DL, bench.NPML, "Ez", bench.L0, comm=comm)
        nloc = (slab.nxl + 2) * n
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * nloc))
        _lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * nloc))
        xe = np.zeros((slab.nxl + 2, n), dtype=np.complex128)
        xe[1:-1] = 1.0
        _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(xe), 16.0 * nloc))
        _lib.check(lib.fdfd_memcpy_h2d(d_y, _lib.ptr(xe), 16.0 * nloc))
        res = {}
        for fused in (1, 0):
            for _ in range(5):
                _lib.check(lib.fdfd_op_apply_dev(slab.h, d_x, d_y, 1, fused))
            dist.barrier()
            ms = C.c_double(0)
            _lib.check(lib.fdfd_timer_start(slab.h))
            for _ in range(args.iters):
                _lib.check(lib.fdfd_op_apply_dev(slab.h, d_x, d_y, 1, fused))
            _lib.check(lib.fdfd_timer_stop(slab.h, C.byref(ms)))
            tt = torch.tensor([ms.value], dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            per = float(tt[0]) / args.iters
            bytes_cell = 48 if fused else 112
            res["fused" if fused else "planes"] = {"ms_per_apply": per, "agg_GBps": bytes_cell * n * n / (per * 1e-3) / 1e9,
                                                   "Gcell_per_s": n * n / (per * 1e-3) / 1e9}
        # a fixed number of BiCGSTAB iterations (2 stencils + 5 all-reduced inner products each)
        it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
        be = np.zeros((slab.nxl + 2, n), dtype=np.complex128)
        if slab.x0 <= n // 2 < slab.x1:
            be[1 + n // 2 - slab.x0, n // 2] = 1j * bench.OMEGA0
        _lib.check(lib.fdfd_memcpy_h2d(d_y, _lib.ptr(be), 16.0 * nloc))
        _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(np.zeros_like(be)), 16.0 * nloc))
        for warm in (1, 0):
            dist.barrier()
            ms = C.c_double(0)
            _lib.check(lib.fdfd_timer_start(slab.h))
            _lib.check(lib.fdfd_krylov_solve_dev(slab.h, None, d_y, d_x, 0, 1e-30, args.iters, 1, args.iters, None, 0,
                                                 C.byref(it), C.byref(rr), C.byref(conv)))
            _lib.check(lib.fdfd_timer_stop(slab.h, C.byref(ms)))
        tt = torch.tensor([ms.value], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["bicgstab"] = {"iters": it.value, "ms_per_iter": float(tt[0]) / max(it.value, 1), "relres_after": rr.value}
        out["slab_size"] = n
        out["slab_throughput"] = res
        lib.fdfd_free(d_x)
        lib.fdfd_free(d_y)
        del slab

    per_rank = [None] * world
    dist.all_gather_object(per_rank, out)
    if rank == 0:
        print(json.dumps({"rank0": out, "others": per_rank[1:]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
