#!/usr/bin/env python
"""Benchmark of the fdfdpy_b200 hot path: full fp64 2-D FDFD solves at 4096 x 4096 (Ez).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 4096]

One "step" = one complete solve of the workload: assemble A from eps_r on the device, numeric
factorisation (batched nested dissection, complex GEMMs on the FP64 tensor pipe), substitution +
iterative refinement with the fp64 stencil residual, derived H fields.

* ``value``  : Mcell/s with eps_r and the source already resident in HBM (C ABI, device pointers).
* ``e2e``    : the same solve through the public API (``Simulation.eps_r = ...; solve_fields()``)
               with HOST numpy arrays in and out, copies inside the timed region.
* N > 1      : independent solves (one omega per GPU, "frequency sweep sharded one solve per GPU",
               no data-path collective) -> weak scaling; time = max over ranks.
* ``--impl reference``: the reference's CPU path (oracle port of its scipy/SuperLU branch) on a
               bounded sub-grid sample of the same workload, on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OMEGA0 = 2 * np.pi * 200e12
DL = 0.02            # cells of 20 nm at lambda0 = 1.5 um: 75 cells per vacuum wavelength
NPML = [15, 15]
L0 = 1e-6


def synthetic_eps(n, seed=0):
    """Random-permittivity photonic-crystal slab: square lattice (period 32 cells) of dielectric rods
    of radius 10 cells whose permittivity is drawn uniformly from [2, 12]; vacuum elsewhere."""
    rng = np.random.default_rng(seed)
    period, radius = 32, 10
    cells = (n + period - 1) // period
    rod_eps = 2 + 10 * rng.random((cells, cells))
    idx = np.arange(n)
    off = (idx % period) - period / 2 + 0.5
    inside = (off[:, None] ** 2 + off[None, :] ** 2) <= radius ** 2
    eps = np.where(inside, rod_eps[(idx // period)[:, None], (idx // period)[None, :]], 1.0)
    return np.ascontiguousarray(eps, dtype=np.float64)


def synthetic_src(n):
    src = np.zeros((n, n))
    src[n // 2, n // 2] = 1.0
    src[n // 3, (2 * n) // 3] = -0.5
    return src


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "power_w_max": float(np.max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (oracle port of linalg.py:139, scipy SuperLU)
# --------------------------------------------------------------------------------------------
def cpu_solve_sample(n_sample, full_n, reps=1):
    from oracle import fdfd_oracle as orc
    eps = synthetic_eps(full_n)[:n_sample, :n_sample]
    src = synthetic_src(n_sample)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        orc.solve_fields(OMEGA0, eps, DL, NPML, "Ez", L0, src)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n_s = args.cpu_sample or (256 if total > 6 else (384 if total > 2 else 512))
    for _ in range(args.warmup):
        cpu_solve_sample(n_s, args.size)
    times = [cpu_solve_sample(n_s, args.size) for _ in range(args.steps)]
    dt = float(np.mean(times))
    val = n_s * n_s / dt / 1e6
    cores = os.cpu_count()
    sample = ("{0}x{0} corner of the {1}x{1} workload, same eps/omega/PML, scipy SuperLU direct solve + derived "
              "fields (the reference's solver='scipy' branch; MKL Pardiso/pyMKL is not installable here)").format(
                  n_s, args.size)
    line = {"impl": "reference", "metric": "fdfd_solve_throughput", "value": val, "unit": "Mcell/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.size, sample=n_s),
            "cpu_baseline": {"value": val, "unit": "Mcell/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n, sample=None):
    cfg = {"workload": "Ez {0}x{0} synthetic random-permittivity photonic-crystal slab (BASELINE config 3 geometry at "
                       "the metric's 4096^2 size), two point sources, one omega per GPU".format(n),
           "grid": [n, n], "npml": NPML, "dl": DL, "omega": OMEGA0, "pol": "Ez",
           "solver": "batched nested-dissection direct solve + fp64 stencil refinement",
           "l2": "inputs larger than L2 (fronts and factors are tens of GB per step)"}
    if sample:
        cfg["cpu_sample_grid"] = [sample, sample]
    return cfg


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fdfdpy_b200 import _lib, core, Simulation
    lib = _lib.load()
    _lib.require_gpu()
    _lib.check(lib.fdfd_set_device(local))
    n = args.size
    ncell = n * n
    omega = OMEGA0 * (1 + 0.01 * rank)          # frequency sweep: one omega per GPU
    eps = synthetic_eps(n)
    src = synthetic_src(n)
    b_host = np.ascontiguousarray(src * 1j * omega, dtype=np.complex128)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- device-resident arm (value) ----------------
    op = core.MaxwellOperator(omega, eps, DL, NPML, "Ez", L0)
    direct = core.DirectSolver(op, tile=args.tile)
    nbytes = 16.0 * ncell
    d_eps, d_b, d_x, d_f = (C.c_void_p() for _ in range(4))
    _lib.check(lib.fdfd_malloc(C.byref(d_eps), nbytes))
    _lib.check(lib.fdfd_malloc(C.byref(d_b), nbytes))
    _lib.check(lib.fdfd_malloc(C.byref(d_x), nbytes))
    _lib.check(lib.fdfd_malloc(C.byref(d_f), 2 * nbytes))
    eps_c = _lib.as_c128(eps)
    _lib.check(lib.fdfd_memcpy_h2d(d_eps, _lib.ptr(eps_c), nbytes))
    _lib.check(lib.fdfd_memcpy_h2d(d_b, _lib.ptr(b_host), nbytes))
    relres, steps_ref = C.c_double(0), C.c_int(0)
    d_f2 = C.c_void_p(d_f.value + int(nbytes))

    def step_dev():
        _lib.check(lib.fdfd_op_assemble_dev(op.h, d_eps, None, 1))
        _lib.check(lib.fdfd_direct_factor(direct.h, op.h))
        _lib.check(lib.fdfd_direct_solve_dev(direct.h, op.h, d_b, d_x, 1, 3, 1e-12, C.byref(relres),
                                             C.byref(steps_ref)))
        _lib.check(lib.fdfd_op_derive_fields_dev(op.h, d_x, d_f, d_f2, -1))

    for _ in range(args.warmup):
        step_dev()
    _lib.check(lib.fdfd_op_sync(op.h))
    lib.fdfd_launch_count(1)
    _lib.check(lib.fdfd_gemm_timing(1))
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    _lib.check(lib.fdfd_timer_start(op.h))
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step_dev()
    ms = C.c_double(0)
    _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms)))
    wall = time.perf_counter() - t_wall
    barrier()
    clocks = sampler.stop()
    launches = lib.fdfd_launch_count(1)
    gt = np.zeros(6)
    _lib.check(lib.fdfd_gemm_timing_read(_lib.ptr(gt)))
    _lib.check(lib.fdfd_gemm_timing(0))
    stats = direct.stats()
    dev_ms = ms.value
    relres_dev = relres.value

    # ---------------- where the step goes: the four library calls timed one by one (outside the timed region)
    breakdown = {}
    calls = [
        ("assemble", lambda: lib.fdfd_op_assemble_dev(op.h, d_eps, None, 1)),
        ("factor", lambda: lib.fdfd_direct_factor(direct.h, op.h)),
        ("solve_refine", lambda: lib.fdfd_direct_solve_dev(direct.h, op.h, d_b, d_x, 1, 3, 1e-12, C.byref(relres),
                                                          C.byref(steps_ref))),
        ("derive_fields", lambda: lib.fdfd_op_derive_fields_dev(op.h, d_x, d_f, d_f2, -1)),
    ]
    for name, fn in calls:
        tms = C.c_double(0)
        _lib.check(lib.fdfd_timer_start(op.h))
        _lib.check(fn())
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(tms)))
        breakdown[name + "_ms"] = tms.value

    # ---------------- stencil kernel on its own (HBM roofline of the matrix-free path) ----------------
    reps = 20
    for _ in range(3):
        _lib.check(lib.fdfd_op_apply_dev(op.h, d_b, d_x, 1, 1))
    _lib.check(lib.fdfd_timer_start(op.h))
    for _ in range(reps):
        _lib.check(lib.fdfd_op_apply_dev(op.h, d_b, d_x, 1, 1))
    sms = C.c_double(0)
    _lib.check(lib.fdfd_timer_stop(op.h, C.byref(sms)))
    stencil_gbs = 48.0 * ncell * reps / (sms.value * 1e-3) / 1e9
    dmma = C.c_double(0)
    _lib.check(lib.fdfd_dmma_peak(C.byref(dmma)))

    for p in (d_eps, d_b, d_x, d_f):
        lib.fdfd_free(p)
    del direct, op

    # ---------------- end-to-end arm through the public API ----------------
    if args.no_e2e:
        print(json.dumps({"profiling_run": True, "ms_per_step": dev_ms / args.steps, "gpu_launches": int(launches),
                          "zgemm_big_ms": gt[0], "zgemm_big_tflops": gt[1] / (gt[0] * 1e-3) / 1e12 if gt[0] else 0,
                          "stencil_gbs": stencil_gbs}), flush=True)
        return
    sim = Simulation(omega, eps, DL, NPML, "Ez", L0)
    e2e_steps = max(1, min(args.steps, 2))

    def step_api():
        sim.eps_r = eps                 # H2D of eps, re-assembly, drops the factorisation
        sim.src = src
        hx, hy, ez = sim.solve_fields()  # factor + solve + D2H of three fields
        return ez

    for _ in range(max(1, min(args.warmup, 3))):
        ez = step_api()                 # warm-up: allocations, plan cache, page-locked result pool
    barrier()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        ez = step_api()
    e2e_s = (time.perf_counter() - t) / e2e_steps
    barrier()
    e2e_relres = sim.last_solve["relres"]
    e2e_timings = {k + "_ms": v * 1e3 for k, v in sim.timings.items()}
    del sim

    # ---------------- reduce over ranks ----------------
    dev_ms_max, e2e_max = dev_ms, e2e_s
    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_max = float(tt[0]), float(tt[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms_max / args.steps
    value = world * ncell / (ms_per_step * 1e-3) / 1e6
    e2e_val = world * ncell / e2e_max / 1e6
    peaks, peak_src = measured_peaks()
    big_ms, big_fl, big_n = gt[0], gt[1], gt[2]
    achieved = big_fl / (big_ms * 1e-3) / 1e12 if big_ms > 0 else 0.0
    # FP64 tensor-pipe ceiling: 64 DFMA/clk/SM (one m8n8k4 DMMA per SM sub-partition every 4 clocks) x 148 SMs at
    # the SM clock sampled under load; the register-resident DMMA probe is reported next to it (it is
    # power-limited when run alone: every SM issuing DMMA back to back pulls the clock below what the solver sees)
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    fp64_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    line = {
        "metric": "fdfd_solve_throughput", "value": value, "unit": "Mcell/s",
        "solves_per_s": world / (ms_per_step * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Mcell/s", "ms_per_step": e2e_max * 1e3,
                "h2d_bytes_per_step": int(eps.nbytes + src.nbytes), "d2h_bytes_per_step": int(3 * nbytes),
                "host_memory": "inputs: the caller's pageable float64 numpy arrays (eps_r setter, src), widened to "
                               "complex on the device; outputs: three complex128 fields in page-locked arrays "
                               "from the library's pool, through Simulation.solve_fields",
                "relres": e2e_relres, "last_step_host_timers": e2e_timings},
        "gpu_launches": int(launches),
        "relres": relres_dev, "refine_steps": steps_ref.value,
        "wall_ms_per_step": wall / args.steps * 1e3,
        "breakdown": breakdown,
        "roofline": {
            "kernel": "zgemm_dmma_kernel<4,2,3> (rank-T sweep update, complex128 on DMMA m8n8k4)",
            "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak if fp64_peak else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu --set full capture
            # (profiles/r01_full_capture_zgemm_persistent_and_planes_stencil.txt): the level-19 Schur update
            # S(8192x8192, lower) -= G(8192x4094) F_RE^T, 1.108e12 flop, 2.15e9 algorithmic bytes; DRAM runs at 8 %
            # of its peak there (tensor-bound kernel, 91.5 % tensor-pipe active), so the 10x re-read is not the limiter
            "traffic": 2.197e10,
            "traffic_note": "bytes of one captured launch (1.108e12 flop, 2.15e9 algorithmic bytes), not of the "
                            "per-step average launch; see profiles/",
            "peak_source": "FP64 tensor pipe: 148 SMs x 128 flop/clk x the SM clock sampled under load "
                           "(MEASURED_PEAKS.json has no FP64 entry; vendor-nominal is 40 TFLOP/s)",
            "dmma_probe_tflops": dmma.value,
            "launches_timed": int(big_n), "kernel_ms_per_step": big_ms / args.steps,
            "share_of_step": big_ms / dev_ms if dev_ms else None,
            "algorithmic_flops_per_step": big_fl / args.steps,
        },
        "roofline_small_gemm": {"kernel": "zgemm_dmma_kernel<2,1,2>", "ms_per_step": gt[3] / args.steps,
                                "tflops": gt[4] / (gt[3] * 1e-3) / 1e12 if gt[3] > 0 else 0.0,
                                "launches": int(gt[5])},
        "roofline_stencil": {"kernel": "stencil_fused_ez_kernel", "bound": "hbm", "achieved": stencil_gbs,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": stencil_gbs / peaks["hbm_gbs"],
                             "peak_source": peak_src, "algorithmic_bytes_per_cell": 48},
        "factor": {"bytes": stats["factor_bytes"], "flops": stats["factor_flops"]},
    }
    # bounded CPU baseline (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu_baseline:
        n_s = 512
        dt = cpu_solve_sample(n_s, n)
        line["cpu_baseline"] = {
            "value": n_s * n_s / dt / 1e6, "unit": "Mcell/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "{0}x{0} corner of the workload, one solve ({1:.1f} s): scipy SuperLU, the reference's "
                      "solver='scipy' branch (linalg.py:139); SuperLU factorisation is single-threaded".format(n_s, dt)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE config 4: broadband sweep of a 2048 x 2048 Hz device, one factorisation per omega reused by 16
    source right-hand sides, the omegas sharded over the GPUs (no collective).  One step = one omega per GPU."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fdfdpy_b200 import _lib, core
    lib = _lib.load()
    _lib.check(lib.fdfd_set_device(local))
    n, nrhs, nfreq = args.size if args.size != 4096 else 2048, 16, 64
    ncell = n * n
    eps = synthetic_eps(n)
    rng = np.random.default_rng(1)
    b_host = np.zeros((nrhs, n, n), dtype=np.complex128)
    for j in range(nrhs):                           # 16 point-dipole sources
        b_host[j, rng.integers(n // 4, 3 * n // 4), rng.integers(n // 4, 3 * n // 4)] = 1j * OMEGA0
    my_freqs = [OMEGA0 * (0.9 + 0.2 * k / (nfreq - 1)) for k in range(nfreq) if k % world == rank]
    nbytes = 16.0 * ncell
    d_eps, d_b, d_x = (C.c_void_p() for _ in range(3))
    _lib.check(lib.fdfd_malloc(C.byref(d_eps), nbytes))
    _lib.check(lib.fdfd_malloc(C.byref(d_b), nbytes * nrhs))
    _lib.check(lib.fdfd_malloc(C.byref(d_x), nbytes * nrhs))
    _lib.check(lib.fdfd_memcpy_h2d(d_eps, _lib.ptr(_lib.as_c128(eps)), nbytes))
    _lib.check(lib.fdfd_memcpy_h2d(d_b, _lib.ptr(b_host), nbytes * nrhs))
    relres, steps_ref = C.c_double(0), C.c_int(0)
    ops = {}

    def step(k):
        omega = my_freqs[k % len(my_freqs)]
        if omega not in ops:                        # operator handles are per omega (PML depends on it); plan is shared
            if len(ops) >= 2:
                ops.pop(next(iter(ops)))
            op = core.MaxwellOperator(omega, eps, DL, NPML, "Hz", L0)
            ops[omega] = (op, core.DirectSolver(op))
        op, direct = ops[omega]
        _lib.check(lib.fdfd_op_assemble_dev(op.h, d_eps, None, 1))
        _lib.check(lib.fdfd_direct_factor(direct.h, op.h))
        _lib.check(lib.fdfd_direct_solve_dev(direct.h, op.h, d_b, d_x, nrhs, 3, 1e-12, C.byref(relres), C.byref(steps_ref)))
        _lib.check(lib.fdfd_op_sync(op.h))
        return relres.value

    for k in range(args.warmup):
        step(k % 2)
    if dist is not None:
        dist.barrier()
    t = time.perf_counter()
    worst = 0.0
    for k in range(args.steps):
        worst = max(worst, step(k % 2))
    dt = time.perf_counter() - t
    if dist is not None:
        import torch
        tt = torch.tensor([dt, worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, worst = float(tt[0]), float(tt[1])
    if rank == 0:
        per = dt / args.steps
        print(json.dumps({
            "metric": "fdfd_sweep_throughput", "value": world * nrhs * ncell / per / 1e6, "unit": "Mcell/s",
            "rhs_solves_per_s": world * nrhs / per, "factorisations_per_s": world / per, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "dtype": "f64", "data": "synthetic", "relres_max": worst,
            "config": {"workload": "Hz {0}x{0} broadband sweep, 1 factorisation + {1} RHS per omega, omegas sharded over "
                                   "GPUs (BASELINE config 4)".format(n, nrhs), "grid": [n, n], "nrhs": nrhs,
                       "timing": "host wall clock around device-synchronised steps (operator re-created per omega)"}}),
              flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--tile", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the public-API arm")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="reference arm: side of the sub-grid sample (default: chosen from steps + warmup)")
    ap.add_argument("--workload", default="solve", choices=["solve", "sweep"],
                    help="solve: the headline 4096^2 Ez solve (default); sweep: BASELINE config 4")
    args = ap.parse_args()
    if args.workload == "sweep" and args.impl == "ours":
        return run_sweep(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
