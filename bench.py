#!/usr/bin/env python
"""Benchmark of the fdfdpy_b200 hot path: full fp64 2-D FDFD solves at 4096 x 4096 (Ez).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 4096]

One "step" = one complete solve of the workload: assemble A from eps_r on the device, numeric
factorisation (batched nested dissection, complex GEMMs on the FP64 tensor pipe), substitution +
iterative refinement with the fp64 stencil residual, derived H fields.

* ``value``  : Mcell/s with eps_r and the source already resident in HBM (C ABI, device pointers).
* ``e2e``    : the same solve through the public API (``Simulation.eps_r = ...; solve_fields()``)
               with HOST numpy arrays in and out, copies inside the timed region, over all ``--steps``.
* N > 1      : independent solves (one omega per GPU, "frequency sweep sharded one solve per GPU",
               no data-path collective) -> weak scaling; time = max over ranks.  The SAME line then
               also carries the two one-grid-on-N-GPUs paths, measured right after the replicas:
               ``sharded`` (4096^2 direct solve, elimination tree split over the ranks, NCCL inside
               the library) and ``slab`` (8192^2 matrix-free stencil on slabs with halo exchange).
* ``--impl reference``: the reference's CPU path (oracle port of its scipy/SuperLU branch) on a
               FIXED 1024^2 corner of the same workload (2 repetitions, independent of --steps),
               preceded by a 256/512 size sweep so the extrapolation to 4096^2 is on the record.
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
CPU_REF_GRID = 1024  # the reference arm's fixed sample (largest that finishes in about a minute per solve)
CPU_REF_REPS = 2


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


def synthetic_device_eps(n):
    """A waveguide device of the kind the reference's notebooks and tests simulate (Examples.ipynb, test_flux.py): a
    straight high-index ridge (eps 12, 0.4 um wide) in oxide (eps 2.1) running along x through the whole grid, with
    a side-coupled rectangular resonator.  Weakly scattering compared with the random-rod crystal of the headline
    workload; used for the Schwarz-preconditioned slab solve."""
    e = np.full((n, n), 2.1)
    c = n // 2
    e[:, c - 10:c + 10] = 12.0
    e[n // 2 - n // 8:n // 2 + n // 8, c + 14:c + 34] = 12.0
    return e


def synthetic_device_src(n):
    src = np.zeros((n, n))
    src[n // 5, n // 2] = 1.0          # a dipole inside the guide
    return src


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
        """``gpu_index``: one index, a comma-separated list (one sampler process watches all of a job's GPUs: N
        nvidia-smi loops next to N launching processes cost host time and driver locks inside the timed region),
        or None for a sampler that does nothing (ranks other than 0)."""
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if self.gpu is None:
            return
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
        if self.gpu is None:
            return {}
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


def captured_traffic():
    """DRAM bytes of the dominant GEMM launch from the committed `ncu --set full` capture; the file is written by
    tools/ncu_traffic.py (tools/run_profiles.sh), never typed in by hand."""
    path = os.path.join(ROOT, "profiles", "zgemm_capture.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


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
    """The reference arm: a FIXED sample and repetition count, whatever --steps / --warmup say (a sparse LU's
    Mcell/s falls with the grid size, so the sample must not shrink when the driver asks for more steps)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = args.cpu_sample or CPU_REF_GRID
    sweep = []
    for n in [v for v in (256, 512) if v < n_s]:            # doubles as the warm-up (imports, scipy, page cache)
        dt = cpu_solve_sample(n, args.size)
        sweep.append({"grid": [n, n], "s_per_solve": dt, "Mcell_per_s": n * n / dt / 1e6})
    times = []
    for rep in range(CPU_REF_REPS):
        times.append(cpu_solve_sample(n_s, args.size))
        if times[-1] > 150.0:                                # a slow host: one repetition has to do
            break
    dt = float(np.mean(times))
    sweep.append({"grid": [n_s, n_s], "s_per_solve": dt, "Mcell_per_s": n_s * n_s / dt / 1e6})
    val = n_s * n_s / dt / 1e6
    # power-law extrapolation of the solve time to the full workload from the last two sweep points
    extrap = None
    if len(sweep) >= 2:
        (a, b) = sweep[-2], sweep[-1]
        p = np.log(b["s_per_solve"] / a["s_per_solve"]) / np.log(b["grid"][0] ** 2 / a["grid"][0] ** 2)
        t_full = b["s_per_solve"] * (args.size ** 2 / b["grid"][0] ** 2) ** p
        extrap = {"exponent_time_vs_cells": float(p), "s_per_solve_at_full_size": float(t_full),
                  "Mcell_per_s_at_full_size": args.size ** 2 / t_full / 1e6,
                  "note": "extrapolated, not measured: the {0}x{0} solve does not fit the time budget on a CPU".format(args.size)}
    cores = os.cpu_count()
    sample = ("{0}x{0} corner of the {1}x{1} workload, same eps/omega/PML, {2} solve(s) of {3:.1f} s: scipy SuperLU direct "
              "solve + derived fields (the reference's solver='scipy' branch, linalg.py:139; MKL Pardiso/pyMKL is not "
              "installable here); SuperLU's factorisation is single-threaded").format(n_s, args.size, len(times), dt)
    line = {"impl": "reference", "metric": "fdfd_solve_throughput", "value": val, "unit": "Mcell/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.size, sample=n_s),
            "cpu_reps": len(times), "size_sweep": sweep, "extrapolation": extrap,
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
        cfg["cpu_sample_note"] = ("the CPU arm solves a corner of the workload, not the workload: a size-matched ratio "
                                  "needs the extrapolation printed in `extrapolation`")
    return cfg


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def stencil_rate(lib, _lib, op, d_in, d_out, ncell, reps=20):
    for _ in range(3):
        _lib.check(lib.fdfd_op_apply_dev(op.h, d_in, d_out, 1, 1))
    _lib.check(lib.fdfd_timer_start(op.h))
    for _ in range(reps):
        _lib.check(lib.fdfd_op_apply_dev(op.h, d_in, d_out, 1, 1))
    sms = C.c_double(0)
    _lib.check(lib.fdfd_timer_stop(op.h, C.byref(sms)))
    return 48.0 * ncell * reps / (sms.value * 1e-3) / 1e9


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
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if rank == 0 else None)   # rank 0 watches every GPU
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
    exec_fl = C.c_double(0)
    _lib.check(lib.fdfd_gemm_timing_exec_flops(C.byref(exec_fl)))
    _lib.check(lib.fdfd_gemm_timing(0))
    stats = direct.stats()
    dev_ms = ms.value
    relres_dev = relres.value

    if args.only_step:                    # ncu launch-list runs: nothing after the timed step
        print(json.dumps({"profiling_run": True, "ms_per_step": dev_ms / args.steps, "gpu_launches": int(launches)}), flush=True)
        return
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

    # ---------------- the matrix-free stencils on their own (HBM roofline of path (a)), Ez and Hz ----------------
    stencil_gbs = stencil_rate(lib, _lib, op, d_b, d_x, ncell)
    op_hz = core.MaxwellOperator(omega, eps, DL, NPML, "Hz", L0)
    stencil_hz_gbs = stencil_rate(lib, _lib, op_hz, d_b, d_x, ncell)
    del op_hz
    # FP64 tensor-pipe probe: register-resident DMMA loop, with the SM clock it ran at measured INSIDE the kernel
    probe = np.zeros(4)
    _lib.check(lib.fdfd_dmma_probe_clocked(8, 8, _lib.ptr(probe)))

    for p in (d_eps, d_b, d_x, d_f):
        lib.fdfd_free(p)
    del direct, op

    # ---------------- end-to-end arm through the public API ----------------
    if args.no_e2e:
        print(json.dumps({"profiling_run": True, "ms_per_step": dev_ms / args.steps, "gpu_launches": int(launches),
                          "zgemm_big_ms": gt[0], "zgemm_big_tflops": gt[1] / (gt[0] * 1e-3) / 1e12 if gt[0] else 0,
                          "stencil_gbs": stencil_gbs, "stencil_hz_gbs": stencil_hz_gbs}), flush=True)
        return
    sim = Simulation(omega, eps, DL, NPML, "Ez", L0)
    e2e_steps = max(1, args.steps)

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
    del sim, ez

    # ---------------- one grid on N GPUs (N > 1): sharded direct solve and slab stencil ----------------
    multi = {}
    if world > 1 and not args.no_multi:
        try:
            multi = run_one_grid_multi_gpu(args, dist, rank, world, local, lib, _lib, core, breakdown)
        except Exception as e:            # a failure here must not take the replica numbers down with it
            multi = {"sharded": {"error": repr(e)[:300]}}

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
    big_ms, big_fl, big_n = gt[0] + gt[3], gt[1] + gt[4], gt[2] + gt[5]
    # ALGORITHMIC flops of a complex GEMM = 8 per complex multiply-add (4 real mul + 4 real add: the count LAPACK, cuBLAS
    # ZGEMM and every FP64 peak figure use).  `achieved` is that count over the measured kernel time.  The kernels
    # EXECUTE only 6 tensor flops per complex MAC (3M / Karatsuba form, zgemm.cuh); the executed rate is what occupies
    # the pipe and is reported next to it (`achieved_executed`, `frac_executed` = tensor-pipe utilisation).
    achieved_exec = exec_fl.value / (big_ms * 1e-3) / 1e12 if big_ms > 0 else 0.0
    achieved = big_fl / (big_ms * 1e-3) / 1e12 if big_ms > 0 else 0.0
    # FP64 tensor-pipe ceiling: 128 flop/clk/SM (one m8n8k4 DMMA per SM sub-partition every 16 clocks) x 148 SMs at
    # the SM clock sampled under load.  MEASURED_PEAKS.json has no FP64 entry, so the denominator is this pipe rate;
    # the register-resident probe next to it shows how much of it an ideal instruction stream reaches on this board.
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    fp64_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    cap = captured_traffic()
    line = {
        "metric": "fdfd_solve_throughput", "value": value, "unit": "Mcell/s",
        "solves_per_s": world / (ms_per_step * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Mcell/s", "ms_per_step": e2e_max * 1e3, "steps": e2e_steps,
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
            "kernel": "zgemm_dmma_persistent_kernel (+ the tiled zgemm_dmma_kernel on the many-small-front levels): "
                      "complex128 GEMM on DMMA m8n8k4, every launch of the step timed with CUDA events",
            "bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak if fp64_peak else None,
            "traffic": cap.get("dram_bytes") if cap else None,
            "traffic_capture": cap,
            "peak_source": "FP64 tensor pipe: 148 SMs x 128 flop/clk x the SM clock sampled under load "
                           "(MEASURED_PEAKS.json has no FP64 entry; vendor-nominal is 40 TFLOP/s)",
            "dmma_probe": {"tflops": probe[0], "sm_mhz_in_kernel": probe[1],
                           "pipe_rate_at_that_clock": 148 * 128 * probe[1] * 1e6 / 1e12 if probe[1] else None,
                           "what": "register-resident mma.sync.m8n8k4.f64 loop, 8 warps/SM x 8 accumulator chains; "
                                   "clock = clock64 / globaltimer measured inside the same kernel"},
            "launches_timed": int(big_n), "kernel_ms_per_step": big_ms / args.steps,
            "share_of_step": big_ms / dev_ms if dev_ms else None,
            "algorithmic_flops_per_step": big_fl / args.steps,
            "algorithmic_flops_per_complex_mac": 8,
            "executed_flops_per_step": exec_fl.value / args.steps, "executed_flops_per_complex_mac": 6,
            "achieved_executed": achieved_exec,
            "frac_executed": achieved_exec / fp64_peak if fp64_peak else None,
            "note": "achieved = algorithmic flops (8 per complex multiply-add, the ZGEMM count; cuBLAS ZGEMM 8192^3 "
                    "reaches 36.9 TFLOP/s on this board, tools/zgemm_vs_cublas.py) / kernel time; the 3M kernels "
                    "execute 6 per complex MAC, so frac_executed (= tensor-pipe utilisation) is 3/4 of frac",
        },
        "roofline_stencil": {"kernel": "stencil_fused_ez_kernel", "bound": "hbm", "achieved": stencil_gbs,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": stencil_gbs / peaks["hbm_gbs"],
                             "peak_source": peak_src, "algorithmic_bytes_per_cell": 48},
        "roofline_stencil_hz": {"kernel": "stencil_fused_hz_kernel", "bound": "hbm", "achieved": stencil_hz_gbs,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": stencil_hz_gbs / peaks["hbm_gbs"],
                                "peak_source": peak_src, "algorithmic_bytes_per_cell": 48},
        "factor": {"bytes": stats["factor_bytes"], "flops": stats["factor_flops"]},
    }
    line.update(multi)
    # bounded CPU baseline (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu_baseline:
        n_s = 512
        dt = cpu_solve_sample(n_s, n)
        line["cpu_baseline"] = {
            "value": n_s * n_s / dt / 1e6, "unit": "Mcell/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "{0}x{0} corner of the workload, one solve ({1:.1f} s): scipy SuperLU, the reference's "
                      "solver='scipy' branch (linalg.py:139); SuperLU factorisation is single-threaded; the reference "
                      "arm (--impl reference) times the 1024^2 corner and prints the size sweep".format(n_s, dt)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_one_grid_multi_gpu(args, dist, rank, world, local, lib, _lib, core, breakdown):
    """ONE grid on all N GPUs, NCCL inside the library (fdfdpy_b200.distributed).

    sharded: the headline 4096^2 Ez solve with the elimination tree split over the ranks (DirectSolver(op, comm));
             efficiency_vs_n1 = t_factor(1 GPU, this run's breakdown) / (N * t_factor(N GPUs)).
             A 200 x 160 solve of the same code path is checked against the CPU oracle first (rank 0), so the
             line carries its own parity evidence for the NCCL path.
    slab:    the matrix-free stencils (Ez and Hz, 48 B/cell) on an 8192^2 grid split into row slabs with halo
             exchange, aggregate GB/s = 48 B x cells / max-over-ranks time."""
    import torch
    from fdfdpy_b200.distributed import Communicator, SlabOperator
    gloo = dist.new_group(backend="gloo")
    comm = Communicator.from_torch(group=gloo)
    out = {}

    def maxr(v):
        tt = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=gloo)
        return float(tt[0])

    # ---- parity of the sharded solve against the oracle (small grid, both polarisations)
    parity = {}
    rng = np.random.default_rng(1)
    nxp, nyp = 200, 160
    eps_p = 1 + 5 * (rng.random((nxp, nyp)) > 0.5)
    b_p = rng.standard_normal((nxp, nyp)) + 1j * rng.standard_normal((nxp, nyp))
    for pol in ("Ez", "Hz"):
        opp = core.MaxwellOperator(OMEGA0, eps_p, 0.04, [10, 8], pol, 1e-6)
        dp = core.DirectSolver(opp, comm=comm)
        xp = dp.solve(b_p).reshape(nxp, nyp)
        if rank == 0:
            from oracle import fdfd_oracle as orc
            ref = orc.sparse_solve(orc.construct_A(OMEGA0, eps_p, 0.04, [10, 8], pol, 1e-6), b_p).reshape(nxp, nyp)
            parity[pol] = {"rel_l2_vs_oracle": float(np.linalg.norm(xp - ref) / np.linalg.norm(ref)),
                           "relres": dp.last_relres}
        del dp, opp

    # ---- the headline grid on N GPUs
    n = args.size
    eps = synthetic_eps(n)
    src = synthetic_src(n)
    op = core.MaxwellOperator(OMEGA0, eps, DL, NPML, "Ez", L0)
    d = core.DirectSolver(op, comm=comm)
    f_ms, s_ms = [], []
    for rep in range(3):
        dist.barrier(group=gloo)
        ms_f, ms_s = C.c_double(0), C.c_double(0)
        _lib.check(lib.fdfd_timer_start(op.h))
        d.factor()
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(ms_f)))
        t0 = time.perf_counter()
        d.solve_fields(src, 1j * OMEGA0)
        ms_s.value = (time.perf_counter() - t0) * 1e3
        f_ms.append(maxr(ms_f.value))
        s_ms.append(maxr(ms_s.value))
    st = d.stats()
    fb, tb = C.c_double(0), C.c_double(0)
    lib.fdfd_mem_info(C.byref(fb), C.byref(tb))
    t1 = breakdown.get("factor_ms")
    out["sharded"] = {
        "what": "ONE {0}x{0} Ez grid on {1} GPUs: elimination tree split over the ranks, top fronts distributed by "
                "block rows, NCCL point-to-point inside the library".format(n, world),
        "factor_ms": min(f_ms[1:]), "factor_ms_all": f_ms, "solve_fields_host_ms": min(s_ms[1:]),
        "relres": d.last_relres, "factor_ms_1gpu_same_run": t1,
        "efficiency_vs_n1": (t1 / (world * min(f_ms[1:]))) if t1 else None,
        "speedup_vs_n1": (t1 / min(f_ms[1:])) if t1 else None,
        "factor_bytes_rank0": st["factor_bytes"], "hbm_used_gb_max": maxr((tb.value - fb.value) / 1e9),
        "parity_200x160": parity, "timing": "CUDA events on the library stream, max over ranks, best of 2 after 1 warm-up"}
    del d, op

    # ---- the same grid solved on the slab path: distributed BiCGSTAB right-preconditioned by restricted additive
    # Schwarz (per-slab direct factors, artificial PML at the cut faces); checked against the sharded direct solution
    def sumr(v):
        tt = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM, group=gloo)
        return float(tt[0])

    def schwarz_solve(ns_, eps_, src_, overlap=4, npml_sub=12, maxiter=600, ref=None):
        import gc
        gc.collect()
        fb0, tb0 = C.c_double(0), C.c_double(0)
        lib.fdfd_mem_info(C.byref(fb0), C.byref(tb0))
        slab = SlabOperator(OMEGA0, eps_, DL, NPML, "Ez", L0, comm=comm)
        dist.barrier(group=gloo)
        t0 = time.perf_counter()
        dsub = slab.setup_schwarz(eps_, overlap=overlap, npml_sub=npml_sub)
        _lib.check(lib.fdfd_op_sync(slab.h))
        setup_s = time.perf_counter() - t0
        sl = slice(slab.x0, slab.x1)
        b_loc = 1j * OMEGA0 * np.asarray(src_)[sl]
        # device-resident solve: right-hand side and iterate in the slab's extended layout, CUDA events around the loop
        nloc = (slab.nxl + 2) * ns_
        be, xe = slab.to_ext(b_loc), np.zeros((slab.nxl + 2, ns_), dtype=np.complex128)
        d_b, d_x = C.c_void_p(), C.c_void_p()
        _lib.check(lib.fdfd_malloc(C.byref(d_b), 16.0 * nloc))
        _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * nloc))
        _lib.check(lib.fdfd_memcpy_h2d(d_b, _lib.ptr(be), 16.0 * nloc))
        _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(xe), 16.0 * nloc))
        it, rr, conv, ms = C.c_int(0), C.c_double(0), C.c_int(0), C.c_double(0)
        dist.barrier(group=gloo)
        t0 = time.perf_counter()
        _lib.check(lib.fdfd_timer_start(slab.h))
        _lib.check(lib.fdfd_krylov_solve_dev(slab.h, None, d_b, d_x, 2, 1e-10, maxiter, 1, 80, None, 0, C.byref(it),
                                             C.byref(rr), C.byref(conv)))     # method 2: GMRES(80)
        _lib.check(lib.fdfd_timer_stop(slab.h, C.byref(ms)))
        solve_s = time.perf_counter() - t0
        _lib.check(lib.fdfd_memcpy_d2h(_lib.ptr(xe), d_x, 16.0 * nloc))
        lib.fdfd_free(d_b)
        lib.fdfd_free(d_x)
        xs, info = xe[1:-1], dict(iters=it.value, relres=rr.value, converged=bool(conv.value))
        fbs, tbs = C.c_double(0), C.c_double(0)
        lib.fdfd_mem_info(C.byref(fbs), C.byref(tbs))
        res = {"grid": [ns_, ns_], "slabs": world, "overlap_rows": overlap, "npml_sub": npml_sub,
               "subdomain_rows": slab.nxl + 2 * (overlap + npml_sub),
               "setup_ms_assemble_plus_factor": maxr(setup_s * 1e3), "solve_ms": maxr(ms.value),
               "solve_ms_host_wall": maxr(solve_s * 1e3),
               "krylov": "GMRES(80), right-preconditioned", "iterations": info["iters"],
               "preconditioner_applications": info["iters"] + 1,
               "relres": info["relres"], "converged": bool(info["converged"]),
               "factor_bytes_per_rank_max": maxr(dsub.stats()["factor_bytes"]),
               "hbm_gb_taken_by_this_solve_max": maxr((fb0.value - fbs.value) / 1e9)}
        if ref is not None:
            num = sumr(float(np.linalg.norm(xs - ref[sl]) ** 2))
            den = sumr(float(np.linalg.norm(ref[sl]) ** 2))
            res["rel_l2_vs_sharded_direct"] = (num / den) ** 0.5
        slab.drop_schwarz()
        del dsub, slab
        return res

    try:
        sw = {"what": "ONE Ez grid as {0} row slabs, matrix-free GMRES over NCCL right-preconditioned by restricted "
                      "additive Schwarz: every rank factorises its slab + 4 overlap rows + 12 artificial PML rows per "
                      "side (a local torus) with the direct solver; one overlap exchange and one local substitution "
                      "per application; tol 1e-10.  `device_*`: a waveguide device (ridge + side-coupled resonator, "
                      "bench.synthetic_device_eps); `crystal_4096_capped`: the headline random-rod crystal, where "
                      "one-level Schwarz is NOT competitive (multiple scattering between slabs), capped at 100 "
                      "iterations and reported for the record".format(world)}
        eps_dev, src_dev = synthetic_device_eps(n), synthetic_device_src(n)
        op = core.MaxwellOperator(OMEGA0, eps_dev, DL, NPML, "Ez", L0)
        d = core.DirectSolver(op, comm=comm)
        dist.barrier(group=gloo)
        t0 = time.perf_counter()
        x_dev = np.array(d.solve_fields(src_dev, 1j * OMEGA0)[0])
        direct_ms = maxr((time.perf_counter() - t0) * 1e3)
        del d, op
        sw["device_%d" % n] = schwarz_solve(n, eps_dev, src_dev, ref=x_dev)
        sw["device_%d" % n]["sharded_direct_factor_plus_solve_ms_host_wall"] = direct_ms
        del x_dev
        sw["crystal_%d_capped" % n] = schwarz_solve(n, eps, src, maxiter=100)
        if world >= 4 and not args.no_big_schwarz:
            n8 = 8192
            sw["device_8192"] = schwarz_solve(n8, synthetic_device_eps(n8), synthetic_device_src(n8))
        out["slab_schwarz"] = sw
    except Exception as e:
        out["slab_schwarz"] = {"error": repr(e)[:300]}

    # ---- slabs: matrix-free stencil with halo exchange
    ns = 8192
    eps_s = synthetic_eps(ns)
    peaks, _ = measured_peaks()
    slab_out = {"grid": [ns, ns]}
    for pol in ("Ez", "Hz"):
        slab = SlabOperator(OMEGA0, eps_s, DL, NPML, pol, L0, comm=comm)
        nloc = (slab.nxl + 2) * ns
        d_x, d_y = C.c_void_p(), C.c_void_p()
        _lib.check(lib.fdfd_malloc(C.byref(d_x), 16.0 * nloc))
        _lib.check(lib.fdfd_malloc(C.byref(d_y), 16.0 * nloc))
        xe = np.ones((slab.nxl + 2, ns), dtype=np.complex128)
        _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(xe), 16.0 * nloc))
        _lib.check(lib.fdfd_memcpy_h2d(d_y, _lib.ptr(xe), 16.0 * nloc))
        iters = 100
        for _ in range(5):
            _lib.check(lib.fdfd_op_apply_dev(slab.h, d_x, d_y, 1, 1))
        dist.barrier(group=gloo)
        ms = C.c_double(0)
        _lib.check(lib.fdfd_timer_start(slab.h))
        for _ in range(iters):
            _lib.check(lib.fdfd_op_apply_dev(slab.h, d_x, d_y, 1, 1))
        _lib.check(lib.fdfd_timer_stop(slab.h, C.byref(ms)))
        per = maxr(ms.value) / iters
        agg = 48.0 * ns * ns / (per * 1e-3) / 1e9
        slab_out[pol] = {"ms_per_apply": per, "agg_GBps": agg, "per_gpu_frac_of_hbm": agg / world / peaks["hbm_gbs"]}
        if pol == "Ez":
            # a fixed number of distributed BiCGSTAB iterations (2 stencils + 3 all-reduced reductions each)
            it, rr, conv = C.c_int(0), C.c_double(0), C.c_int(0)
            be = np.zeros((slab.nxl + 2, ns), dtype=np.complex128)
            if slab.x0 <= ns // 2 < slab.x1:
                be[1 + ns // 2 - slab.x0, ns // 2] = 1j * OMEGA0
            _lib.check(lib.fdfd_memcpy_h2d(d_y, _lib.ptr(be), 16.0 * nloc))
            _lib.check(lib.fdfd_memcpy_h2d(d_x, _lib.ptr(np.zeros_like(be)), 16.0 * nloc))
            for warm in (1, 0):
                dist.barrier(group=gloo)
                ms = C.c_double(0)
                _lib.check(lib.fdfd_timer_start(slab.h))
                _lib.check(lib.fdfd_krylov_solve_dev(slab.h, None, d_y, d_x, 0, 1e-30, 100, 1, 100, None, 0,
                                                     C.byref(it), C.byref(rr), C.byref(conv)))
                _lib.check(lib.fdfd_timer_stop(slab.h, C.byref(ms)))
            slab_out["bicgstab_ms_per_iter"] = maxr(ms.value) / max(it.value, 1)
        lib.fdfd_free(d_x)
        lib.fdfd_free(d_y)
        del slab
    slab_out["what"] = ("ONE 8192x8192 grid as {0} row slabs: fused matrix-free stencil (48 B/cell) incl. NCCL halo "
                        "exchange overlapped with the interior rows; device time, max over ranks").format(world)
    out["slab"] = slab_out
    del comm
    return out


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
    parts = {"factor_ms": 0.0, "solve_ms": 0.0}

    def step(k, timed=False):
        omega = my_freqs[k % len(my_freqs)]
        if omega not in ops:                        # operator handles are per omega (PML depends on it); plan is shared
            if len(ops) >= 2:
                ops.pop(next(iter(ops)))
            op = core.MaxwellOperator(omega, eps, DL, NPML, "Hz", L0)
            ops[omega] = (op, core.DirectSolver(op))
        op, direct = ops[omega]
        tms = C.c_double(0)
        _lib.check(lib.fdfd_op_assemble_dev(op.h, d_eps, None, 1))
        _lib.check(lib.fdfd_timer_start(op.h))
        _lib.check(lib.fdfd_direct_factor(direct.h, op.h))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(tms)))
        if timed:
            parts["factor_ms"] += tms.value
        _lib.check(lib.fdfd_timer_start(op.h))
        _lib.check(lib.fdfd_direct_solve_dev(direct.h, op.h, d_b, d_x, nrhs, 3, 1e-12, C.byref(relres), C.byref(steps_ref)))
        _lib.check(lib.fdfd_timer_stop(op.h, C.byref(tms)))
        if timed:
            parts["solve_ms"] += tms.value
        return relres.value

    for k in range(args.warmup):
        step(k % 2)
    if dist is not None:
        dist.barrier()
    t = time.perf_counter()
    worst = 0.0
    for k in range(args.steps):
        worst = max(worst, step(k % 2, timed=True))
    dt = time.perf_counter() - t
    if dist is not None:
        import torch
        tt = torch.tensor([dt, worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, worst = float(tt[0]), float(tt[1])
    if rank == 0:
        per = dt / args.steps
        fstats = ops[next(iter(ops))][1].stats()
        print(json.dumps({
            "metric": "fdfd_sweep_throughput", "value": world * nrhs * ncell / per / 1e6, "unit": "Mcell/s",
            "rhs_solves_per_s": world * nrhs / per, "factorisations_per_s": world / per, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "dtype": "f64", "data": "synthetic", "relres_max": worst,
            "refine_steps_last": steps_ref.value,
            "factor_ms_per_omega": parts["factor_ms"] / args.steps,
            "solve_refine_ms_per_omega": parts["solve_ms"] / args.steps,
            "factor_bytes": fstats["factor_bytes"],
            "substitution_floor_ms": 2 * fstats["factor_bytes"] / (measured_peaks()[0]["hbm_gbs"] * 1e9) * 1e3,
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
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the one-grid-on-N-GPUs arms")
    ap.add_argument("--no-big-schwarz", action="store_true", help="N >= 4: skip the 8192^2 Schwarz-preconditioned slab solve")
    ap.add_argument("--only-step", action="store_true", help="profiling runs only: exit right after the timed steps")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="reference arm: side of the sub-grid sample (default {})".format(CPU_REF_GRID))
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
