"""The bar for zgemm.cuh: cuBLAS ZGEMM (through torch.matmul on complex128 CUDA tensors) on the same box, same shapes,
next to the hand-written DMMA kernel and the register-resident DMMA probe.  Prints one JSON line.

    python tools/zgemm_vs_cublas.py            # 8192^3 and two factorisation shapes
"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from fdfdpy_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.require_gpu()
import torch  # noqa: E402

out = {"shapes": []}
# (M, N, K, transb): NN and NT (C = A B^T, the Schur-update form)
for (M, N, K, tb) in [(8192, 8192, 8192, 0), (8192, 8192, 8192, 1), (8192, 8192, 512, 1), (4096, 4096, 4096, 1)]:
    a = torch.randn(M, K, dtype=torch.complex128, device="cuda")
    b = torch.randn((N, K) if tb else (K, N), dtype=torch.complex128, device="cuda")
    bt = b.t() if tb else b                    # a view: cuBLAS gets op(B) = T, no copy
    for _ in range(2):
        c = a @ bt
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        c = a @ bt
    e1.record()
    torch.cuda.synchronize()
    ms_cublas = e0.elapsed_time(e1) / reps
    del a, b, c, bt
    torch.cuda.empty_cache()
    ms = C.c_double(0)
    _lib.check(lib.fdfd_zgemm_bench(M, N, K, 1, 0, tb, 0, reps, C.byref(ms)))
    fl = 8.0 * M * N * K
    out["shapes"].append({"M": M, "N": N, "K": K, "transb": tb,
                          "cublas_zgemm_ms": ms_cublas, "cublas_tflops": fl / ms_cublas / 1e9,
                          "ours_ms": ms.value, "ours_tflops": fl / ms.value / 1e9})
probe = np.zeros(4)
rows = []
for warps in (4, 8, 16, 32):
    for nacc in (4, 8, 16):
        _lib.check(lib.fdfd_dmma_probe_clocked(warps, nacc, _lib.ptr(probe)))
        rows.append({"warps_per_sm": warps, "acc_chains": nacc, "tflops": probe[0], "sm_mhz": probe[1],
                     "frac_of_pipe_at_clock": probe[0] / (148 * 128 * probe[1] * 1e-6) if probe[1] else None})
out["dmma_probe"] = rows
print(json.dumps(out), flush=True)
