"""DMMA issue-order probe (fdfd_dmma_pattern_probe): fraction of the FP64 tensor pipe the 3M inner loop's 48 DMMAs
reach in different issue orders, register operands only."""
import sys
import numpy as np
sys.path.insert(0, ".")
from fdfdpy_b200 import _lib  # noqa: E402
lib = _lib.load()
_lib.require_gpu()
out = np.zeros(4)
names = ["k-step order of the GEMM", "accumulator-major (2 k-steps back to back)", "term-major", "tile-major"]
for warps in (8, 4):
    for pat in range(4):
        _lib.check(lib.fdfd_dmma_pattern_probe(pat, warps, _lib.ptr(out)))
        print(f"warps/SM={warps} pattern {pat} ({names[pat]}): {out[0]:.2f} TFLOP/s executed, {out[3]*100:.1f} % of the pipe at {out[1]:.0f} MHz", flush=True)
for mode, what in enumerate(["LDS fragments + DADD sums, no barrier", "+ a CTA barrier per k-tile of 32", "+ 16 cp.async per thread per k-tile"]):
    _lib.check(lib.fdfd_dmma_smem_probe(mode, _lib.ptr(out)))
    print(f"smem probe mode {mode} ({what}): {out[0]:.2f} TFLOP/s executed, {out[3]*100:.1f} % of the pipe at {out[1]:.0f} MHz", flush=True)
