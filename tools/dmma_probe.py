"""DMMA issue-rate probe: TFLOP/s vs warps per SM and independent accumulator chains per warp."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from fdfdpy_b200 import _lib  # noqa: E402
lib = _lib.load()
_lib.require_gpu()
for warps in (4, 8, 16, 24, 32):
    row = []
    for nacc in (1, 2, 4, 8, 16):
        t = C.c_double(0)
        _lib.check(lib.fdfd_dmma_probe(warps, nacc, C.byref(t)))
        row.append(f"{t.value:6.2f}")
    print(f"warps/SM={warps:2d}  nacc 1,2,4,8,16 -> " + " ".join(row), flush=True)
