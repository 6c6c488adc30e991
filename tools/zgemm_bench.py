"""Time the batched complex GEMM on a few sweep-update shapes (device only)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
from fdfdpy_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.require_gpu()
shapes = [(8192, 8192, 64, 1), (8192, 8192, 32, 1), (4096, 4096, 64, 4), (2048, 2048, 64, 16), (1024, 1024, 64, 64),
          (448, 448, 64, 1024), (224, 224, 56, 4096), (64, 8192, 64, 1), (100, 100, 24, 16384)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
import os
lib.fdfd_zgemm_set_variant(int(os.environ.get("ZGEMM_VARIANT", "0")))
for (M, N, K, batch) in shapes:
    ms = C.c_double(0)
    _lib.check(lib.fdfd_zgemm_bench(M, N, K, batch, 1, 5, C.byref(ms)))
    fl = 8.0 * M * N * K * batch
    print(f"M={M} N={N} K={K} batch={batch}: {ms.value:.3f} ms  {fl / ms.value / 1e9:.2f} TFLOP/s", flush=True)
