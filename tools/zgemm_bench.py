"""Time the batched complex GEMM on a few sweep-update shapes (device only)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
from fdfdpy_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.require_gpu()
# (M, N, K, batch, transb, lower)
shapes = [(8192, 8192, 64, 1, 0, 0), (8192, 8192, 2048, 1, 1, 1), (8192, 4096, 4096, 2, 1, 0), (8192, 8192, 2048, 1, 1, 0),
          (4096, 4096, 1024, 16, 1, 1), (4096, 1024, 1024, 16, 1, 0), (2048, 2048, 512, 64, 1, 1),
          (1024, 1024, 256, 256, 1, 1), (512, 512, 128, 1024, 1, 1), (256, 256, 64, 4096, 1, 1),
          (128, 128, 32, 16384, 1, 1), (64, 64, 16, 65536, 1, 1), (64, 8192, 64, 1, 0, 0), (24, 24, 3, 524288, 1, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
import os
lib.fdfd_zgemm_set_variant(int(os.environ.get("ZGEMM_VARIANT", "0")))
for (M, N, K, batch, tb, low) in shapes:
    ms = C.c_double(0)
    _lib.check(lib.fdfd_zgemm_bench(M, N, K, batch, 1, tb, low, 5, C.byref(ms)))
    fl = 8.0 * (M * (N + 64) / 2 if low else M * N) * K * batch
    print(f"M={M} N={N} K={K} batch={batch} transb={tb} lower={low}: {ms.value:.3f} ms  "
          f"{fl / ms.value / 1e9:.2f} TFLOP/s", flush=True)
