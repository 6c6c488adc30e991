set -x
mkdir -p gpurun_out
STENCIL_POL=Hz python tools/stencil_bench.py 4096 1003 > gpurun_out/r2_11_hz.log 2>&1
STENCIL_POL=Hz STENCIL_LOSSY=1 python tools/stencil_bench.py 4096 517 >> gpurun_out/r2_11_hz.log 2>&1
cat gpurun_out/r2_11_hz.log
