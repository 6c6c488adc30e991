set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_distfront.py -x -q -k "schwarz or slab" > gpurun_out/r2_12_pytest.log 2>&1; tail -15 gpurun_out/r2_12_pytest.log
python -m pytest tests -x -q -m gpu -k "hz or Hz or stencil" >> gpurun_out/r2_12_pytest.log 2>&1; tail -5 gpurun_out/r2_12_pytest.log
STENCIL_POL=Hz python tools/stencil_bench.py 4096 1003 517 > gpurun_out/r2_12_hz.log 2>&1; cat gpurun_out/r2_12_hz.log
(python tools/schwarz_probe.py 1024 4; python tools/schwarz_probe.py 2048 4; python tools/schwarz_probe.py 2048 8) > gpurun_out/r2_12_schwarz.log 2>&1; cat gpurun_out/r2_12_schwarz.log
