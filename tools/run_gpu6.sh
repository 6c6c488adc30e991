set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_core.py tests/test_gpu_distfront.py -x -q -k "krylov or schwarz" > gpurun_out/r2_18_pytest.log 2>&1; tail -12 gpurun_out/r2_18_pytest.log
(python tools/schwarz_probe.py 2048 4; python tools/schwarz_probe.py 4096 8; python tools/schwarz_probe.py 1024 4 4 12 rods; python tools/schwarz_probe.py 2048 4 4 12 rods) > gpurun_out/r2_18_schwarz.log 2>&1; cat gpurun_out/r2_18_schwarz.log
