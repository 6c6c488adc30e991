set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_distfront.py -x -q -k "schwarz" > gpurun_out/r2_13_pytest.log 2>&1; tail -5 gpurun_out/r2_13_pytest.log
(python tools/schwarz_probe.py 1024 4; python tools/schwarz_probe.py 2048 4; python tools/schwarz_probe.py 2048 8; python tools/schwarz_probe.py 4096 8; python tools/schwarz_probe.py 4096 8 8 16) > gpurun_out/r2_13_schwarz.log 2>&1; cat gpurun_out/r2_13_schwarz.log
