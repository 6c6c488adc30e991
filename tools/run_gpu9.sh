mkdir -p gpurun_out
python -m pytest tests/test_gpu_core.py -x -q -k "direct" > gpurun_out/r2_22_pytest.log 2>&1; tail -3 gpurun_out/r2_22_pytest.log
python tools/diag_perf.py 4096 2>&1 | grep -E "phases ms|^N=|^   L0[0-5] " | cut -c 1-200 > gpurun_out/r2_22_diag.log; cat gpurun_out/r2_22_diag.log
python tools/diag_perf.py 1000 2>&1 | grep -E "^N=" | cut -c 1-200
