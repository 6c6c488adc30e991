set -x
mkdir -p gpurun_out
python tools/dmma_pattern.py > gpurun_out/r2_08_dmma_pattern.log 2>&1
cat gpurun_out/r2_08_dmma_pattern.log
export FDFD_LOCAL_TIMEOUT_S=60
timeout 600 python -m pytest tests/test_gpu_distfront.py -m gpu -q -k "lookahead or multi_rhs" --timeout=400 2>&1 | tail -5
python tools/diag_perf.py 4096 2>&1 | tail -1
