set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_core.py tests/test_gpu_distfront.py -x -q -k "direct or dist or residual" > gpurun_out/r2_20_pytest.log 2>&1; tail -4 gpurun_out/r2_20_pytest.log
for cfg in "default" "FDFD_BLOCK_GJ=0"; do
  echo "== $cfg"; env $( [ "$cfg" = default ] || echo $cfg ) python tools/diag_perf.py 4096 2>&1 | grep -E "phases ms|^N=" | cut -c 1-330
done > gpurun_out/r2_20_diag.log 2>&1
cat gpurun_out/r2_20_diag.log
