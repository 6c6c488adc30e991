# round 2, GPU run 3
set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_distfront.py tests/test_gpu_core.py -m gpu -q --durations=5 --timeout=400 > gpurun_out/r2_03_pytest.log 2>&1
tail -30 gpurun_out/r2_03_pytest.log
STENCIL_POL=Hz python tools/stencil_bench.py 4096 > gpurun_out/r2_03_stencil_hz.log 2>&1
STENCIL_POL=Hz STENCIL_LOSSY=1 python tools/stencil_bench.py 4096 >> gpurun_out/r2_03_stencil_hz.log 2>&1
cat gpurun_out/r2_03_stencil_hz.log
DIAG_BRIEF=1 python tools/diag_perf.py 4096 > gpurun_out/r2_03_diag.log 2>&1
cat gpurun_out/r2_03_diag.log | cut -c 1-260
PROFILE_ONLY=zgemm timeout 600 ncu --set full --import-source on --clock-control none -k regex:zgemm_dmma_persistent -s 2 -c 1 -o gpurun_out/r2_03_zgemm3m python tools/profile_kernels.py > gpurun_out/r2_03_ncu.log 2>&1
tail -4 gpurun_out/r2_03_ncu.log
