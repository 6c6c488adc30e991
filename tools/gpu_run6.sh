# round 2, GPU run 6 (1 GPU): device-resident nonlinear loops, small-K mrhs tiles, config-4 sweep
set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_simulation.py tests/test_gpu_core.py -m gpu -q --durations=5 --timeout=400 > gpurun_out/r2_06_pytest.log 2>&1
tail -30 gpurun_out/r2_06_pytest.log
python tools/multirhs_probe.py 16 > gpurun_out/r2_06_multirhs.log 2>&1
head -3 gpurun_out/r2_06_multirhs.log
python bench.py --workload sweep --steps 4 --warmup 2 > gpurun_out/r2_06_sweep.json 2> gpurun_out/r2_06_sweep.err
cat gpurun_out/r2_06_sweep.json | cut -c 1-600
