# round 2, GPU run 6 (1 GPU): device-resident nonlinear loops, small-K mrhs tiles, look-ahead A/B, config-4 sweep
set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_simulation.py tests/test_gpu_core.py tests/test_gpu_distfront.py -m gpu -q --durations=5 --timeout=400 > gpurun_out/r2_06_pytest.log 2>&1
tail -30 gpurun_out/r2_06_pytest.log
FDFD_LOOKAHEAD=0 python tools/diag_perf.py 4096 > gpurun_out/r2_06_diag_nolookahead.log 2>&1
tail -1 gpurun_out/r2_06_diag_nolookahead.log
python tools/diag_perf.py 4096 > gpurun_out/r2_06_diag_full.log 2>&1
tail -1 gpurun_out/r2_06_diag_full.log
python tools/multirhs_probe.py 16 > gpurun_out/r2_06_multirhs.log 2>&1
head -3 gpurun_out/r2_06_multirhs.log
python bench.py --workload sweep --steps 4 --warmup 2 > gpurun_out/r2_06_sweep.json 2> gpurun_out/r2_06_sweep.err
cat gpurun_out/r2_06_sweep.json | cut -c 1-600
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_06_bench.json 2> gpurun_out/r2_06_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_06_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_4m_equivalent'], d['relres'], d['refine_steps'])"
