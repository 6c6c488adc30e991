mkdir -p gpurun_out
python -m pytest tests/test_gpu_simulation.py tests/test_gpu_configs.py tests/test_gpu_core.py -x -q > gpurun_out/r2_26_pytest.log 2>&1; tail -4 gpurun_out/r2_26_pytest.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_26_bench.json 2> gpurun_out/r2_26_bench.err; tail -c 300 gpurun_out/r2_26_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_26_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
