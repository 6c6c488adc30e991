# round 2, GPU run 4: pipelined 3M kernel, tensor-pipe multi-RHS substitution
set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_distfront.py tests/test_gpu_core.py -m gpu -q --durations=5 --timeout=400 > gpurun_out/r2_04_pytest.log 2>&1
tail -30 gpurun_out/r2_04_pytest.log
python tools/zgemm_vs_cublas.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print(s) for s in d['shapes']]"
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_04_bench.json 2> gpurun_out/r2_04_bench.err
tail -c 300 gpurun_out/r2_04_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_04_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_4m_equivalent'], d['roofline_stencil_hz']['achieved'], d['relres'], d['refine_steps'])"
python bench.py --workload sweep --steps 4 --warmup 2 > gpurun_out/r2_04_sweep.json 2> gpurun_out/r2_04_sweep.err
cat gpurun_out/r2_04_sweep.json | cut -c 1-700
tail -c 300 gpurun_out/r2_04_sweep.err
