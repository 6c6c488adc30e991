set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_simulation.py tests/test_gpu_distfront.py -m gpu -q --timeout=400 2>&1 | tail -6
python tools/diag_perf.py 4096 > gpurun_out/r2_10_diag.log 2>&1
grep -E "^   L0[0-5]|totals|^N=" gpurun_out/r2_10_diag.log | cut -c 1-200
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_10_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_4m_equivalent'], d['relres'], d['refine_steps'])"
