set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py -m gpu -q -k "zgemm or direct_solver or many_rhs" --timeout=400 2>&1 | tail -4
python tools/zgemm_vs_cublas.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print(s) for s in d['shapes']]"
python tools/diag_perf.py 4096 2>&1 | tail -1
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_09_bench.json 2> gpurun_out/r2_09_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_09_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_4m_equivalent'], d['relres'], d['refine_steps'])"
