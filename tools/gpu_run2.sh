# round 2, GPU run 2: 3M GEMM + Hz flux stencil + distributed fronts (in-process ranks)
set -x
mkdir -p gpurun_out
export FDFD_LOCAL_TIMEOUT_S=60
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_distfront.py tests/test_gpu_simulation.py -m gpu -q --durations=10 --timeout=400 > gpurun_out/r2_02_pytest.log 2>&1
tail -60 gpurun_out/r2_02_pytest.log
STENCIL_POL=Hz python tools/stencil_bench.py 4096 > gpurun_out/r2_02_stencil_hz.log 2>&1
STENCIL_POL=Hz STENCIL_LOSSY=1 python tools/stencil_bench.py 4096 >> gpurun_out/r2_02_stencil_hz.log 2>&1
cat gpurun_out/r2_02_stencil_hz.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_02_bench.json 2> gpurun_out/r2_02_bench.err
tail -c 300 gpurun_out/r2_02_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_02_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['achieved_4m_equivalent'], d['roofline_stencil_hz']['achieved'], d['relres'], d['refine_steps'])"
python tools/zgemm_vs_cublas.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print(s) for s in d['shapes']]"
STENCIL_POL=Hz STENCIL_ONLY=1 timeout 300 ncu --set full --clock-control none -k regex:stencil_fused_hz -c 2 -o gpurun_out/r2_02_hz_stencil python tools/stencil_bench.py 4096 > gpurun_out/r2_02_ncu.log 2>&1
tail -3 gpurun_out/r2_02_ncu.log
