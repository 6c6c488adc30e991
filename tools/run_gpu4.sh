set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_core.py -x -q -k "zgemm or direct" > gpurun_out/r2_15_pytest.log 2>&1; tail -3 gpurun_out/r2_15_pytest.log
for cfg in "default" "FDFD_LA_HELPER=0" "ZGEMM_VARIANT=8"; do
  echo "== $cfg"; env $( [ "$cfg" = default ] || echo $cfg ) python tools/diag_perf.py 4096 2>&1 | grep -E "phases ms|^N=" | cut -c 1-330
done > gpurun_out/r2_15_diag.log 2>&1
cat gpurun_out/r2_15_diag.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err; tail -c 300 gpurun_out/r2_15_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_15_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_stencil_hz']['achieved'], d['relres'])"
