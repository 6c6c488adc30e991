# final-state evidence: bench line (native), then the ncu launch list of ONE timed step of the same command
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err
tail -c 300 gpurun_out/s23_bench.err
PER=$(python -c "
import json; d=json.loads([l for l in open('gpurun_out/s23_bench.json') if l.startswith('{')][-1]); print(int(d['gpu_launches'])//3)")
SKIP=$((PER - 60))
CNT=$((PER + 140))
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file gpurun_out/s23_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s23_ncu_bench.log 2>&1
tail -2 gpurun_out/s23_ncu_bench.log | cut -c 1-300
python -c "
import json; d=json.loads([l for l in open('gpurun_out/s23_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_stencil']['achieved'], d['gpu_launches'])"
