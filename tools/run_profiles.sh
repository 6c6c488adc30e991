set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s17_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s17_ncu_bench.log 2>&1
tail -2 gpurun_out/s17_ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:'stencil_fused_ez_kernel|stencil_planes_kernel|zgemm_dmma_persistent_kernel' -s 2 -c 6 -o gpurun_out/s17_full python tools/profile_kernels.py > gpurun_out/s17_ncu_full.log 2>&1
tail -3 gpurun_out/s17_ncu_full.log
python bench.py --steps 3 --warmup 3 > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err
tail -c 300 gpurun_out/s17_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s17_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['breakdown'], d['roofline']['frac'])"
