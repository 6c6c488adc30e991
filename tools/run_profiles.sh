# Round-2 evidence at HEAD, everything lands in gpurun_out/ (copied to profiles/r02_* by the author):
#  1. the native bench line (not under a profiler) + the reference arm + config 4
#  2. the ncu launch list (gpu__time_duration.sum) of exactly ONE timed bench step
#  3. ncu --set full captures of the dominant kernels, exported as raw CSV, and the traffic file bench.py reads
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 300 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --workload sweep --steps 6 --warmup 2 > gpurun_out/r02_bench_sweep_config4.json 2> gpurun_out/r02_bench_sweep_config4.err
python tools/zgemm_vs_cublas.py > gpurun_out/r02_zgemm_vs_cublas.json 2>/dev/null
python tools/dmma_pattern.py > gpurun_out/r02_dmma_probes.txt 2>&1
PER=$(python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench.json') if l.startswith('{')][-1]); print(int(d['gpu_launches'])//3)")
# warm-up step + probes come first: skip to the timed step (the launch list is cut to it by tools/summarize_launches.py --last)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_all.csv python bench.py --steps 1 --warmup 1 --only-step > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c 1-300
python tools/cut_step.py gpurun_out/r02_launches_all.csv $PER > gpurun_out/r02_launches_bench_step.csv
python tools/summarize_launches.py gpurun_out/r02_launches_bench_step.csv > gpurun_out/r02_launches_bench_step_summary.txt
cat gpurun_out/r02_launches_bench_step_summary.txt
rm -f gpurun_out/r02_launches_all.csv
# full captures: 3M persistent zgemm on a chain-step Schur update, Ez / Hz fused stencils, the 16-RHS substitution kernel
PROFILE_ONLY=zgemm timeout 600 ncu --set full --import-source on --clock-control none -k regex:zgemm_dmma_persistent -s 2 -c 1 -f -o gpurun_out/r02_zgemm3m python tools/profile_kernels.py > gpurun_out/r02_ncu_zgemm.log 2>&1
PROFILE_ONLY=stencil timeout 600 ncu --set full --clock-control none -k "regex:stencil_fused_ez|stencil_march_hz" -c 4 -f -o gpurun_out/r02_stencils python tools/profile_kernels.py > gpurun_out/r02_ncu_stencil.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mrhs_dmma -s 40 -c 2 -f -o gpurun_out/r02_mrhs python tools/multirhs_probe.py 16 > gpurun_out/r02_ncu_mrhs.log 2>&1
for f in r02_zgemm3m r02_stencils r02_mrhs; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/$f.raw.csv > gpurun_out/${f/r02_/r02_full_capture_}.txt; done
rm -f gpurun_out/r02_zgemm3m.ncu-rep gpurun_out/r02_stencils.ncu-rep gpurun_out/r02_mrhs.ncu-rep
python tools/ncu_traffic.py gpurun_out/r02_zgemm3m.raw.csv zgemm_dmma_persistent --flops 2.923e11 --bytes 1.76e9 --round 2 \
  --launch "chain-step Schur update S(9727x9727, lower) -= G(9727x512) F_RE^T: 6 flops per complex MAC (3M) x 9727 x 9791/2 x 512; bytes = S read+write + G + F_RE" --out gpurun_out/zgemm_capture.json
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['breakdown'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_stencil']['achieved'], d['roofline_stencil_hz']['achieved'], d['gpu_launches'])"
