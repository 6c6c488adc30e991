set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"thread_kernel|merge_gather|child_scatter|leaf_gather|leaf_scatter" -c 14 -f -o gpurun_out/r2_16_smallsolve python tools/multirhs_probe.py 16 > gpurun_out/r2_16_ncu.log 2>&1
ncu -i gpurun_out/r2_16_smallsolve.ncu-rep --page raw --csv > gpurun_out/r2_16_smallsolve.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_16_smallsolve.raw.csv > gpurun_out/r2_16_smallsolve.txt
rm -f gpurun_out/r2_16_smallsolve.ncu-rep
grep -E "===|duration|dram__bytes|inst_executed.sum|warps_active|long_scoreboard|lg_throttle|mio_throttle|registers" gpurun_out/r2_16_smallsolve.txt | cut -c 1-150
