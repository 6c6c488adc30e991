mkdir -p gpurun_out
for t in 128 256 512 1024; do
  echo "== FDFD_BLOCK_GJ_MAXTILES=$t"; FDFD_BLOCK_GJ_MAXTILES=$t python tools/diag_perf.py 4096 2>&1 | grep -E "phases ms|^N=|^   L(1[4-9]|2[04]|3[29]) " | cut -c 1-200
done > gpurun_out/r2_21_diag.log 2>&1
cat gpurun_out/r2_21_diag.log
