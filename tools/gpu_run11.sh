set -x
python tools/dmma_pattern.py 2>&1 | tail -4
python tools/diag_perf.py 4096 2>&1 | grep -E "^   L0[0-5]|^N=" | cut -c 1-200
