export DIAG_BRIEF=1
for cfg in "1000 4" "1000 -384" "1000 -512" "1000 -768" "1000 -1024" "700 -512"; do
set -- $cfg
echo "=== split_min=$1 parts=$2"
FDFD_SPLIT_MIN=$1 FDFD_SPLIT_PARTS=$2 python tools/diag_perf.py 4096 2>&1 | grep -E "^N=|phases ms" | tail -2
done
