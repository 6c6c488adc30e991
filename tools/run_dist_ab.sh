# 8-GPU A/B of the sharded 4096^2 factorisation (dist_check: factor timing + per-level phases of ranks 0, N/2, N-1)
N=${1:-8}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/dist_check.py --parity "" --size 4096 --reps 2 $PH > gpurun_out/r2_25_dist_$tag.log 2>&1; echo "== $tag"; grep -E "factor_ms_all_reps|^rank|^   r" gpurun_out/r2_25_dist_$tag.log | cut -c 1-400 | sed 's/"others".*//' ; }
PH=--phases run base FDFD_DIST_GJ_GROUP=0
PH= run gj8 FDFD_DIST_GJ_GROUP=8
PH= run gj4 FDFD_DIST_GJ_GROUP=4
PH= run steps16_gj8 FDFD_DIST_GJ_GROUP=8 FDFD_SPLIT_MAX_STEPS=16
PH= run steps16_base FDFD_DIST_GJ_GROUP=0 FDFD_SPLIT_MAX_STEPS=16
