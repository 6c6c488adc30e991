# usage: bash tools/run_dist_scaling.sh "8"   or   "4 2"   (world sizes to run on this box)
set -x
for N in $1; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/dist_check.py --parity 200x160 --size 8192 --reps 1 --slab-parity 40x36 --slab-size 8192 > gpurun_out/s22_dist_n$N.log 2>&1
tail -c 300 gpurun_out/s22_dist_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/dist_check.py --parity "" --size 4096 --reps 2 > gpurun_out/s22_dist4096_n$N.log 2>&1
tail -c 300 gpurun_out/s22_dist4096_n$N.log
done
