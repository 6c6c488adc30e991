# round 2, GPU run 7 (8-GPU box): bench line at N = 8 (replicas + one-grid arms), 8192^2 sharded at N = 2 and 8, config-4 sweep on 8
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/r2_07_bench_n8.json 2> gpurun_out/r2_07_bench_n8.err
tail -c 600 gpurun_out/r2_07_bench_n8.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_07_bench_n8.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('sharded'))[:1500]); print(json.dumps(d.get('slab'))[:900])"
for N in 2 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/dist_check.py --parity "" --size 8192 --reps 2 > gpurun_out/r2_07_dist8192_n$N.json 2> gpurun_out/r2_07_dist8192_n$N.err
  tail -c 300 gpurun_out/r2_07_dist8192_n$N.err
  cat gpurun_out/r2_07_dist8192_n$N.json | cut -c 1-700
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --workload sweep --steps 8 --warmup 2 > gpurun_out/r2_07_sweep_n8.json 2> gpurun_out/r2_07_sweep_n8.err
cat gpurun_out/r2_07_sweep_n8.json | cut -c 1-500
