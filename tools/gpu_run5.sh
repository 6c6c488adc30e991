# round 2, GPU run 5 (4-GPU box): the N = 2 and N = 4 bench lines with the one-grid arms (real NCCL), config-4 substitution profile
set -x
mkdir -p gpurun_out
nvidia-smi -L
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/r2_05_bench_n$N.json 2> gpurun_out/r2_05_bench_n$N.err
  tail -c 600 gpurun_out/r2_05_bench_n$N.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_05_bench_n$N.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('sharded'))[:1500]); print(json.dumps(d.get('slab'))[:900])"
done
python tools/multirhs_probe.py 16 > gpurun_out/r2_05_multirhs.log 2>&1
cat gpurun_out/r2_05_multirhs.log
