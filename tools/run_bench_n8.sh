# N-GPU bench line (replicas + sharded + slab + slab_schwarz); usage: bash tools/run_bench_n8.sh N TAG
N=${1:-8}; TAG=${2:-r2_19}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -c 400 gpurun_out/${TAG}_bench_n$N.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/${TAG}_bench_n$N.json') if l.startswith('{')][-1]); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches','clocks']}, d['e2e']['value'], d['e2e']['ms_per_step']); print(d['sharded']); print(d['slab']); print(json.dumps(d.get('slab_schwarz'), indent=1))"
