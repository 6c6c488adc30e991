# 2-GPU bench line (replicas + sharded + slab + slab_schwarz) and the GPU tests that need >= 2 devices
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_24_bench_n2.json 2> gpurun_out/r2_24_bench_n2.err; tail -c 600 gpurun_out/r2_24_bench_n2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_24_bench_n2.json') if l.startswith('{')][-1]); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['e2e']['ms_per_step']); print(d['sharded']); print(d['slab']); print(json.dumps(d.get('slab_schwarz'), indent=1))"
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
