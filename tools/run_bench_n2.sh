python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/s20_bench_n2.json 2> gpurun_out/s20_bench_n2.err; tail -c 400 gpurun_out/s20_bench_n2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/s20_bench_n2.json') if l.startswith('{')][-1]); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches','breakdown']}, d['e2e']['ms_per_step'], d['roofline']['frac'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference 2>&1 | tail -2 | cut -c 1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 4 --warmup 2 --workload sweep 2>&1 | tail -1 | cut -c 1-600
python bench.py --steps 4 --warmup 2 --workload sweep 2>&1 | tail -1 | cut -c 1-400
