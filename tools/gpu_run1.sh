# round 2, GPU run 1: the whole GPU suite (with durations), the bench line, the cuBLAS comparator
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2_01_pytest.log 2>&1
tail -40 gpurun_out/r2_01_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
tail -c 400 gpurun_out/r2_01_bench.err
python tools/zgemm_vs_cublas.py > gpurun_out/r2_01_zgemm_vs_cublas.json 2> gpurun_out/r2_01_zgemm_vs_cublas.err
tail -c 600 gpurun_out/r2_01_zgemm_vs_cublas.json
python bench.py --workload sweep --steps 4 --warmup 2 > gpurun_out/r2_01_sweep.json 2> gpurun_out/r2_01_sweep.err
cat gpurun_out/r2_01_sweep.json | cut -c 1-900
