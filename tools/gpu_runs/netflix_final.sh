set -x
mkdir -p gpurun_out
timeout 280 python bench.py --gpus 1 --steps 3 --warmup 3 --workload netflix --no-cpu-baseline --no-secondary > gpurun_out/r02_netflix_1gpu_final.json 2> gpurun_out/r02_netflix_1gpu_final.err
tail -2 gpurun_out/r02_netflix_1gpu_final.err
