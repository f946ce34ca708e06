set -x
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 --workload netflix --no-cpu-baseline --no-secondary > gpurun_out/r02_netflix_1gpu.json 2> gpurun_out/r02_netflix_1gpu.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 3 --warmup 3 --workload netflix > gpurun_out/r02_netflix_${N}gpu.json 2> gpurun_out/r02_netflix_${N}gpu.err
fi
tail -2 gpurun_out/r02_netflix_${N}gpu.err
