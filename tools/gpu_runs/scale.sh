set -x
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_scale_1gpu.json 2> gpurun_out/r02_scale_1gpu.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err
fi
tail -2 gpurun_out/r02_scale_${N}gpu.err
