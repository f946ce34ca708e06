set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s8.json 2> gpurun_out/s8.err
tail -2 gpurun_out/s8.err
