set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not ml20m and not netflix" > gpurun_out/v1_pytest.log 2>&1
tail -15 gpurun_out/v1_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/v1_bench1.json 2> gpurun_out/v1_bench1.err
tail -3 gpurun_out/v1_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/v1_bench2.json 2> gpurun_out/v1_bench2.err
tail -3 gpurun_out/v1_bench2.err
