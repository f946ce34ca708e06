set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "n_gpus or multi_process or shard_bounds or sharded" > gpurun_out/v12_multi.log 2>&1
tail -5 gpurun_out/v12_multi.log
