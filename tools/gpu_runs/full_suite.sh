set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/v17_full.log 2>&1
tail -14 gpurun_out/v17_full.log
