set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -x -q -m gpu --durations=6 > gpurun_out/v9_full.log 2>&1
tail -14 gpurun_out/v9_full.log
python __graft_entry__.py smoke > gpurun_out/v9_smoke.log 2>&1; tail -4 gpurun_out/v9_smoke.log
