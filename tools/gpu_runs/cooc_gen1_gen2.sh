set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cooc.py -x -q -m gpu -k "tiny or small or ml-100k or nonpositive" > gpurun_out/v6_cooc_small.log 2>&1
tail -3 gpurun_out/v6_cooc_small.log
timeout 900 python -m pytest tests/test_gpu_cooc.py -x -q -m gpu > gpurun_out/v6_cooc.log 2>&1
tail -3 gpurun_out/v6_cooc.log
FY_COOC_V1=1 timeout 600 python tools/cooc_bench.py ml-1m ml-20m > gpurun_out/v6_cooc_bench_v1.json 2>gpurun_out/v6.err
timeout 600 python tools/cooc_bench.py ml-1m ml-20m > gpurun_out/v6_cooc_bench_v2.json 2>>gpurun_out/v6.err
cat gpurun_out/v6_cooc_bench_v1.json gpurun_out/v6_cooc_bench_v2.json
