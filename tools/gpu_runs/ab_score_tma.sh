set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab1.jsonl
for shape in ml20m netflix; do
  FY_SCORE_TMA=0 timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
  FY_SCORE_TMA=1 timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not ml20m and not netflix and not multi_process and not n_gpus" > gpurun_out/v7_pytest.log 2>&1
tail -3 gpurun_out/v7_pytest.log
FY_SCORE_TMA=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/v7_bench_ldg.json 2> gpurun_out/v7_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/v7_bench_tma.json 2>> gpurun_out/v7_bench.err
