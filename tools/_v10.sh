set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab1.jsonl
for shape in ml20m netflix; do
  timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic or overflow or sharded or fine_seam" > gpurun_out/v10_pytest.log 2>&1
tail -3 gpurun_out/v10_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/v10_bench.json 2> gpurun_out/v10_bench.err
