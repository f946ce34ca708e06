set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab1.jsonl
for shape in ml20m netflix; do
  timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
  FY_H2_BULK=0 timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/v8_bench.json 2> gpurun_out/v8_bench.err
