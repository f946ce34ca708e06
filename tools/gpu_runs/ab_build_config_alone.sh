set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab1.jsonl
for cfg in 0 6 7 8; do
  FY_H2_CFG=$cfg timeout 300 python tools/one_cluster.py 4 ml20m >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
  FY_H2_CFG=$cfg timeout 300 python tools/one_cluster.py 4 netflix >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.jsonl
