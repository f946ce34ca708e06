set -x
mkdir -p gpurun_out
for shape in ml20m netflix; do
  FY_BUILD_H=1 timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
  for cfg in 0 1 3 5; do
    FY_H2_BULK=1 FY_H2_CFG=$cfg timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
    FY_H2_BULK=0 FY_H2_CFG=$cfg timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
  done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic or overflow" > gpurun_out/ab1_pytest.log 2>&1
tail -5 gpurun_out/ab1_pytest.log
