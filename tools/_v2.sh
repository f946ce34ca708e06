set -x
mkdir -p gpurun_out
rm -f gpurun_out/ab1.jsonl
for shape in ml20m netflix; do
  timeout 300 python tools/one_cluster.py 4 $shape >> gpurun_out/ab1.jsonl 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic or overflow or compact" > gpurun_out/v2_pytest_quick.log 2>&1
tail -3 gpurun_out/v2_pytest_quick.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "ml20m_full" > gpurun_out/v2_pytest_ml20m.log 2>&1
tail -8 gpurun_out/v2_pytest_ml20m.log
