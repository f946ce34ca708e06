set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_neighbours.py tests/test_jni_stub.py -x -q -m gpu -k "not ml20m and not netflix and not multi_process and not n_gpus" > gpurun_out/v16_pytest.log 2>&1
tail -3 gpurun_out/v16_pytest.log
FY_RM2_LIB=$PWD/filmyou_core_b200/libfilmyou_rm2_checked.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic or sharded or overflow" > gpurun_out/v16_checked.log 2>&1
tail -2 gpurun_out/v16_checked.log
python __graft_entry__.py smoke > gpurun_out/v16_smoke.log 2>&1; tail -3 gpurun_out/v16_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_ml20m_1gpu_final.json 2> gpurun_out/v16_bench.err
tail -1 gpurun_out/v16_bench.err
