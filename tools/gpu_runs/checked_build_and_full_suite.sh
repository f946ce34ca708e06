set -x
mkdir -p gpurun_out
FY_RM2_LIB=$PWD/filmyou_core_b200/libfilmyou_rm2_checked.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_neighbours.py tests/test_gpu_cooc.py -x -q -m gpu -k "not ml20m and not netflix and not multi_process and not n_gpus and not ml-1m and not cpp_host" > gpurun_out/v5_checked.log 2>&1
tail -3 gpurun_out/v5_checked.log
timeout 3000 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/v5_full.log 2>&1
tail -15 gpurun_out/v5_full.log
