set -x
mkdir -p gpurun_out
export FY_H2_CFG=3
python tools/one_cluster.py 2 ml20m > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_build_H2 -s 1 -c 1 -o gpurun_out/prof_h2_pf4 python tools/one_cluster.py 2 ml20m > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
