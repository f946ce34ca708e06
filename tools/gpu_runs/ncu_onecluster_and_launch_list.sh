set -x
mkdir -p gpurun_out
python tools/one_cluster.py 2 ml20m > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_build_H2|k_score_f32|k_topn|k_refine_score' -s 4 -c 4 -o gpurun_out/prof_r02_final_onecluster python tools/one_cluster.py 2 ml20m > gpurun_out/ncu_r02_c.log 2>&1
tail -2 gpurun_out/ncu_r02_c.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/plain_bench3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 800 --csv --log-file gpurun_out/r02_launches_ml20m_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_r02_d.log 2>&1
tail -2 gpurun_out/ncu_r02_d.log
