mkdir -p gpurun_out
B="python bench.py --no-secondary --no-cpu-baseline --no-e2e"
timeout 55 $B --steps 3 --warmup 3 > gpurun_out/v19_a.json 2> gpurun_out/v19_a.err
FY_H_BUFS=3 timeout 110 $B --steps 2 --warmup 3 --workload netflix > gpurun_out/v19_b.json 2> gpurun_out/v19_b.err
R=$((185-SECONDS)); if [ $R -gt 25 ]; then FY_H_BUFS=3 timeout $R $B --steps 3 --warmup 3 > gpurun_out/v19_c.json 2> gpurun_out/v19_c.err; fi
echo elapsed $SECONDS
