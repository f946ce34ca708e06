set -x
mkdir -p gpurun_out
rm -f gpurun_out/v13.jsonl
for rep in 1; do
  for cfg in 0 1 8 6; do
    FY_H2_CFG=$cfg timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e 2>>gpurun_out/v13.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'cfg': '$cfg', 'ms': d['ms_per_step'], 'value': d['value'], 'stage': d['roofline']['stage_ms_per_step'], 'clk': d['clocks']['sm_mhz']}))" >> gpurun_out/v13.jsonl
  done
done
cat gpurun_out/v13.jsonl
