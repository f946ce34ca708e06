set -x
mkdir -p gpurun_out
rm -f gpurun_out/v14.jsonl
run() {
  timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e 2>>gpurun_out/v14.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'tag': '$1', 'ms': d['ms_per_step'], 'stage': d['roofline']['stage_ms_per_step'], 'clk': d['clocks']['sm_mhz']}))" >> gpurun_out/v14.jsonl
}
run base
FY_H2_PAD=4096 run pad4k_7ctas
FY_H2_PAD=12288 run pad12k_6ctas
FY_H2_CFG=8 run 1536x1
FY_H2_CFG=8 FY_H2_PAD=4096 run 1536x1_pad4k
run base2
cat gpurun_out/v14.jsonl
