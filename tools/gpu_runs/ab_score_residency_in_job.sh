set -x
mkdir -p gpurun_out
rm -f gpurun_out/v15.jsonl
run() {
  timeout 600 python bench.py --steps 4 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e 2>>gpurun_out/v15.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'tag': '$1', 'ms': d['ms_per_step'], 'stage': d['roofline']['stage_ms_per_step'], 'clk': d['clocks']['sm_mhz'], 'sha': d['result_digest']['result_sha256'][:12]}))" >> gpurun_out/v15.jsonl
}
run base_6ctas
FY_SCORE_PAD=40000 run pad_5ctas
FY_SCORE_PAD=52000 run pad_4ctas
FY_H2_CFG=8 run build1536x2
FY_H2_CFG=5 run build1024x2
run base2
cat gpurun_out/v15.jsonl
