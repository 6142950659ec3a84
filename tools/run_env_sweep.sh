#!/bin/bash
# bench-only sweep over env settings: usage run_env_sweep.sh "VAR=a" "VAR=b" ...
mkdir -p gpurun_out
i=0
for e in "$@"; do
  env $e timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - "$e" gpurun_out/sweep_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
kp = d['kernel_profile']
print(sys.argv[1], 'te_mm', round(d['config']['max_translation_error_vs_gt_mm'],3), 'pairs/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'attn_sc', round(kp['attn_sc']['ms_per_step'],2), 'attn_fusion', round(kp['attn_fusion']['ms_per_step'],2))
PY
  i=$((i+1))
done
