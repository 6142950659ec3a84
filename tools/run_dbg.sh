#!/bin/bash
# timing experiments: rebuild with GMF_SC_DBG variants ON THE BOX and bench only
mkdir -p gpurun_out
for v in ${DBG_LIST:-1 2 0}; do
  GMF_SC_DBG=$v python gmf_b200/build.py > /dev/null 2>&1
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg$v.json 2> gpurun_out/bench_dbg$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_dbg$v.json').read().strip().splitlines()[-1])
kp = d['kernel_profile']
print('DBG=$v', 'ms/step', round(d['ms_per_step'],2), 'attn_sc', round(kp['attn_sc']['ms_per_step'],2), 'attn_fusion', round(kp['attn_fusion']['ms_per_step'],2))
PY
done
