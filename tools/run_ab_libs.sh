#!/bin/bash
# A/B of library variants on the same box: usage run_ab_libs.sh <variant> ... (build/libgmf_<variant>.so; "base" = the in-tree library); 2 rounds
mkdir -p gpurun_out
QUICK="--no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train"
cp gmf_b200/libgmf_b200.so /tmp/base.so
for round in 1 2; do
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so gmf_b200/libgmf_b200.so; else cp build/libgmf_$v.so gmf_b200/libgmf_b200.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 $QUICK > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err || echo "[bench $v failed]"
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{v}.json').read().strip().splitlines()[-1])
    pk = d['roofline']['per_kernel']
    print(v, 'ms/step', round(d['ms_per_step'], 2), ' '.join(f"{k}={pk[k]['ms_per_step']:.2f}" for k in ('attn_sc', 'attn_fusion', 'ffn_geglu', 'pcn_qkv', 'fusion_q_proj', 'fusion_kv_proj') if k in pk))
except Exception as e:
    print(v, 'parse failed', e)
PY
done
done
cp /tmp/base.so gmf_b200/libgmf_b200.so
