#!/bin/bash
# time the SC / fusion attention kernels with development variants of the library (build/libgmf_*.so)
mkdir -p gpurun_out
QUICK="--no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train"
cp gmf_b200/libgmf_b200.so /tmp/orig.so
for v in "$@"; do
  cp build/libgmf_$v.so gmf_b200/libgmf_b200.so
  timeout 600 python bench.py --steps 3 --warmup 3 $QUICK > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "[bench $v exit $?]"
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{v}.json').read().strip().splitlines()[-1])
    pk=d['roofline']['per_kernel']
    print(v, 'ms/step', round(d['ms_per_step'],2), ' '.join(f"{k}={pk[k]['ms_per_step']:.2f}" for k in ('attn_sc','attn_fusion','ffn_geglu','pcn_qkv')))
except Exception as e:
    print(v, 'parse failed', e)
PY
done
cp /tmp/orig.so gmf_b200/libgmf_b200.so
