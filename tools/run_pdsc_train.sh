#!/bin/bash
# GPU box: PointDSC training-step tests + bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pdsc_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pdsc_tests.log 2>&1; echo "tests rc $?"
tail -12 gpurun_out/pdsc_tests.log
timeout 600 python tools/bench_pdsc_train.py --iters 3 > gpurun_out/pdsc_train_bench.jsonl 2> gpurun_out/pdsc_train_bench.err; echo "bench rc $?"
timeout 300 python tools/bench_pdsc_train.py --iters 3 --precision tf32 --no-cpu >> gpurun_out/pdsc_train_bench.jsonl 2>> gpurun_out/pdsc_train_bench.err; echo "bench tf32 rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/pdsc_train_bench.jsonl'):
    d = json.loads(l); print(d['precision'], 'ms/step', round(d['ms_per_step'], 2), 'pairs/s', round(d['value'], 1), 'launches', d['gpu_launches_per_step'], 'ws GB', round(d['workspace_gb'], 1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['pairs_per_s'], 2))
PY
timeout 300 python tools/profile_pdsc_train.py tf32x3 > gpurun_out/pdsc_train_profile_x3.txt 2>&1
