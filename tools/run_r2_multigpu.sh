#!/bin/bash
# round 2: N-GPU legs (N = $1): headline bench (weak cfg#2 + strong cfg#3) and the DGR training step with its NCCL all-reduce
N=$1
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-backbone > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "[bench $N exit $?]"
tail -n 3 gpurun_out/bench_${N}gpu.err
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_dgr_train.py --iters 30 --no-cpu > gpurun_out/dgr_train_${N}gpu.jsonl 2> gpurun_out/dgr_train_${N}gpu.err; echo "[dgr train $N exit $?]"
grep -m3 -E "NVLS|nRanks|via P2P" gpurun_out/dgr_train_${N}gpu.err | cut -c1-200
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = [json.loads(l) for l in open(f'gpurun_out/bench_{n}gpu.json') if l.startswith('{')][-1]
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'clocks', d['clocks'])
print('strong', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['strong_scaling'].items() if k != 'workload'})
for l in [x for x in open(f'gpurun_out/dgr_train_{n}gpu.jsonl') if x.startswith('{')]:
    t = json.loads(l); print(t['workload'][:40], 'ms/step', round(t['ms_per_step'], 3), t['split_ms'], t['allreduce'])
PY
