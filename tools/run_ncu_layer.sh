#!/bin/bash
# full ncu capture of every kernel of one encoder layer (12 launches after Fusion-1 + layer 0)
mkdir -p gpurun_out
SMALL="python bench.py --pairs ${NCU_PAIRS:-37} --steps 1 --warmup 1 --min-warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 && \
timeout 2400 ncu --set full --clock-control none --import-source on -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-12} -o gpurun_out/$1 -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "[ncu exit $?]"; tail -n 3 gpurun_out/ncu_full.log | cut -c1-300
