#!/bin/bash
# usage: run_ncu_full.sh <kernel-regex> <skip> <count> <outname>   (plain run first, then ncu --set full)
mkdir -p gpurun_out
SMALL="python bench.py --pairs ${NCU_PAIRS:-8} --steps 1 --warmup 1 --min-warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "[ncu exit $?]"; tail -n 4 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
