#!/bin/bash
# every step in its own process (a device trap poisons the CUDA context) and under a timeout
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,memory.total --format=csv >> $LOG 2>&1
for s in ${@:-linear attn attn_sc fusion layer tail e2e big}; do
  timeout 300 python tools/gpu_probe.py $s >> $LOG 2>&1
  echo "[exit $?] $s" >> $LOG
done
tail -c 6000 $LOG
