#!/bin/bash
# tests -> bench -> ncu launch list (each only after the previous exited 0)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
SMALL="python bench.py --pairs 16 --steps 1 --warmup 1 --min-warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 175 -c 180 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu.log 2>&1
echo "[ncu exit $?]"; tail -n 3 gpurun_out/ncu.log
