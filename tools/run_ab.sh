#!/bin/bash
# micro-benchmark + quick test/bench of the current build
mkdir -p gpurun_out
./tools/ubench/sm_rates 2>&1 | grep -A3 "== T8"
bash tools/run_quick.sh
