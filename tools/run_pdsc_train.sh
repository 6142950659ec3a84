#!/bin/bash
# GPU box: PointDSC training-step diagnostics + tests, each in its own process (a CUDA fault must not poison the next one)
mkdir -p gpurun_out
timeout 600 python tools/check_pdsc_train.py 2 2 256 300 1 > gpurun_out/pdsc_check_l2.log 2>&1; echo "check l2 rc $?"
timeout 600 python tools/check_pdsc_train.py 12 2 200 150 0 > gpurun_out/pdsc_check_l12.log 2>&1; echo "check l12 rc $?"
timeout 900 python -m pytest tests/test_pdsc_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pdsc_tests.log 2>&1; echo "tests rc $?"
tail -8 gpurun_out/pdsc_tests.log
timeout 600 python tools/bench_pdsc_train.py --iters 3 > gpurun_out/pdsc_train_bench.jsonl 2> gpurun_out/pdsc_train_bench.err; echo "bench rc $?"
timeout 300 python tools/bench_pdsc_train.py --iters 3 --precision tf32 --no-cpu >> gpurun_out/pdsc_train_bench.jsonl 2>> gpurun_out/pdsc_train_bench.err; echo "bench tf32 rc $?"
cat gpurun_out/pdsc_train_bench.jsonl; tail -3 gpurun_out/pdsc_train_bench.err
