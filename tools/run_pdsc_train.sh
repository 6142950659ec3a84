#!/bin/bash
# GPU box: PointDSC training-step tests + bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pdsc_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pdsc_tests.log 2>&1; echo "tests rc $?"
tail -25 gpurun_out/pdsc_tests.log
