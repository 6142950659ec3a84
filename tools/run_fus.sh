#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "flash or fusion or forward or drop_in or full_size or encoder" > gpurun_out/pytest_fus.log 2>&1
echo "[pytest exit $?]"; tail -n 8 gpurun_out/pytest_fus.log
bash tools/run_env_sweep.sh ${SWEEP:-GMF_FUS_IMPL=2 GMF_FUS_IMPL=1}
