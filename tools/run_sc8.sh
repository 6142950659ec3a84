#!/bin/bash
# validate + time the gen-8 SC attention variants (GMF_SC_IMPL = 8 / 9 / 10: 0 / 1 / 2 of every 4 exponentials on the FMA pipe)
mkdir -p gpurun_out
for impl in ${IMPLS:-8 10}; do
  GMF_SC_IMPL=$impl timeout 600 python -m pytest tests -m gpu -x -q -k "sc_ or encoder or forward or drop_in or full_size" > gpurun_out/pytest_sc$impl.log 2>&1
  echo "[impl $impl pytest exit $?]"; tail -n 6 gpurun_out/pytest_sc$impl.log
done
bash tools/run_env_sweep.sh ${SWEEP:-GMF_SC_IMPL=8 GMF_SC_IMPL=9 GMF_SC_IMPL=10 GMF_SC_IMPL=1}
