#!/bin/bash
# clock64 timeline of one tile of the persistent FFN kernel (GMF_FFN_TRACE build in build/libgmf_ffntrace.so)
mkdir -p gpurun_out
cp gmf_b200/libgmf_b200.so /tmp/orig.so; cp build/libgmf_ffntrace.so gmf_b200/libgmf_b200.so
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train --no-roofline > gpurun_out/bench_ffntrace.json 2> gpurun_out/bench_ffntrace.err; echo "[ffn trace exit $?]"
python tools/ffn_trace.py | tee gpurun_out/ffn_trace.txt
python tools/kv_trace.py | tee gpurun_out/kv_trace.txt
cp /tmp/orig.so gmf_b200/libgmf_b200.so
