#!/bin/bash
QUICK="--no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train"
cp gmf_b200/libgmf_b200.so /tmp/orig.so
for v in "$@"; do
cp build/libgmf_$v.so gmf_b200/libgmf_b200.so
timeout 600 python bench.py --steps 2 --warmup 3 $QUICK --no-roofline > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "[bench trace $v exit $?]"
python tools/sc_trace.py > gpurun_out/sc_trace_$v.txt 2>&1; sed -n 20,30p gpurun_out/sc_trace_$v.txt; tail -n 33 gpurun_out/sc_trace_$v.txt
done
cp /tmp/orig.so gmf_b200/libgmf_b200.so
