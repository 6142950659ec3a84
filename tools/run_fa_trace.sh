#!/bin/bash
# per-CTA phase timings of the fusion attention kernel (GMF_FFN_TRACE build prints them) + a quick bench of the product library
mkdir -p gpurun_out
cp gmf_b200/libgmf_b200.so /tmp/orig.so; cp build/libgmf_ffntrace.so gmf_b200/libgmf_b200.so
timeout 600 python bench.py --steps 1 --warmup 1 --min-warmup 1 --no-cpu-baseline --no-e2e --no-backbone --no-cfg3 --no-train --no-roofline 2>&1 | grep "fus_attn trace" | head -${TRACE_LINES:-6}
python tools/ffn_trace.py | head -3
cp /tmp/orig.so gmf_b200/libgmf_b200.so
bash tools/run_ab_libs.sh base
