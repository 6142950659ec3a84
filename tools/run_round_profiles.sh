#!/bin/bash
# round-end evidence: GPU tests, full bench + reference arm, ncu launch list of one small step, ncu --set full of the dominant kernels at cfg#2
# (launch order of a step after Fusion-1: kv_proj_all, then per layer pcn_qkv, sc_attn | fus_attn, ffn_fused; 51 matching launches per step; -s 53 skips the warm-up step and Fusion-1 of the second)
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "[ref exit $?]"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "[bench exit $?]"
tail -c 400 gpurun_out/bench_full.json; echo
Q="--no-cpu-baseline --no-e2e --no-roofline --no-backbone --no-cfg3 --no-train"
SMALL="python bench.py --pairs 16 --steps 1 --warmup 1 --min-warmup 1 $Q"
timeout 600 $SMALL > gpurun_out/plain_small.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu.log 2>&1; echo "[ncu list exit $?]"
BIG="python bench.py --pairs 64 --steps 1 --warmup 1 --min-warmup 1 $Q"
timeout 600 $BIG > gpurun_out/plain_big.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"sc_attn_v9|fus_attn_v2|ffn_fused|pcn_qkv_persist|kv_proj_all" -s 53 -c 5 -o gpurun_out/top5_cfg2 -f $BIG > gpurun_out/ncu_full.log 2>&1; echo "[ncu full exit $?]"
python __graft_entry__.py smoke
