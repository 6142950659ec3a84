#!/bin/bash
# GPU box: ncu evidence for the PointDSC training step (launch list of one small step + full metric sets of its dominant kernels)
mkdir -p gpurun_out
CMD="python tools/bench_pdsc_train.py --iters 1 --no-cpu --batch 4"
timeout 300 $CMD > gpurun_out/train_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1; echo "[ncu list exit $?]"
BIG="python tools/bench_pdsc_train.py --iters 1 --no-cpu"
timeout 300 $BIG > gpurun_out/train_plain_big.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"img_gemm_kernel|mat_to_img_b|sm_loss_fused|softmax_mul_rows" -s 6000 -c 40 -o gpurun_out/train_top -f $BIG > gpurun_out/train_ncu_full.log 2>&1; echo "[ncu full exit $?]"
