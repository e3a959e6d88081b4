#!/bin/bash
# ncu --set full of the training-mode GATE GEMM (EPI 1) and the dz GEMM with the fused gate backward (EPI 4), one-product mode
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled \
  -k 'regex:fd_tapgemm_tc_kernel<\(int\)256, \(int\)64, \(int\)[14],' --launch-skip 58 -c 4 \
  -o gpurun_out/r02h_train_x1_gate python tools/bench_train.py --steps 1 --warmup 1 --precision f16x1 > gpurun_out/ncu_h_train.log 2>&1
ls -la gpurun_out/r02h_train* ; tail -3 gpurun_out/ncu_h_train.log
