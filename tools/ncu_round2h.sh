#!/bin/bash
# ncu --set full captures of the final fused-pair variants and of the training-mode GATE GEMM (one launch each)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
for cfg in "128 7 1" "64 7 3" "32 7 3"; do
  set -- $cfg
  C=$1 K=$2 D=$3 N=3 timeout 600 $NCU -k regex:fd_respair --launch-skip 2 -c 1 -o gpurun_out/r02h_respair_c$1k$2 \
    python tools/prof_respair_one.py > gpurun_out/ncu_h_c$1.log 2>&1
done
# training step, one-product mode: the forward GATE GEMM (EPI 1) and the dz GEMM with the fused gate backward (EPI 4)
timeout 900 $NCU --kernel-name-base demangled -k "regex:fd_tapgemm_tc_kernel<256, 64, [14]," --launch-skip 58 -c 4 \
  -o gpurun_out/r02h_train_x1_gate python tools/bench_train.py --steps 1 --warmup 1 --precision f16x1 > gpurun_out/ncu_h_train.log 2>&1
ls -la gpurun_out/r02h_* | tail
