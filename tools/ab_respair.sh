#!/bin/bash
# Same-box A/B of fused-pair kernel variants: correctness of each variant first, then alternating timing runs.
#   old  = build/ab/libfishdiff_old.so (the previous commit's library, FISHDIFF_B200_LIB)
#   new  = in-tree library; FD_RP_EW=16 / FD_RP_OCC=2 select the epilogue-warp / two-CTAs-per-SM variants
mkdir -p gpurun_out
L=gpurun_out/ab_respair.log
: > $L
run() { echo "=== $1" >> $L; shift; env "$@" >> $L 2>&1; }
run "pytest new default" timeout 600 python -m pytest tests/test_gpu_respair.py -x -q
run "pytest new EW16" FD_RP_EW=16 timeout 600 python -m pytest tests/test_gpu_respair.py -x -q
run "pytest new OCC2" FD_RP_OCC=2 timeout 600 python -m pytest tests/test_gpu_respair.py -x -q
for rep in 1 2; do
  run "prof old rep$rep" FISHDIFF_B200_LIB=$PWD/build/ab/libfishdiff_old.so timeout 300 python tools/prof_respair.py
  run "prof new rep$rep" timeout 300 python tools/prof_respair.py
  run "prof new EW16 rep$rep" FD_RP_EW=16 C32_64=1 timeout 300 python tools/prof_respair.py
  run "prof new OCC2 rep$rep" FD_RP_OCC=2 C=32 timeout 300 python tools/prof_respair.py
done
tail -5 $L
