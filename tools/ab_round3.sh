#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/ab_round3.log
: > $L
OLD=$PWD/build/ab/libfishdiff_old.so
run() { echo "=== $1" >> $L; shift; env "$@" >> $L 2>&1; }
run "respair tests new" timeout 600 python -m pytest tests/test_gpu_respair.py -x -q
for rep in 1 2; do
  run "prof old rep$rep" FISHDIFF_B200_LIB=$OLD C32_64=1 timeout 300 python tools/prof_respair.py
  run "prof new rep$rep" C32_64=1 timeout 300 python tools/prof_respair.py
  run "prof new OCC1 EW8 rep$rep" FD_RP_OCC=1 FD_RP_EW=8 C32_64=1 timeout 300 python tools/prof_respair.py
  run "voc old rep$rep" FISHDIFF_B200_LIB=$OLD timeout 300 python tools/voc_once.py
  run "voc new rep$rep" timeout 300 python tools/voc_once.py
done
tail -3 $L
