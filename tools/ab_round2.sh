#!/bin/bash
# Same-box A/B: previous commit's library (build/ab/libfishdiff_old.so) against the in-tree build -- sampler (GEMM2 operand
# prefetch), vocoder pass (fused-pair variants, block-sparse folded stage), then the full GPU suite on the new build.
mkdir -p gpurun_out
L=gpurun_out/ab_round2.log
: > $L
OLD=$PWD/build/ab/libfishdiff_old.so
run() { echo "=== $1" >> $L; shift; env "$@" >> $L 2>&1; }
S="python bench.py --steps 2 --warmup 3 --evals 20 --no-vocoder --no-cpu-baseline --no-e2e --no-train --no-extras"
run "full gpu suite new" timeout 1200 python -m pytest tests -m gpu -x -q
for rep in 1 2; do
  run "sampler old rep$rep" FISHDIFF_B200_LIB=$OLD timeout 300 $S
  run "sampler new rep$rep" timeout 300 $S
  run "voc old rep$rep" FISHDIFF_B200_LIB=$OLD timeout 300 python tools/voc_once.py
  run "voc new rep$rep" timeout 300 python tools/voc_once.py
done
run "prof new" timeout 300 python tools/prof_respair.py
tail -3 $L
