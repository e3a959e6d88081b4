#!/bin/bash
# wave-quantisation-aware BLOCK_N for LINEAR / GATE_BWD: parity tests, then same-box A/B of the training step
mkdir -p gpurun_out
L=gpurun_out/ab_round5.log
: > $L
OLD=$PWD/build/ab/libfishdiff_old.so
run() { echo "=== $1" >> $L; shift; env "$@" >> $L 2>&1; }
run "tests new" timeout 900 python -m pytest tests/test_gpu_tapgemm.py tests/test_gpu_wavenet.py tests/test_gpu_train.py tests/test_gpu_r2_golden.py tests/test_gpu_vocoder.py tests/test_gpu_mel.py -x -q
for rep in 1 2; do
  run "train x1 old rep$rep" FISHDIFF_B200_LIB=$OLD timeout 300 python tools/bench_train.py --steps 20 --warmup 5 --precision f16x1
  run "train x1 new rep$rep" timeout 300 python tools/bench_train.py --steps 20 --warmup 5 --precision f16x1
  run "train f16 old rep$rep" FISHDIFF_B200_LIB=$OLD timeout 300 python tools/bench_train.py --steps 10 --warmup 3 --precision f16
  run "train f16 new rep$rep" timeout 300 python tools/bench_train.py --steps 10 --warmup 3 --precision f16
done
grep -E "passed|failed|error" $L | head; grep -E "^=== |ms_per_step" $L | sed -E 's/.*"ms_per_step": ([0-9.]+).*/   \1/' | paste - - | head -20
