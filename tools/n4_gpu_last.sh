#!/bin/bash
# last N4 check of the round: the N4 GPU tests and the training bench tool on the final code
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_voc_train.py -q -s --timeout 120 -m gpu > gpurun_out/n4_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n4_tests3.log
grep -E "passed|failed|rc=" gpurun_out/n4_tests3.log | tail -3
grep -E "wav vs|gradients|input log-mel|^E  " gpurun_out/n4_tests3.log | tail -16
timeout 150 python tools/bench_voc_train.py --steps 3 --cpu > gpurun_out/n4_bench3.json 2> gpurun_out/n4_bench3.err; echo "bench rc=$?" >> gpurun_out/n4_bench3.err
tail -1 gpurun_out/n4_bench3.err; cut -c1-1400 gpurun_out/n4_bench3.json
