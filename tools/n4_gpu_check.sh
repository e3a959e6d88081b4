#!/bin/bash
# N4 check on one box: the vocoder-training GPU tests (all of them, no -x) and the training bench tool.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_voc_train.py -q -s --timeout 150 -m gpu > gpurun_out/n4_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n4_tests.log
grep -E "passed|failed|rc=" gpurun_out/n4_tests.log | tail -3
if [ "$1" != "nobench" ]; then
  timeout 300 python tools/bench_voc_train.py --steps 3 --cpu > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "bench rc=$?" >> gpurun_out/n4_bench.err
  tail -1 gpurun_out/n4_bench.err; cut -c1-1500 gpurun_out/n4_bench.json
fi
