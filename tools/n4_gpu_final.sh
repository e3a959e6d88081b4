#!/bin/bash
# N4 round-end check on one box: full GPU suite (incl. the N4 tests), the training bench tool, ncu launch list of one step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/n4_full_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n4_full_tests.log
grep -E "passed|failed|rc=" gpurun_out/n4_full_tests.log | tail -3
grep -E "wav vs|gradients|input log-mel" gpurun_out/n4_full_tests.log | tail -12
timeout 200 python tools/bench_voc_train.py --steps 3 --cpu > gpurun_out/n4_bench2.json 2> gpurun_out/n4_bench2.err; echo "bench rc=$?" >> gpurun_out/n4_bench2.err
tail -1 gpurun_out/n4_bench2.err; cut -c1-1200 gpurun_out/n4_bench2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_voc_train_launches.csv python tools/voc_train_once.py > gpurun_out/n4_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/n4_ncu.log
python tools/launch_table.py gpurun_out/r02i_voc_train_launches.csv 2>&1 | head -40
