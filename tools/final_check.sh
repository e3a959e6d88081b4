#!/bin/bash
# Round-end check on one box: full GPU suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" >> gpurun_out/final_bench.err
tail -2 gpurun_out/final_tests.log; tail -2 gpurun_out/final_smoke.log; tail -1 gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json
