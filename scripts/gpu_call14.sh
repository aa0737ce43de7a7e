#!/bin/bash
# 1 GPU, what the driver runs at round end: the whole -m gpu suite, smoke(), the default bench and its reference arm
set -x
mkdir -p gpurun_out
S=gpurun_out/summary14.txt; rm -f $S
timeout 420 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_all.log 2>&1
echo "exit pytest -m gpu: $?" >> $S; tail -6 gpurun_out/test_gpu_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit smoke: $?" >> $S; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
echo "exit bench: $?" >> $S; cat gpurun_out/bench_n1_final.json | cut -c1-3000; tail -3 gpurun_out/bench_n1_final.err | cut -c1-300
cat $S
