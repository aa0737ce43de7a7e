#!/bin/bash
# 2 GPUs: new product paths -- context parallelism, in-model fused GEMM+RS, checkpoint load/save/resume
set -x
mkdir -p gpurun_out
S=gpurun_out/summary13.txt; rm -f $S
timeout 700 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "checkpoint" > gpurun_out/test_gpu_model_2gpu_c.log 2>&1
echo "exit tests: $?" >> $S; tail -25 gpurun_out/test_gpu_model_2gpu_c.log | cut -c1-500
cat $S
