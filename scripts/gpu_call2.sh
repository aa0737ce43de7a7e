#!/bin/bash
# e2e model parity on N GPUs (N = number visible)
set -x
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -x --timeout 600 > gpurun_out/test_gpu_model.log 2>&1
echo "exit test_gpu_model: $?" > gpurun_out/summary2.txt
tail -40 gpurun_out/test_gpu_model.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 > gpurun_out/test_gpu_ops.log 2>&1
echo "exit test_gpu_ops: $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
