#!/bin/bash
# first GPU call: kernel parity tests per file (separate processes) + kernel micro-benchmarks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
export CUDA_DEVICE_MAX_CONNECTIONS=32
for f in test_gpu_ops test_gpu_gemm test_gpu_collectives; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "exit $f: $?" >> gpurun_out/summary.txt
  tail -15 gpurun_out/$f.log
done
timeout 600 python scripts/bench_kernels.py gemm cast > gpurun_out/bench_kernels.jsonl 2> gpurun_out/bench_kernels.err
echo "exit bench: $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -30 gpurun_out/bench_kernels.jsonl
