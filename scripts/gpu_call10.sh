#!/bin/bash
# 4 GPUs: fused GEMM+RS (PDL reducer), zero3 pool on real GPUs (2- and 4-rank strategy tests)
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
rm -f gpurun_out/summary10.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu.err
echo "exit fused_gemm_rs ${N}: $?" >> gpurun_out/summary10.txt; cat gpurun_out/fused_gemm_rs_${N}gpu.jsonl; grep "error info\|arrival" gpurun_out/fused_gemm_rs_${N}gpu.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29735 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_2gpu.jsonl 2> gpurun_out/fused_gemm_rs_2gpu.err
echo "exit fused_gemm_rs 2: $?" >> gpurun_out/summary10.txt; cat gpurun_out/fused_gemm_rs_2gpu.jsonl
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "four_gpus or (two_gpus and zero3)" > gpurun_out/test_gpu_model_zero3.log 2>&1
echo "exit tests zero3/4gpu: $?" >> gpurun_out/summary10.txt; tail -12 gpurun_out/test_gpu_model_zero3.log | cut -c1-600
cat gpurun_out/summary10.txt
