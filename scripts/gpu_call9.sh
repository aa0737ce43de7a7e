#!/bin/bash
# diagnose fused GEMM+RS at N GPUs
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu.err
echo "exit fused_gemm_rs: $?"; cat gpurun_out/fused_gemm_rs_${N}gpu.jsonl; grep "error info\|arrival" gpurun_out/fused_gemm_rs_${N}gpu.err | cut -c1-300
HGB_DIAG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu_diag.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu_diag.err
echo "exit fused_gemm_rs diag: $?"; cat gpurun_out/fused_gemm_rs_${N}gpu_diag.jsonl; grep "error info\|arrival" gpurun_out/fused_gemm_rs_${N}gpu_diag.err | cut -c1-300
