#!/bin/bash
# 2 GPUs: first hardware run of the opt-in VMM arena + NVLS all-reduce
set -x
mkdir -p gpurun_out
NVLS_MAX_MB=256 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $(nvidia-smi -L | wc -l) --master-addr 127.0.0.1 --master-port 29761 scripts/test_nvls.py > gpurun_out/nvls_ngpu.jsonl 2> gpurun_out/nvls_ngpu.err
echo "exit nvls: $?"; cat gpurun_out/nvls_ngpu.jsonl | cut -c1-400; grep -v "^\*\|OMP\|^$" gpurun_out/nvls_ngpu.err | grep -i "error\|bg_galvatron\|Traceback\|File \"/" | head -20 | cut -c1-300
