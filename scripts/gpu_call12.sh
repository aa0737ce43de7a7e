#!/bin/bash
# 2 GPUs: ZeRO-3 + checkpointing with REAL layer sizes (8 Llama-3-8B layers): concurrent all-gather / reduce-scatter at 296 CTAs each
set -x
mkdir -p gpurun_out
S=gpurun_out/summary12.txt; rm -f $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
HGB_TIMEOUT_MS=20000 timeout 400 $TR --master-port 29734 bench.py --gpus 2 --layers 8 --strategy configs/debug_llama3-8b_2gpus_zero3_ckpt.json --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_zero3_8layers_2gpu.json 2> gpurun_out/bench_zero3_8layers_2gpu.err
echo "exit bench zero3 8 layers: $?" >> $S; cat gpurun_out/bench_zero3_8layers_2gpu.json | cut -c1-1500; grep -v "^\*\|OMP" gpurun_out/bench_zero3_8layers_2gpu.err | grep -i "error\|info" | head -5 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "two_gpus" > gpurun_out/test_gpu_model_2gpu_b.log 2>&1
echo "exit tests 2gpu: $?" >> $S; tail -5 gpurun_out/test_gpu_model_2gpu_b.log | cut -c1-400
cat $S
