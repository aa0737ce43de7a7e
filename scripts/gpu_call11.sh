#!/bin/bash
# 8 GPUs: bench N=8 (searched strategy), collectives vs NCCL at p=8, fused GEMM+RS at p=8, Llama-3-70B ZeRO-3+ckpt (BASELINE config 5),
# NVLS all-reduce at p=8, the reference's 8-GPU hybrid corpus on the product path
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary11.txt; rm -f $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29733 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit bench: $?" >> $S; cat gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
timeout 300 $TR --master-port 29732 scripts/bench_collectives.py --max-mb 1024 > gpurun_out/collectives_${N}gpu.jsonl 2> gpurun_out/collectives_${N}gpu.err
echo "exit collectives: $?" >> $S; tail -4 gpurun_out/collectives_${N}gpu.jsonl; tail -3 gpurun_out/collectives_${N}gpu.err | cut -c1-300
timeout 300 $TR --master-port 29731 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu.err
echo "exit fused_gemm_rs: $?" >> $S; cat gpurun_out/fused_gemm_rs_${N}gpu.jsonl; grep "error info" gpurun_out/fused_gemm_rs_${N}gpu.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29736 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_4gpu.jsonl 2> gpurun_out/fused_gemm_rs_4gpu.err
echo "exit fused_gemm_rs 4: $?" >> $S; cat gpurun_out/fused_gemm_rs_4gpu.jsonl
timeout 900 $TR --master-port 29734 bench.py --gpus $N --model llama3-70b --strategy configs/galvatron_config_llama3-70b_8gpus_zero3_ckpt.json --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_70b_zero3_${N}gpu.json 2> gpurun_out/bench_70b_zero3_${N}gpu.err
echo "exit bench 70b zero3: $?" >> $S; cat gpurun_out/bench_70b_zero3_${N}gpu.json; grep -v "^\*\|OMP" gpurun_out/bench_70b_zero3_${N}gpu.err | tail -8 | cut -c1-300
NVLS_MAX_MB=1024 timeout 300 $TR --master-port 29737 scripts/test_nvls.py > gpurun_out/nvls_${N}gpu.jsonl 2> gpurun_out/nvls_${N}gpu.err
echo "exit nvls: $?" >> $S; tail -12 gpurun_out/nvls_${N}gpu.jsonl | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "eight_gpus" > gpurun_out/test_gpu_model_${N}gpu.log 2>&1
echo "exit tests eight_gpus: $?" >> $S; tail -5 gpurun_out/test_gpu_model_${N}gpu.log | cut -c1-400
cat $S
