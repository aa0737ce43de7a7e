#!/bin/bash
# N GPUs (4 or 8): multi-GPU strategy tests (N<=4), fused GEMM+RS, collectives vs NCCL, bench at N
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
rm -f gpurun_out/summary8_${N}.txt
if [ "$N" -le 4 ]; then
  timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "four_gpus" > gpurun_out/test_gpu_model_${N}gpu.log 2>&1
  echo "exit tests ${N}gpu: $?" >> gpurun_out/summary8_${N}.txt; tail -12 gpurun_out/test_gpu_model_${N}gpu.log | cut -c1-600
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu.err
echo "exit fused_gemm_rs: $?" >> gpurun_out/summary8_${N}.txt; cat gpurun_out/fused_gemm_rs_${N}gpu.jsonl; tail -5 gpurun_out/fused_gemm_rs_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 scripts/bench_collectives.py --max-mb 1024 > gpurun_out/collectives_${N}gpu.jsonl 2> gpurun_out/collectives_${N}gpu.err
echo "exit collectives: $?" >> gpurun_out/summary8_${N}.txt; cat gpurun_out/collectives_${N}gpu.jsonl; tail -3 gpurun_out/collectives_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit bench: $?" >> gpurun_out/summary8_${N}.txt; cat gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
cat gpurun_out/summary8_${N}.txt
