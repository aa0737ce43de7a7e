#!/bin/bash
# multi-GPU validation (run with gpurun --gpus N): e2e strategy tests, collective microbench vs NCCL, bench at N
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -k "two_gpus or four_gpus" > gpurun_out/test_gpu_model_${N}gpu.log 2>&1
echo "exit test_gpu_model ${N}gpu: $?" > gpurun_out/summary4.txt
tail -25 gpurun_out/test_gpu_model_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 scripts/bench_collectives.py --max-mb 1024 > gpurun_out/collectives_${N}gpu.jsonl 2> gpurun_out/collectives_${N}gpu.err
echo "exit collectives: $?" >> gpurun_out/summary4.txt
cat gpurun_out/collectives_${N}gpu.jsonl; tail -5 gpurun_out/collectives_${N}gpu.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit bench: $?" >> gpurun_out/summary4.txt
cat gpurun_out/bench_${N}gpu.json; tail -8 gpurun_out/bench_${N}gpu.err
cat gpurun_out/summary4.txt
