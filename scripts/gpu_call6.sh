#!/bin/bash
# N GPUs: pipeline tests after the TMA-context fix, fused GEMM+reduce-scatter, collectives v2, attention libraries, bench N
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
rm -f gpurun_out/summary6.txt
timeout 300 python scripts/bench_kernels.py attn > gpurun_out/attn_libs.jsonl 2>&1; cat gpurun_out/attn_libs.jsonl
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -k "pp2 or fused_gemm or tp2_dp2 or hybrid" > gpurun_out/test_gpu_model_pp_${N}gpu.log 2>&1
echo "exit test pp/fused ${N}gpu: $?" >> gpurun_out/summary6.txt; tail -30 gpurun_out/test_gpu_model_pp_${N}gpu.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 scripts/test_fused_gemm_rs.py > gpurun_out/fused_gemm_rs_${N}gpu.jsonl 2> gpurun_out/fused_gemm_rs_${N}gpu.err
echo "exit fused_gemm_rs: $?" >> gpurun_out/summary6.txt; cat gpurun_out/fused_gemm_rs_${N}gpu.jsonl; tail -5 gpurun_out/fused_gemm_rs_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 scripts/bench_collectives.py --max-mb 1024 > gpurun_out/collectives_${N}gpu_v2.jsonl 2> gpurun_out/collectives_${N}gpu_v2.err
echo "exit collectives: $?" >> gpurun_out/summary6.txt; cat gpurun_out/collectives_${N}gpu_v2.jsonl; tail -3 gpurun_out/collectives_${N}gpu_v2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_fused.json 2> gpurun_out/bench_${N}gpu_fused.err
echo "exit bench: $?" >> gpurun_out/summary6.txt; cat gpurun_out/bench_${N}gpu_fused.json; tail -5 gpurun_out/bench_${N}gpu_fused.err
cat gpurun_out/summary6.txt
