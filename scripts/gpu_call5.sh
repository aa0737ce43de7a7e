#!/bin/bash
# 1 GPU: kernel tests after the comm-kernel restructure, fused optimizer, N=1 bench with the fused optimizer and fewer checkpoints
set -x
mkdir -p gpurun_out
for f in test_gpu_collectives test_gpu_ops; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1; echo "exit $f: $?" >> gpurun_out/summary5.txt; tail -6 gpurun_out/$f.log
done
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x --timeout 600 -k "one_gpu or smoke" > gpurun_out/test_gpu_model_1.log 2>&1; echo "exit model1: $?" >> gpurun_out/summary5.txt; tail -5 gpurun_out/test_gpu_model_1.log
for ck in 0 8 16; do
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --optimizer fused --checkpoint-layers $ck > gpurun_out/bench_fused_ck$ck.json 2> gpurun_out/bench_fused_ck$ck.err
  rc=$?; echo "exit bench fused ck$ck: $rc" >> gpurun_out/summary5.txt; tail -3 gpurun_out/bench_fused_ck$ck.err; cat gpurun_out/bench_fused_ck$ck.json
  if [ $rc -eq 0 ]; then break; fi
done
timeout 300 python scripts/bench_kernels.py cast > gpurun_out/bench_kernels2.jsonl 2>&1; cat gpurun_out/bench_kernels2.jsonl
cat gpurun_out/summary5.txt
