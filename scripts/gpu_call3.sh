#!/bin/bash
# bench bring-up on 1 GPU: smoke, truncated-model debug run, full bench, ncu launch list + full capture of the GEMM
set -x
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?" > gpurun_out/summary3.txt
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --layers 4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_l4.json 2> gpurun_out/bench_l4.err; echo "exit bench_l4: $?" >> gpurun_out/summary3.txt
tail -3 gpurun_out/bench_l4.err; cat gpurun_out/bench_l4.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit bench_full: $?" >> gpurun_out/summary3.txt
tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/summary3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_l2.csv python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo "exit ncu_list: $?" >> gpurun_out/summary3.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 30 -c 3 -o gpurun_out/prof_gemm python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "exit ncu_gemm: $?" >> gpurun_out/summary3.txt
cat gpurun_out/summary3.txt
