#!/bin/bash
# 1 GPU: validate cuDNN attention path + full N=1 bench on the searched strategy, ncu launch list and captures
set -x
mkdir -p gpurun_out; rm -f gpurun_out/summary7.txt
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x --timeout 600 -k "one_gpu or smoke" > gpurun_out/test_gpu_model_1b.log 2>&1; echo "exit model1: $?" >> gpurun_out/summary7.txt; tail -5 gpurun_out/test_gpu_model_1b.log
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1_searched.json 2> gpurun_out/bench_n1_searched.err; echo "exit bench: $?" >> gpurun_out/summary7.txt
tail -3 gpurun_out/bench_n1_searched.err; cat gpurun_out/bench_n1_searched.json
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --checkpoint-layers 0 > gpurun_out/bench_n1_nockpt.json 2> gpurun_out/bench_n1_nockpt.err; echo "exit bench nockpt: $?" >> gpurun_out/summary7.txt
cat gpurun_out/bench_n1_nockpt.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_l2_v2.csv python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l2_v2.log 2>&1; echo "exit ncu_list: $?" >> gpurun_out/summary7.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"all_gather_push|reduce_scatter_adamw" -c 4 -o gpurun_out/prof_comm python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_comm.log 2>&1; echo "exit ncu_comm: $?" >> gpurun_out/summary7.txt
cat gpurun_out/summary7.txt
