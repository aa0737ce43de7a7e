#!/bin/bash
# One 8-GPU lease: stand-alone collectives and fused GEMM+collective micro-benchmarks at p = 8, the reference's own 8-GPU hybrid
# corpus on the product path, and the path legs that only exist at N = 8 (Llama-3-70B ZeRO-3, TP = 8).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 240 29711 scripts/bench_collectives.py --max-mb 1024 --sizes-mb 4,16,64,256,1024 > gpurun_out/r02_collectives_8gpu.jsonl 2> gpurun_out/r02_collectives_8gpu.err
tail -2 gpurun_out/r02_collectives_8gpu.jsonl | cut -c1-300
run 240 29712 scripts/test_fused_collectives.py > gpurun_out/r02_fused_8gpu.jsonl 2> gpurun_out/r02_fused_8gpu.err
tail -1 gpurun_out/r02_fused_8gpu.jsonl; tail -2 gpurun_out/r02_fused_8gpu.err | cut -c1-300
run 600 29500 bench.py --gpus 8 --legs-only --legs llama3-70b_zero3_ckpt_dp8,tp8_megatron_sp,tp8,ulysses8 > gpurun_out/r02_legs_8gpu.json 2> gpurun_out/r02_legs_8gpu.err
tail -c 1500 gpurun_out/r02_legs_8gpu.json; tail -3 gpurun_out/r02_legs_8gpu.err | cut -c1-300
timeout 420 python -m pytest tests/test_gpu_model.py -q -m gpu -k "eight_gpus" > gpurun_out/r02_tests_8gpu.log 2>&1
tail -5 gpurun_out/r02_tests_8gpu.log
