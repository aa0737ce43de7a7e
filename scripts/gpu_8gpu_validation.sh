#!/bin/bash
# One 8-GPU lease, bounded to LIMIT seconds in total: the driver's own N = 8 bench command with all six path legs (70B ZeRO-3
# first), the reference's 8-GPU hybrid corpus on the product path, then the stand-alone collectives and the fused GEMM+collective
# kernels at p = 8.  Every step gets min(its own limit, what is left).  Outputs -> gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
LIMIT=${LIMIT:-700}
T0=$(date +%s)
left() { local l=$(( LIMIT - ($(date +%s) - T0) )); [ "$l" -lt "$1" ] && echo "$l" || echo "$1"; }
t=$(left 520); timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
    bench.py --gpus 8 --steps 4 --warmup 3 --total-budget-s $(( t - 20 )) \
    --legs llama3-70b_zero3_ckpt_dp8,tp8_megatron_sp,tp8,ulysses8,pp2_tp4_1f1b,zero3_ckpt_dp8 \
    > gpurun_out/r02_bench_n8_with_legs.json 2> gpurun_out/r02_bench_n8_with_legs.err
echo "== bench rc=$? at $(( $(date +%s) - T0 )) s"; tail -c 600 gpurun_out/r02_bench_n8_with_legs.json; tail -3 gpurun_out/r02_bench_n8_with_legs.err | cut -c1-300
t=$(left 190); if [ "$t" -gt 40 ]; then
  timeout "$t" python -m pytest tests/test_gpu_model.py -q -m gpu -k "eight_gpus and (hybrid or tp1248_vtp8_sp or tp2_cp2)" > gpurun_out/r02_tests_8gpu.log 2>&1
  echo "== tests rc=$? at $(( $(date +%s) - T0 )) s"; tail -4 gpurun_out/r02_tests_8gpu.log | cut -c1-300
fi
t=$(left 70); if [ "$t" -gt 40 ]; then
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 \
      scripts/bench_collectives.py --max-mb 1024 --sizes-mb 4,64,1024 > gpurun_out/r02_collectives_8gpu.jsonl 2> gpurun_out/r02_collectives_8gpu.err
  echo "== collectives rc=$? at $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/r02_collectives_8gpu.jsonl | cut -c1-300
fi
t=$(left 60); if [ "$t" -gt 30 ]; then
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 \
      scripts/test_fused_collectives.py > gpurun_out/r02_fused_8gpu.jsonl 2> gpurun_out/r02_fused_8gpu.err
  echo "== fused rc=$? at $(( $(date +%s) - T0 )) s"; tail -1 gpurun_out/r02_fused_8gpu.jsonl | cut -c1-300; tail -2 gpurun_out/r02_fused_8gpu.err | cut -c1-300
fi
