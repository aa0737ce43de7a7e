#!/bin/bash
# ncu evidence on ONE GPU (never a bench value): (1) --set full of every kernel of this library inside one training step of 2
# real-size Llama-3-8B layers, one microbatch (cudaProfilerStart/Stop around the step: bench.py --ncu-step); (2) the same for the GPT
# family's LayerNorm / bias-GeLU kernels; (3) the per-launch time list of the bench command on 2 layers.  Cross-rank kernels are
# captured at p = 1 only (a kernel that waits for a peer cannot be replayed).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
OURS='regex:gemm_bf16_kernel|all_gather_push|reduce_scatter|all_reduce|all_to_all|tile_reduce|qkv_rope|rmsnorm|layernorm|swiglu|bias_gelu|ce_rowmax|ce_sumexp|ce_bwd|cast_kernel'
COMMON="--no-cpu-baseline --no-probe --steps 1 --warmup 1 --layers 2 --strategy configs/ncu_2layers_1microbatch.json"
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$OURS" -c 110 -f -o gpurun_out/r02_ncu_llama_2layers \
    python bench.py $COMMON --ncu-step > gpurun_out/r02_ncu_llama_2layers.log 2>&1
echo "== llama rc=$?"; tail -2 gpurun_out/r02_ncu_llama_2layers.log | cut -c1-200
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:layernorm|bias_gelu' -c 16 -f -o gpurun_out/r02_ncu_gpt_rows \
    python bench.py $COMMON --ncu-step --model gpt-6.7b --seq 2048 > gpurun_out/r02_ncu_gpt_rows.log 2>&1
echo "== gpt rc=$?"; tail -2 gpurun_out/r02_ncu_gpt_rows.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_ncu_launches_bench_layers2.csv \
    python bench.py --no-cpu-baseline --no-probe --steps 1 --warmup 1 --layers 2 > gpurun_out/r02_ncu_launches_bench_layers2.log 2>&1
echo "== launches rc=$?"; wc -l gpurun_out/r02_ncu_launches_bench_layers2.csv; ls -la gpurun_out/*.ncu-rep
