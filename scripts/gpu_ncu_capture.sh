#!/bin/bash
# ncu evidence on ONE GPU (never a bench value): (1) --set full of every kernel of this library inside one training step of ONE
# real-size Llama-3-8B layer, one microbatch (cudaProfilerStart/Stop around the step: bench.py --ncu-step); (2) the same for the GPT
# family's LayerNorm / bias-GeLU kernels; (3) the per-launch time list of one step of the bench command on 2 layers.  Cross-rank
# kernels are captured at p = 1 only (a kernel that waits for a peer cannot be replayed).  The reports are exported to CSV on the box
# and deleted (gpurun_out/ is capped at 64 MiB).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
OURS='regex:gemm_bf16_kernel|all_gather_push|reduce_scatter|all_reduce|all_to_all|tile_reduce|qkv_rope|rmsnorm|layernorm|swiglu|bias_gelu|ce_rowmax|ce_sumexp|ce_bwd|cast_kernel'
COMMON="--no-cpu-baseline --no-probe --steps 1 --warmup 1 --layers 1 --strategy configs/ncu_2layers_1microbatch.json"
export_rep() {   # $1 = report stem
  if [ -f "$1.ncu-rep" ]; then
    ncu -i "$1.ncu-rep" --page raw --csv 2>/dev/null | gzip -9 > "$1_raw.csv.gz"
    ncu -i "$1.ncu-rep" --page details --csv 2>/dev/null | gzip -9 > "$1_details.csv.gz"
    ls -la "$1.ncu-rep" "$1_raw.csv.gz" "$1_details.csv.gz"
    rm -f "$1.ncu-rep"
  fi
}
timeout 400 ncu --set full --clock-control none --profile-from-start off -k "$OURS" -c 48 -f -o gpurun_out/r02_ncu_llama_1layer \
    python bench.py $COMMON --ncu-step > gpurun_out/r02_ncu_llama_1layer.log 2>&1
echo "== llama rc=$?"; tail -2 gpurun_out/r02_ncu_llama_1layer.log | cut -c1-200
export_rep gpurun_out/r02_ncu_llama_1layer
timeout 200 ncu --set full --clock-control none --profile-from-start off -k 'regex:layernorm|bias_gelu' -c 8 -f -o gpurun_out/r02_ncu_gpt_rows \
    python bench.py $COMMON --ncu-step --model gpt-6.7b --seq 2048 > gpurun_out/r02_ncu_gpt_rows.log 2>&1
echo "== gpt rc=$?"; tail -2 gpurun_out/r02_ncu_gpt_rows.log | cut -c1-200
export_rep gpurun_out/r02_ncu_gpt_rows
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r02_ncu_launches_step_layers2.csv \
    python bench.py --no-cpu-baseline --no-probe --steps 1 --warmup 1 --layers 2 --ncu-step > gpurun_out/r02_ncu_launches_step_layers2.log 2>&1
echo "== launches rc=$?"; wc -l gpurun_out/r02_ncu_launches_step_layers2.csv; gzip -9 gpurun_out/r02_ncu_launches_step_layers2.csv
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
