#!/bin/bash
# Per-layer activation memory of the runtime, measured the way the reference's memory profiler does it (runs with different layer
# counts, subtract): Llama-3-8B layers, seq 8192, microbatch 1, no checkpointing, with and without --recompute_activations.
# Then the kernel-by-kernel breakdown of one full step (CUPTI, shares only).  One GPU.  Outputs -> gpurun_out/.
mkdir -p gpurun_out
for mode in plain recompute; do
  strat=configs/galvatron_config_llama3-8b_1gpus.json
  [ $mode = recompute ] && strat=configs/galvatron_config_llama3-8b_1gpus_recompute.json
  for L in 4 12; do
    timeout 200 python bench.py --no-cpu-baseline --no-probe --steps 1 --warmup 3 --layers $L --checkpoint-layers 0 --strategy $strat \
        > gpurun_out/r02_mem_${mode}_L$L.json 2> gpurun_out/r02_mem_${mode}_L$L.err
    python - "$mode" "$L" <<'P'
import json, sys
d = json.load(open("gpurun_out/r02_mem_%s_L%s.json" % (sys.argv[1], sys.argv[2])))
print("MEM", sys.argv[1], sys.argv[2], d["memory_gib"], d["ms_per_step"])
P
  done
done
timeout 400 python bench.py --no-cpu-baseline --no-probe --steps 2 --warmup 3 --kernel-breakdown gpurun_out/r02_kernel_breakdown_n1.json \
    ${BENCH_STRATEGY:+--strategy $BENCH_STRATEGY} > gpurun_out/r02_bench_n1_breakdown_run.json 2> gpurun_out/r02_bench_n1_breakdown_run.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_kernel_breakdown_n1.json"))
print(d["event_timed_ms_per_step"], d["sum_kernel_ms"])
for k in d["kernels"][:28]:
    print("%8.2f ms %5.1f%% %6d  %s" % (k["ms"], 100 * k["share_of_sum"], k["launches"], k["name"][:90]))
P
