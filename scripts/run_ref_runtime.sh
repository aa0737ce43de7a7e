#!/bin/bash
# Runs the UNMODIFIED reference runtime (baseline/_ref, under oracle/ref_shim) on this box's GPUs for the fixture cases of
# oracle/ref_runtime/run_ref.py; outputs land in gpurun_out/ref_runtime_<case>.json (+ .log).   usage: run_ref_runtime.sh case:world ...
mkdir -p gpurun_out
port=29900
for cw in "$@"; do
  c=${cw%%:*}; w=${cw##*:}; port=$((port+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $port \
      oracle/ref_runtime/run_ref.py --case $c > gpurun_out/ref_runtime_$c.log 2>&1
  echo "== $c rc=$?"; grep "REF_RUNTIME" gpurun_out/ref_runtime_$c.log | cut -c1-300 || true
  grep -v "Warning\|warn" gpurun_out/ref_runtime_$c.log | tail -4 | cut -c1-400
done
