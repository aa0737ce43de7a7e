#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_collectives.py tests/test_gpu_gemm.py -q -m gpu -x > gpurun_out/test_gpu_quick.log 2>&1; echo "exit quick: $?"; tail -3 gpurun_out/test_gpu_quick.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
