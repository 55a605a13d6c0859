#!/bin/bash
# session-2: tcgen05 issue-loop timing probe + first run of the streamed-weight halo kernel (conv_tc3.cu)
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a tools/umma_timing.cu -o /tmp/umma_timing && timeout 120 /tmp/umma_timing > gpurun_out/umma_timing.log 2>&1; echo "probe rc=$?"; cat gpurun_out/umma_timing.log
echo "== conv parity (tc3 default)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_kernel_parity and bf16" -p no:cacheprovider 2>&1 | tail -15
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
AB_CONFIGS="base MC_TC3=0 MC_TC3_NT=256 MC_TC3_SUB=4 MC_TC3_SUB=1 MC_TC3_ROWPAD=1" bash scripts/gpu_ab.sh
