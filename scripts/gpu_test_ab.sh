#!/bin/bash
# parity tests, then bench A/B over AB_CONFIGS, then role traces of TRACE_LAYERS (MC_TRACE_LAYER)
set -u
mkdir -p gpurun_out
echo "== conv parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_kernel_parity and bf16" -p no:cacheprovider 2>&1 | tail -3
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
AB_CONFIGS="${AB_CONFIGS:-base}" bash scripts/gpu_ab.sh
for layer in ${TRACE_LAYERS:-}; do
  MC_TRACE_LAYER=$layer PROF_PASSES=2 timeout 300 python scripts/prof_forward.py 2>&1 | grep "\[trace" | tail -1
done
