#!/bin/bash
# session-2 probe: (1) does TMA accept 128-B (not 1024-B) aligned swizzled destinations?  (2) role wait traces
set -u
mkdir -p gpurun_out
echo "== ASPLIT=2 conv parity"; MC_TC2_ASPLIT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "conv_kernel_parity and bf16" -p no:cacheprovider 2>&1 | tail -5
echo "== ASPLIT=18 conv parity"; MC_TC2_ASPLIT=18 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "conv_kernel_parity and bf16" -p no:cacheprovider 2>&1 | tail -5
echo "== ASPLIT=18 bench"; MC_TC2_ASPLIT=18 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
TRACE_LAYERS="neck.ida_2.node_1 backbone.level2.tree1.conv2 head.stems backbone.level0" bash scripts/gpu_trace.sh
