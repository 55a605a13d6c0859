#!/bin/bash
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a tools/umma_timing.cu -o /tmp/umma_timing && timeout 120 /tmp/umma_timing > gpurun_out/umma_timing2.log 2>&1; echo "probe rc=$?"; tail -24 gpurun_out/umma_timing2.log
for layer in backbone.level3.tree1.tree2.conv1 backbone.level4.tree2.tree1.conv1 neck.ida_1.node_1 backbone.level5.tree2.conv1; do
  MC_TRACE_LAYER=$layer PROF_PASSES=2 timeout 300 python scripts/prof_forward.py 2>&1 | grep "\[trace" | tail -1
  MC_TC3_NT=256 MC_TRACE_LAYER=$layer PROF_PASSES=2 timeout 300 python scripts/prof_forward.py 2>&1 | grep "\[trace" | tail -1 | sed 's/^/NT256 /'
done
