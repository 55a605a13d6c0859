#!/bin/bash
set -u
mkdir -p gpurun_out
for layer in ${TRACE_LAYERS}; do
  MC_TRACE_LAYER=$layer PROF_PASSES=2 timeout 300 python scripts/prof_forward.py 2>&1 | grep "\[trace" | tail -1
  MC_DIAG=3 MC_TRACE_LAYER=$layer PROF_PASSES=2 timeout 300 python scripts/prof_forward.py 2>&1 | grep "\[trace" | tail -1 | sed 's/^/DIAG3 /'
done
