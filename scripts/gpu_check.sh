#!/bin/bash
# Runs on the GPU box (via gpurun): smoke, GPU parity tests, a short bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-run}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/${TAG}_smoke.log
echo "== tests" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1 ; echo "tests rc=$?" ; tail -40 gpurun_out/${TAG}_tests.log
echo "== bench" ; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --stage-table gpurun_out/${TAG}_stages.csv > gpurun_out/${TAG}_bench.log 2>&1 ; echo "bench rc=$?" ; tail -3 gpurun_out/${TAG}_bench.log
