#!/bin/bash
# ncu passes on the GPU box (1 GPU).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
# (1) launch list of one eager step: every kernel with its device time (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 201 -c 70 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_launches_stdout.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/${TAG}_launches.csv
# (2) full capture of the convolution kernel over one forward pass (50 launches)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 150 -c 50 -o gpurun_out/${TAG}_conv_tc \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_conv_tc_stdout.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/${TAG}_conv_tc.ncu-rep
