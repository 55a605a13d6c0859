#!/bin/bash
# ncu passes on the GPU box (1 GPU).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
NK=66     # kernels per eager forward+decode pass
# (1) launch list of one eager step: every kernel with its device time (cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $NK -c $NK --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/prof_forward.py > gpurun_out/${TAG}_launches_stdout.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${TAG}_launches.csv | cut -c1-200
# (2) all metrics (--set full) of every convolution launch of the second pass, as a raw CSV (small)
timeout 1500 ncu --set full --clock-control none -k regex:conv_tc -s 50 -c 50 --csv --page raw --log-file gpurun_out/${TAG}_conv_raw.csv \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_raw_stdout.log 2>&1
echo "conv raw rc=$?"; ls -la gpurun_out/${TAG}_conv_raw.csv
# (3) full captures with source: one streamed-weight halo launch (level3 3x3) and the head stems (resident-weight halo kernel)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 25 -c 1 -o gpurun_out/${TAG}_conv_tc3 \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_tc3_stdout.log 2>&1
echo "tc3 capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 27 -c 1 -o gpurun_out/${TAG}_conv_tc2 \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_tc2_stdout.log 2>&1
echo "tc2 capture rc=$?"; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
