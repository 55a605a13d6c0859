#!/bin/bash
# ncu passes on the GPU box (1 GPU), round 2.  Numbers printed under ncu are never bench values.
#   bash scripts/gpu_profile.sh <tag> [precision]      precision: fp32 (default, the fp32-accurate tensor-core mode) | bf16
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
export PROF_PRECISION=${2:-fp32}
NCU="ncu --clock-control none --profile-from-start off"
# (1) launch list of one eager step: every kernel with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/prof_forward.py > gpurun_out/${TAG}_launches_stdout.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${TAG}_launches.csv | cut -c1-200
# (2) DRAM bytes + tensor-pipe + throughput metrics of EVERY convolution launch of the step (a short metric list: few replays)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
M=$M,sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed
M=$M,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed
timeout 1500 $NCU --metrics $M -k regex:"conv_tc|head_apply" --csv --page raw --log-file gpurun_out/${TAG}_conv_raw.csv \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_raw_stdout.log 2>&1
echo "conv metrics rc=$?"; ls -la gpurun_out/${TAG}_conv_raw.csv
# (3) --set full with source of three representative launches: a streamed-weight halo layer (level4 3x3), the stem, a 1x1 Root
timeout 900 $NCU --set full --import-source on -k regex:conv_tc3 -s 12 -c 1 -o gpurun_out/${TAG}_conv_tc3 \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_tc3_stdout.log 2>&1
echo "tc3 capture rc=$?"
timeout 900 $NCU --set full --import-source on -k regex:conv_tc2 -s 0 -c 1 -o gpurun_out/${TAG}_conv_tc2 \
    python scripts/prof_forward.py > gpurun_out/${TAG}_conv_tc2_stdout.log 2>&1
echo "tc2 capture rc=$?"; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
