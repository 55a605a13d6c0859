#!/bin/bash
# ncu --set full of the non-convolution kernels of one eager pass (raw CSV).
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
timeout 900 ncu --set full --clock-control none -k regex:'head_apply|attn_stats|attn_mix|decode|upsample2|pack_input|maxpool2' -s 15 -c 15 --csv --page raw \
    --log-file gpurun_out/${TAG}_misc_raw.csv python scripts/prof_forward.py > gpurun_out/${TAG}_misc_stdout.log 2>&1
echo "misc raw rc=$?"; ls -la gpurun_out/${TAG}_misc_raw.csv
