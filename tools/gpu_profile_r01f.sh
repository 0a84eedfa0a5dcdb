#!/bin/bash
# Round-1f ncu evidence for the loss / metric epilogue and the tile scheduler / blended merge kernels.
mkdir -p gpurun_out
python tools/prof_te.py > gpurun_out/plain_te.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'loss_l1|dem_metrics|tiles_' -s 8 -c 4 -f -o gpurun_out/prof_te python tools/prof_te.py > gpurun_out/ncu_te.log 2>&1
echo "te exit: $?"
