#!/bin/bash
# Round-1f ncu evidence for the loss / metric epilogue and the tile scheduler / blended merge kernels.
mkdir -p gpurun_out
python tools/prof_te.py > gpurun_out/plain_te.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'loss_l1|dem_metrics|tiles_' -s 10 -c 5 -f -o gpurun_out/prof_te python tools/prof_te.py > gpurun_out/ncu_te.log 2>&1
echo "te exit: $?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'loss_l1|dem_metrics|tiles_' -c 15 --csv --log-file gpurun_out/launches_te.csv python tools/prof_te.py > gpurun_out/ncu_te_launches.log 2>&1
echo "launch list exit: $?"
