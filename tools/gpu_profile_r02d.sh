#!/bin/bash
# Final round-2 ncu evidence of the bench workload (1 GPU) with the kernels as committed at the end of the round:
# launch list of the bench command, then a full capture of one forward and one backward launch of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --no-strips"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_ -s 6 -c 2 -f -o gpurun_out/prof_r02d $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
