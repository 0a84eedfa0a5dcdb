#!/bin/bash
# ncu evidence for the bench workload: (1) launch list with device times, (2) one full capture of our kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_ -s 6 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
tail -3 gpurun_out/plain.log | cut -c1-400; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/
