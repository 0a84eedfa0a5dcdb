#!/bin/bash
# Round-2 ncu evidence: launch list + full capture of the bench workload (1 GPU), the Generator tail's weight-gradient
# kernel, the single-launch fixed-affinity loop.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --no-strips"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_ -s 6 -c 2 -f -o gpurun_out/prof_r02 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
python tools/prof_gw.py > gpurun_out/plain_gw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_grad_weight -s 1 -c 1 -f -o gpurun_out/prof_gw python tools/prof_gw.py > gpurun_out/ncu_gw.log 2>&1
echo "gw exit: $?"
python tools/prof_iter_fused.py > gpurun_out/plain_if.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_iterate_fused -s 1 -c 1 -f -o gpurun_out/prof_if python tools/prof_iter_fused.py > gpurun_out/ncu_if.log 2>&1
echo "iter fused exit: $?"
