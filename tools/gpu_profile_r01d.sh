#!/bin/bash
# Round-1d ncu evidence: launch list + full capture of the bench workload, the generator-tail kernel, the grad_init backward.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch-list exit: $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_ -s 6 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full-capture exit: $?"
python tools/prof_gen.py > gpurun_out/plain_gen.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_spn -s 2 -c 1 -f -o gpurun_out/prof_gen python tools/prof_gen.py > gpurun_out/ncu_gen.log 2>&1
echo "gen exit: $?"
python tools/prof_case.py f32 2048 128 128 gi > gpurun_out/plain_gi.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_backward -s 2 -c 1 -f -o gpurun_out/prof_gi python tools/prof_case.py f32 2048 128 128 gi > gpurun_out/ncu_gi.log 2>&1
echo "gi exit: $?"
ls -la gpurun_out/*.ncu-rep
