#!/bin/bash
mkdir -p gpurun_out
python tools/prof_gen.py > gpurun_out/plain_gen.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_spn -s 2 -c 1 -f -o gpurun_out/prof_gen python tools/prof_gen.py > gpurun_out/ncu_gen.log 2>&1
echo "gen exit $?"; tail -3 gpurun_out/ncu_gen.log
