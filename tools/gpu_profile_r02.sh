#!/bin/bash
# Round-2 ncu evidence: the weight-gradient kernel of the Generator tail.
mkdir -p gpurun_out
python tools/prof_gw.py > gpurun_out/plain_gw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_grad_weight -s 1 -c 1 -f -o gpurun_out/prof_gw python tools/prof_gw.py > gpurun_out/ncu_gw.log 2>&1
echo "gw exit: $?"
python tools/prof_iter_fused.py > gpurun_out/plain_if.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_iterate_fused -s 1 -c 1 -f -o gpurun_out/prof_if python tools/prof_iter_fused.py > gpurun_out/ncu_if.log 2>&1
echo "iter fused exit: $?"
