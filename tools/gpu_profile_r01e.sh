#!/bin/bash
# Round-1e ncu evidence for the generator-tail kernels (final versions) and the NLSPN affinity backward.
mkdir -p gpurun_out
python tools/prof_gen.py > gpurun_out/plain_gen.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_spn -s 2 -c 1 -f -o gpurun_out/prof_gen python tools/prof_gen.py > gpurun_out/ncu_gen.log 2>&1
echo "gen exit: $?"
python tools/prof_gf.py > gpurun_out/plain_gf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gen_grad_feature -s 1 -c 1 -f -o gpurun_out/prof_gf python tools/prof_gf.py > gpurun_out/ncu_gf.log 2>&1
echo "gf exit: $?"
python tools/prof_aff.py > gpurun_out/plain_aff.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nlspn_affinity_bwd -s 1 -c 1 -f -o gpurun_out/prof_aff python tools/prof_aff.py > gpurun_out/ncu_aff.log 2>&1
echo "aff exit: $?"
