#!/bin/bash
# usage: gpu_prof_case.sh <tag> <args to prof_case.py...>
tag=$1; shift
mkdir -p gpurun_out
python tools/prof_case.py "$@" > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spn_ -s 4 -c 2 -f -o gpurun_out/prof_$tag python tools/prof_case.py "$@" > gpurun_out/ncu_$tag.log 2>&1
echo "$tag exit $?"
