#!/bin/bash
mkdir -p gpurun_out
for a in "-4 0 32 8" "-8 0 32 8" "-6 0 32 8" "-1 0 32 8" "0 -6 32 8" "0 -1 32 8" "6 0 32 8" "1 0 32 8" "0 6 32 8" "100 0 32 8" "0 125 32 8" "-32 0 32 8" "-4 -4 144 29" "0 0 32 8 -1"; do
  echo -n "c0 c1 bw bh [c2] = $a : "; timeout 60 ./tools/tma_probe 100 $a 2>&1 | tail -1; done > gpurun_out/tma_probe2.log 2>&1
cat gpurun_out/tma_probe2.log
