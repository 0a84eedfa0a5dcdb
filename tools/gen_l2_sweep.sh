#!/bin/bash
mkdir -p gpurun_out
for a in 0 1 2 3 4 6 8 12; do
  echo "== L2_AHEAD=$a"
  JSPSR_GEN_L2_AHEAD=$a python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from jspsr_b200 import functional as F
from tools.quick_bench import timeit, PEAK
B, H, W, C = 2048, 128, 128, 64
init = torch.rand(B, 1, H, W, device="cuda"); feat = torch.randn(B, C, H, W, device="cuda")
cw = torch.randn(25, C, device="cuda") * 0.15; cb = torch.randn(25, device="cuda") * 0.1
w = torch.ones(1, 1, 3, 3, device="cuda"); b = torch.zeros(1, device="cuda")
npx = B * H * W
m, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, False), n=20)
m2, _ = timeit(lambda: F.gen_spn_forward(init, feat, cw, cb, w, b, 1, 1.0, True), n=20)
print(f"fused {m*1e3:.1f} us ({npx*264/m/1e6/PEAK:.3f})  with w/o out {m2*1e3:.1f} us ({npx*372/m2/1e6/PEAK:.3f})")
PY
done 2>&1 | tee gpurun_out/gen_l2_sweep.log
for a in 0 2 4; do
JSPSR_GEN_L2_AHEAD=$a ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gen_spn -s 2 -c 1 python tools/prof_gen.py 2>&1 | grep -E "dram__bytes_read|gpu__time" | sed "s/^/ahead=$a /"
done | tee -a gpurun_out/gen_l2_sweep.log
