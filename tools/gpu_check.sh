#!/bin/bash
# One GPU session: parity tests, smoke, a short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -c 3000 gpurun_out/bench.log
