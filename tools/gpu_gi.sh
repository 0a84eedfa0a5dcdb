#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gi.log
python tools/gi_sweep.py 2>&1 | tee gpurun_out/gi_sweep.log
