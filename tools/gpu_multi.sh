#!/bin/bash
# multi-GPU session (gpurun --gpus N): bench at N, strip inference at N
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N exit $?"
grep '^{' gpurun_out/bench_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks','gpu_launches')}); print(d['roofline'])"
H=$((4096*N))
timeout 600 $RUN tools/strip_bench.py --H $H --W 32768 --T 1 > gpurun_out/strip_n${N}_T1.log 2>&1; echo "strip T=1 exit $?"; grep -v "^W\|^\[" gpurun_out/strip_n${N}_T1.log | tail -3
timeout 600 $RUN tools/strip_bench.py --H $H --W 32768 --T 6 > gpurun_out/strip_n${N}_T6.log 2>&1; echo "strip T=6 exit $?"; grep -v "^W\|^\[" gpurun_out/strip_n${N}_T6.log | tail -3
