#!/bin/bash
# Round 2, call 7: full GPU suite on the current build + ncu of the chain kernels at the multi-GPU working point.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call7.log) 2>&1
date
echo "== full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15
echo "== chain kernels at W = 1024 (one GPU of eight)"
PIPE_PLANS=3 timeout 120 python tools/pipe_once.py
echo "== ncu chain kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'steer_x|commit_x' -s 80 -c 4 -o gpurun_out/chain_r2 python tools/pipe_once.py > gpurun_out/chain_ncu.log 2>&1; tail -2 gpurun_out/chain_ncu.log
date
