#!/bin/bash
# Round 2, call 2: first run of the device-scheduled wave pipeline.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call2.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== new pipeline tests first"
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q --timeout 500 2>&1 | tail -15
echo "== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_pipeline.py 2>&1 | tail -25
echo "== bench (device schedule)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -3
echo "== bench (host schedule)"
LQRRT_SCHEDULE=host timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -3
date
