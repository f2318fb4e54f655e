#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check4.log) 2>&1
date
timeout 80 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -4
timeout 45 python bench.py --steps 10 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_err4.log; tail -c 300 gpurun_out/bench_r1_final2.json
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench4.log 2>&1
date
