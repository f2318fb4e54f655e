#!/bin/bash
# Second (last) round-1 GPU call: the failing plugin test with its full traceback, one ncu --set full capture
# of the packed and the scalar scan at the same working point, and the bench line with the new defaults.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check2.log) 2>&1
date
timeout 90 python -m pytest tests/test_gpu_zz_plugin.py -x -q --tb=long -k unicycle --timeout 80 > gpurun_out/plugin_unicycle.log 2>&1; tail -30 gpurun_out/plugin_unicycle.log | cut -c1-400
echo "== ncu full: packed then scalar scan"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:nearest_diag -c 2 -o gpurun_out/ncu_k1_r1 -f python tools/k1_once.py > gpurun_out/ncu_k1.log 2>&1; tail -3 gpurun_out/ncu_k1.log
echo "== bench, new defaults"
timeout 150 python bench.py --steps 10 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_r1_defaults.json 2> gpurun_out/bench_err2.log; tail -c 400 gpurun_out/bench_r1_defaults.json
date
