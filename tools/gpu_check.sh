#!/bin/bash
# One gpurun call near the end of round 1: correctness of the new option / plugin paths first, then the A/B
# measurements, then (if the budget allows) the full GPU suite and a bench line with the candidate defaults.
# Everything lands in gpurun_out/ incrementally.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
date
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== options tests"; timeout 240 python -m pytest tests/test_gpu_zz_options.py -x -q --timeout 200 2>&1 | tail -5
echo "== A/B k1 + plan"; timeout 200 python tools/gpu_ab.py k1 plan 2>&1 | tail -12
echo "== plugin tests"; timeout 240 python -m pytest tests/test_gpu_zz_plugin.py -x -q --timeout 200 2>&1 | tail -5
echo "== A/B wave policy + c5k"; timeout 200 python tools/gpu_ab.py wave c5k 2>&1 | tail -10
echo "== bench with candidate defaults"
LQRRT_NEAREST=packed LQRRT_PREFETCH=1 timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_r1_packed_prefetch.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_r1_packed_prefetch.json
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "== ncu launch list (packed + prefetch)"
LQRRT_NEAREST=packed LQRRT_PREFETCH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1_packed.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full capture of the packed scan"
LQRRT_NEAREST=packed timeout 300 ncu --set full --clock-control none --import-source on -k regex:nearest_diag_pk -s 30 -c 1 -o gpurun_out/ncu_nearest_pk_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
date
