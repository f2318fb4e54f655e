set -x
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --cpu-seconds 15 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 2500 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nearest_diag -s 50 -c 2 -o gpurun_out/prof_nearest_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full1.log 2>&1; tail -2 gpurun_out/ncu_full1.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:steer_kernel|sample_kernel' -s 100 -c 2 -o gpurun_out/prof_steer_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log
python bench.py --impl reference --steps 1 --warmup 0 --cpu-seconds 10 > gpurun_out/bench_r1_ref.json 2>&1; tail -c 1500 gpurun_out/bench_r1_ref.json
ls -la gpurun_out
