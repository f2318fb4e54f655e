#!/bin/bash
# Round 2, call 15 (2 GPUs): multi-GPU parity tests and a 2-GPU bench line on the final build (sampler ahead of the
# chain, seed by kernel argument, multi-CTA closest-to-guide).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call15.log) 2>&1
date; nvidia-smi --query-gpu=name --format=csv,noheader
echo "== multi-GPU tests"
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 2>&1 | tail -4
echo "== bench, 1 and 2 GPUs (digests must match)"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/final_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --plan-leg 0 2>&1 | tail -1 > gpurun_out/final_n2.json
python - <<'PY'
import json
for n in (1, 2):
    try:
        d=json.load(open('gpurun_out/final_n%d.json' % n))
        print('N=%d'%d['n_gpus'], round(d['value']/1e6,2),'M nodes/s', round(d['ms_per_step'],2),'ms/plan  e2e',round(d['e2e']['value']/1e6,2), d['e2e']['host_ms_per_step'], 'digest',d.get('tree_digest'), 'clocks', d['clocks'], 'roofline', d['roofline']['frac'], d['roofline'].get('alone'))
    except Exception as e:
        print('N=%d unreadable' % n, e, open('gpurun_out/final_n%d.json' % n).read()[-800:])
PY
date
