#!/bin/bash
# Round 2, call 1: baseline sanity of the round-1 engine + the measurements round 1 left open
# (prefix-filter scan, non-identity weights tests, rollout latency attribution).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call1.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"
timeout 400 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "== opt-in tests"
LQRRT_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_zz_options.py -x -q --timeout 150 -k "prefix or non_identity" 2>&1 | tail -5
echo "== scan sweep"
timeout 150 python tools/gpu_sweep_k1.py 2>&1 | grep -v '"what": "k1"' | tail -14
python - <<'PY'
import json
rows = [json.loads(l) for l in open('gpurun_out/sweep_k1_r1.jsonl')]
for r in rows:
    if r.get('what') == 'k1' and r['ctas_per_sm'] in (8, 16):
        print('%-4s N %-8d %-10s ctas %2d  %.4f ms  %.3e pairs/s  same=%s' % (r['config'], r['N'], r['kernel'], r['ctas_per_sm'], r['ms'], r['pairs_per_s'], r['bit_identical_to_scalar']))
PY
echo "== rollout latency attribution"
timeout 150 python tools/steer_probe.py 2>&1 | tail -24
date
