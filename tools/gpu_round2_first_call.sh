#!/bin/bash
# Ready-made first GPU call of round 2 (everything below was written after round 1's GPU budget was spent):
#   1. the tensor-core bound prototype: self-check + pairs/s                      (tools/tc_filter)
#   2. the opt-in tests: prefix-filter scan, non-identity weights                   (LQRRT_EXPERIMENTAL=1)
#   3. scan sweep including the prefix variant                                      (tools/gpu_sweep_k1.py)
#   4. rollout latency attribution                                                  (tools/steer_probe.py)
# usage: gpurun --timeout 400 -- 'bash tools/gpu_round2_first_call.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/round2_first_call.log) 2>&1
date
echo "== 1. tensor-core bound prototype"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/tc_filter tools/tc_filter/tc_filter.cu && {
  timeout 60 /tmp/tc_filter 8192 100000 200; timeout 60 /tmp/tc_filter 8192 20000 60; timeout 60 /tmp/tc_filter 1000 777 30; }
echo "== 2. opt-in tests"
LQRRT_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_zz_options.py -x -q --timeout 150 -k "prefix or non_identity" 2>&1 | tail -5
echo "== 3. scan sweep"
timeout 120 python tools/gpu_sweep_k1.py 2>&1 | grep -v '"what": "k1"' | tail -14
python - <<'PY'
import json
rows = [json.loads(l) for l in open('gpurun_out/sweep_k1_r1.jsonl')]
for r in rows:
    if r.get('what') == 'k1' and r['ctas_per_sm'] in (8, 16):
        print('%-4s N %-8d %-10s ctas %2d  %.4f ms  %.3e pairs/s  same=%s' % (r['config'], r['N'], r['kernel'], r['ctas_per_sm'], r['ms'], r['pairs_per_s'], r['bit_identical_to_scalar']))
PY
echo "== 4. rollout latency attribution"
timeout 120 python tools/steer_probe.py 2>&1 | tail -24
date
