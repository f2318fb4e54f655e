#!/bin/bash
# Round 2, call 24 (1 GPU, short): scan grid target on the 12-state whitened scan (C5).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call24.log) 2>&1
date
for c in 4 8 16; do
  LQRRT_SCAN_CTAS=$c timeout 120 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu --plan-leg 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 scan CTAs/SM $c:', round(d['value']/1e6,3),'M nodes/s', round(d['ms_per_step'],1),'ms/plan', {k:round(v/d['steps'],1) for k,v in d['stage_ms'].items()})"
done
date
