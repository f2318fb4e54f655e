#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check3.log) 2>&1
date
timeout 80 python tools/gpu_sweep_k1.py 2>&1 | grep -v '"what": "k1"' | tail -16
timeout 60 python -m pytest tests/test_gpu_zz_plugin.py tests/test_gpu_zz_options.py -x -q --timeout 50 2>&1 | tail -4
date
