#!/bin/bash
set -x
mkdir -p gpurun_out
( python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q ) > gpurun_out/o_tests.log 2>&1; tail -15 gpurun_out/o_tests.log
for b in 2 4; do python tools/time_small.py $b 16 10; python tools/time_small.py $b 64 5; DCTC_EDGES=0.8 DCTC_TEXTURES=0.2 python tools/time_small.py $b 16 10; done 2>&1 | tee gpurun_out/o_time.log
