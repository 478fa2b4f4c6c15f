#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x ) > gpurun_out/x_tests.log 2>&1; tail -3 gpurun_out/x_tests.log | cut -c1-200
timeout 60 python tools/time_tc.py 16 10; timeout 60 python tools/time_tc.py 1 40; timeout 60 python tools/time_tc.py 16 10
DCTC_EDGES=0.8 DCTC_TEXTURES=0.2 timeout 60 python tools/time_tc.py 16 10
timeout 60 python tools/time_tc.py 512 20
