#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py tests/test_multi_gpu.py -m gpu -q -x ) > gpurun_out/u_tests.log 2>&1; tail -5 gpurun_out/u_tests.log | cut -c1-200
python tools/check_tc16.py 2>&1 | tail -5
python tools/time_tc16.py 16 5; DCTC_EDGES=0.8 DCTC_TEXTURES=0.2 python tools/time_tc16.py 16 5
