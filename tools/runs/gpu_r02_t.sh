#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x ) > gpurun_out/t_tests.log 2>&1; tail -5 gpurun_out/t_tests.log | cut -c1-200
python tools/check_tc.py 2>&1 | tail -6
python tools/time_tc.py 16 10; python tools/time_tc.py 1 40; DCTC_EDGES=0.8 DCTC_TEXTURES=0.2 python tools/time_tc.py 16 10
python tools/time_tc.py 512 20
