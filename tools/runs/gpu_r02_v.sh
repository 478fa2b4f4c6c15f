#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py tests/test_multi_gpu.py -m gpu -q -x ) > gpurun_out/v_tests.log 2>&1; tail -5 gpurun_out/v_tests.log | cut -c1-200
timeout 120 python tools/check_tc.py 2>&1 | tail -3
timeout 60 python tools/time_tc.py 16 10; timeout 60 python tools/time_tc.py 1 40
DCTC_TC_NO_TENSORMAP=1 timeout 60 python tools/time_tc.py 16 10; DCTC_TC_NO_TENSORMAP=1 timeout 60 python tools/time_tc.py 1 40
timeout 60 python tools/time_tc.py 16 10; DCTC_TC_NO_TENSORMAP=1 timeout 60 python tools/time_tc.py 16 10
