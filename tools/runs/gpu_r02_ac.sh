#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/check_tc.py 2>&1 | tail -2
for i in 1 2; do
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc.py 16 10
timeout 60 python tools/time_tc.py 16 10
done
timeout 60 python tools/time_tc.py 1 40
timeout 60 python tools/time_tc.py 512 20
( timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x ) > gpurun_out/ac_tests.log 2>&1; tail -3 gpurun_out/ac_tests.log | cut -c1-200
