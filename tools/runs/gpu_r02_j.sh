#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_k2_gpu.py tests/test_fullsize_gpu.py -q -x -k "seam or retarget or c3 or carver or incremental or enlarg or vmap" > gpurun_out/j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j_pytest.log
tail -5 gpurun_out/j_pytest.log
python tools/bench_retarget.py > gpurun_out/j_retarget.log 2>&1; tail -c 400 gpurun_out/j_retarget.log
python tools/bench_retarget.py 40 > gpurun_out/j_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/j_launches_seamloop.csv python tools/bench_retarget.py 40 > gpurun_out/j_ncu.log 2>&1
