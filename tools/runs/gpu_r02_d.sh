#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_k2_gpu.py tests/test_fullsize_gpu.py -q -x -k "seam or retarget or c3 or carver or incremental" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/d_pytest.log
tail -8 gpurun_out/d_pytest.log
python tools/bench_retarget.py > gpurun_out/d_retarget.log 2>&1; tail -c 900 gpurun_out/d_retarget.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --frames 16 > gpurun_out/d_bench16.log 2>&1; tail -c 300 gpurun_out/d_bench16.log
python tools/bench_retarget.py 40 > gpurun_out/d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/d_launches_seamloop.csv python tools/bench_retarget.py 40 > gpurun_out/d_ncu.log 2>&1
tail -30 gpurun_out/d_launches_seamloop.csv | cut -c1-200
