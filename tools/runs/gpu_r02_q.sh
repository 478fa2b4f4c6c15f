#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_small.py 2>&1 | tee gpurun_out/q_dbg.log | cut -c1-150
( python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py tests/test_multi_gpu.py -m gpu -q ) > gpurun_out/q_tests.log 2>&1; tail -5 gpurun_out/q_tests.log | cut -c1-200
for b in 2 4; do python tools/time_small.py $b 16 10; python tools/time_small.py $b 64 5; done 2>&1 | tee gpurun_out/q_time.log
python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/q_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_small -s 3 -c 1 -o gpurun_out/q_prof_b4 -f \
    python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/q_ncu.log 2>&1
tail -2 gpurun_out/q_ncu.log | cut -c1-300
