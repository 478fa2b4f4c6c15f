#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.log 2>&1; tail -c 600 gpurun_out/b_bench.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --frames 1 > gpurun_out/b_bench_f1.log 2>&1; tail -c 300 gpurun_out/b_bench_f1.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --frames 16 > gpurun_out/b_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc8 -s 4 -c 1 -o gpurun_out/b_prof_tc8 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --frames 16 > gpurun_out/b_ncu1.log 2>&1
