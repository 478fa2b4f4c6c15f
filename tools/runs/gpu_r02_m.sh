#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/m_bench.log 2>&1; tail -c 400 gpurun_out/m_bench.log
python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 16 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/m_plain16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/m_launches_tc16.csv \
    python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 16 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/m_ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_tc16 -s 3 -c 1 -o gpurun_out/m_prof_tc16 -f \
    python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 16 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/m_ncu2.log 2>&1
tail -3 gpurun_out/m_plain16.log
