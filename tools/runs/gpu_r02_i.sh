#!/bin/bash
set -x
mkdir -p gpurun_out
for v in a b c d; do
  DCTC_LIB=$PWD/tools/exp/libdctc_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --frames 8 --blocksize 16 > gpurun_out/i_bench_b16_$v.log 2>&1; tail -c 200 gpurun_out/i_bench_b16_$v.log
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench.log 2>&1; tail -c 300 gpurun_out/i_bench.log
