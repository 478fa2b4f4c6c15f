#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/final_bench_n1.log 2>&1; tail -c 300 gpurun_out/final_bench_n1.log
