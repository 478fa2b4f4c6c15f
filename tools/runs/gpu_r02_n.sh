#!/bin/bash
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/n_tests.log 2>&1; tail -5 gpurun_out/n_tests.log
( time python bench.py ) > gpurun_out/n_bench.log 2>&1; tail -c 600 gpurun_out/n_bench.log
( time python bench.py --impl reference ) > gpurun_out/n_bench_ref.log 2>&1; tail -c 600 gpurun_out/n_bench_ref.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; tail -2 gpurun_out/n_smoke.log
