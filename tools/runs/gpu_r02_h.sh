#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "rc $?" >> gpurun_out/h_smoke.log; tail -3 gpurun_out/h_smoke.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/h_pytest.log; tail -5 gpurun_out/h_pytest.log
python tools/bench_retarget.py > gpurun_out/h_retarget.log 2>&1; tail -c 400 gpurun_out/h_retarget.log
DCTC_DP_P=2 python tools/bench_retarget.py > gpurun_out/h_retarget_p2.log 2>&1; tail -c 400 gpurun_out/h_retarget_p2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.log 2>&1; echo "rc $?" >> gpurun_out/h_bench.log; tail -c 600 gpurun_out/h_bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.log 2>&1; tail -c 400 gpurun_out/h_bench_ref.log
