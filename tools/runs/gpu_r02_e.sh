#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_k1_gpu.py tests/test_k2_gpu.py tests/test_fullsize_gpu.py tests/test_multi_gpu.py -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/e_pytest.log
tail -8 gpurun_out/e_pytest.log
for b in 4 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --frames 64 --blocksize $b > gpurun_out/e_bench_b$b.log 2>&1; tail -c 200 gpurun_out/e_bench_b$b.log; done
python tools/bench_retarget.py > gpurun_out/e_retarget.log 2>&1; tail -c 900 gpurun_out/e_retarget.log
DCTC_LIB=$PWD/tools/exp/libdctc_dbg.so python tools/bench_retarget.py 4 2>&1 | grep "dp kernel w 19" | tail -4 > gpurun_out/e_dbg.log
DCTC_LIB=$PWD/tools/exp/libdctc_dbg_noex.so python tools/bench_retarget.py 4 2>&1 | grep "dp kernel w 19" | tail -4 > gpurun_out/e_dbg_noex.log
cat gpurun_out/e_dbg.log gpurun_out/e_dbg_noex.log
python tools/bench_retarget.py 40 > gpurun_out/e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/e_launches_seamloop.csv python tools/bench_retarget.py 40 > gpurun_out/e_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/e_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_small -s 3 -c 1 -o gpurun_out/e_prof_b4 -f \
    python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/e_ncu3.log 2>&1
