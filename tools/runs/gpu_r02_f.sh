#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_k1_gpu.py tests/test_k2_gpu.py tests/test_fullsize_gpu.py -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
python tools/bench_retarget.py > gpurun_out/f_retarget.log 2>&1; tail -c 700 gpurun_out/f_retarget.log
for f in 1 16; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --frames $f > gpurun_out/f_bench_f$f.log 2>&1; tail -c 250 gpurun_out/f_bench_f$f.log; done
for v in "" t16p4 t16p2 t16w16; do
  if [ -n "$v" ]; then export DCTC_LIB=$PWD/tools/exp/libdctc_$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --frames 8 --blocksize 16 > gpurun_out/f_bench_b16_$v.log 2>&1; tail -c 200 gpurun_out/f_bench_b16_$v.log
done
unset DCTC_LIB
