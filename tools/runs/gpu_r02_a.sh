#!/bin/bash
# round-2 GPU call A: baseline evidence (tests, bench, ncu of the benched launch, seam-loop kernels)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.log 2>&1; tail -c 1500 gpurun_out/a_bench.log
python tools/bench_retarget.py > gpurun_out/a_retarget.log 2>&1; tail -c 1200 gpurun_out/a_retarget.log
# ncu: the exact launch the bench times (16 frames of 4K), full set with source
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc8 -s 4 -c 1 -o gpurun_out/a_prof_tc8 -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu1.log 2>&1
# ncu: seam DP kernel (config 3), 2 launches in steady state
python tools/bench_retarget.py 40 > gpurun_out/a_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:seam_dp -s 60 -c 1 -o gpurun_out/a_prof_dp -f \
    python tools/bench_retarget.py 40 > gpurun_out/a_ncu2.log 2>&1
# ncu: b=4 streaming kernel and b=16 tile kernel, 4 frames
python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e > gpurun_out/a_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_small -s 3 -c 1 -o gpurun_out/a_prof_b4 -f \
    python bench.py --steps 2 --warmup 3 --frames 4 --blocksize 4 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu3.log 2>&1
ls -la gpurun_out
