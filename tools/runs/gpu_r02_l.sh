#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2; do
python tools/bench_retarget.py > gpurun_out/l_retarget.log 2>&1; grep -o '"device_loop_us_per_seam": [0-9.]*' gpurun_out/l_retarget.log
DCTC_NO_PDL=1 python tools/bench_retarget.py > gpurun_out/l_retarget_nopdl.log 2>&1; grep -o '"device_loop_us_per_seam": [0-9.]*' gpurun_out/l_retarget_nopdl.log
DCTC_LIB=$PWD/tools/exp/libdctc_pdlw.so python tools/bench_retarget.py > gpurun_out/l_retarget_pdlw.log 2>&1; grep -o '"device_loop_us_per_seam": [0-9.]*' gpurun_out/l_retarget_pdlw.log
done
