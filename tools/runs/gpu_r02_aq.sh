#!/bin/bash
mkdir -p gpurun_out
for b in 2 16; do
B="python bench.py --steps 2 --warmup 3 --frames 4 --blocksize $b --no-cpu-baseline --no-e2e --no-configs"
$B > gpurun_out/aq_plain$b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_ -s 3 -c 1 -o gpurun_out/aq_prof_b$b -f $B > gpurun_out/aq_ncu$b.log 2>&1
tail -1 gpurun_out/aq_ncu$b.log | cut -c1-200
done
