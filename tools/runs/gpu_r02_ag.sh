#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 16 --blocksize 4 --no-cpu-baseline --no-e2e --no-configs"
$B > gpurun_out/ag_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_tc4 -s 3 -c 1 -o gpurun_out/ag_prof_tc4 -f $B > gpurun_out/ag_ncu2.log 2>&1
tail -c 300 gpurun_out/ag_plain.log
ncu -i gpurun_out/ag_prof_tc4.ncu-rep --page source --csv > gpurun_out/ag_src.csv 2>/dev/null
python tools/ncu_roles.py gpurun_out/ag_src.csv $((16*3840*2160)) > gpurun_out/ag_roles.txt; head -30 gpurun_out/ag_roles.txt | cut -c1-200
