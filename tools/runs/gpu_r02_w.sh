#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --frames 16 --no-cpu-baseline --no-e2e --no-configs"
$B > gpurun_out/w_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/w_launches_tc8.csv $B > gpurun_out/w_ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1_tc8 -s 3 -c 1 -o gpurun_out/w_prof_tc8 -f $B > gpurun_out/w_ncu2.log 2>&1
tail -c 300 gpurun_out/w_plain.log
ncu -i gpurun_out/w_prof_tc8.ncu-rep --page source --csv > gpurun_out/w_src.csv 2>/dev/null
python tools/ncu_roles.py gpurun_out/w_src.csv $((16*3840*2160)) > gpurun_out/w_roles.txt; head -30 gpurun_out/w_roles.txt | cut -c1-200
