#!/bin/bash
for i in 1 2; do
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc16.py 1 10 | tail -2
timeout 60 python tools/time_tc16.py 1 10 | tail -2
done
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 120 python tools/time_tc_ramp.py 16 | head -2
timeout 120 python tools/time_tc_ramp.py 16 | head -2
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
