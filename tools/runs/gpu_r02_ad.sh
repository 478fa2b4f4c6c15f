#!/bin/bash
for i in 1 2; do
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc.py 1 40
timeout 60 python tools/time_tc.py 1 40
done
timeout 60 python tools/time_tc.py 16 10
timeout 120 python tools/time_tc_ramp.py 8 | head -3
timeout 120 python tools/time_tc_ramp.py 16 | head -3
timeout 60 python tools/time_tc16.py 1 10 | tail -2
timeout 120 python tools/check_tc.py 2>&1 | tail -1
timeout 120 python tools/check_tc16.py 2>&1 | tail -1
