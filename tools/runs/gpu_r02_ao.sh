#!/bin/bash
timeout 60 python tools/time_tc.py 16 10
for a in 1 2 3; do echo "ablation $a"; DCTC_LIB=tools/exp/libdctc_abl$a.so timeout 60 python tools/time_tc.py 16 10; done
timeout 60 python tools/time_tc.py 16 10
