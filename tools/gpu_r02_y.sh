#!/bin/bash
for i in 1 2 3; do
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc.py 16 10
timeout 60 python tools/time_tc.py 16 10
done
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc.py 512 20
timeout 60 python tools/time_tc.py 512 20
DCTC_LIB=tools/exp/libdctc_oldstore.so timeout 60 python tools/time_tc.py 1 40
timeout 60 python tools/time_tc.py 1 40
