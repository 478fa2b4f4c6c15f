#!/bin/bash
for i in 1 2; do
timeout 60 python tools/time_small.py 4 16 10
DCTC_LIB=tools/exp/libdctc_b4six.so timeout 60 python tools/time_small.py 4 16 10
timeout 60 python tools/time_small.py 4 64 5
DCTC_LIB=tools/exp/libdctc_b4six.so timeout 60 python tools/time_small.py 4 64 5
done
