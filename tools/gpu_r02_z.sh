#!/bin/bash
DCTC_LIB=tools/exp/libdctc_timing.so timeout 120 python tools/time_tc.py 16 1 2>&1 | tail -14
