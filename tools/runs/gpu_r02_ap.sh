#!/bin/bash
DCTC_TC_WIDE=1 timeout 300 python tools/check_tc.py 2>&1 | tail -26
timeout 60 python tools/time_tc.py 16 10
DCTC_TC_WIDE=1 timeout 60 python tools/time_tc.py 16 10
DCTC_TC_WIDE=1 timeout 60 python tools/time_tc.py 1 40
