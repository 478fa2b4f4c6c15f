#!/bin/bash
set -x
mkdir -p gpurun_out
DCTC_LIB=$PWD/tools/exp/libdctc_dbg16.so timeout 300 python tools/check_tc16.py --quick > gpurun_out/k_dbg.log 2>&1; echo "rc $?" >> gpurun_out/k_dbg.log
grep -o "tag [0-9]* parity [0-9]" gpurun_out/k_dbg.log | sort | uniq -c
grep -v "mbar timeout" gpurun_out/k_dbg.log | tail -40
if grep -q "CHECK_TC16 PASS" gpurun_out/k_dbg.log; then
timeout 600 python tools/check_tc16.py > gpurun_out/k_check16.log 2>&1; echo "rc $?" >> gpurun_out/k_check16.log
tail -12 gpurun_out/k_check16.log
fi
