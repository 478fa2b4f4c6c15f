#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/k_pytest.log
tail -15 gpurun_out/k_pytest.log
timeout 300 python tools/check_tc16.py > gpurun_out/k_check16.log 2>&1; echo "rc $?" >> gpurun_out/k_check16.log
tail -7 gpurun_out/k_check16.log
