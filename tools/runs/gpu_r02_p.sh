#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_small.py 2>&1 | tee gpurun_out/p_dbg.log
( python -m pytest tests/test_k1_gpu.py tests/test_fullsize_gpu.py tests/test_multi_gpu.py -m gpu -q ) > gpurun_out/p_tests.log 2>&1; tail -15 gpurun_out/p_tests.log | cut -c1-200
