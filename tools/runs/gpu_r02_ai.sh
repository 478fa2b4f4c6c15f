#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_k1_gpu.py -m gpu -q -k "stream" ) > gpurun_out/ai_tests.log 2>&1; tail -15 gpurun_out/ai_tests.log | cut -c1-220
