#!/bin/bash
mkdir -p gpurun_out
for b in 8 2 4 16; do python tools/time_tc_ramp.py $b; done 2>&1 | tee gpurun_out/s_ramp.log
